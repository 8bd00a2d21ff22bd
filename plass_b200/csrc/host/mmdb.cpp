#include "mmdb.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <numeric>
#include <omp.h>
#include <parallel/algorithm>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mmdb {

static int g_threads = 0;

int hostThreads() {
    if (g_threads > 0) return g_threads;
    int n = 0;
    if (const char *e = getenv("MMSEQS_NUM_THREADS")) n = atoi(e);     // Parameters.cpp:2124
    if (n <= 0) n = omp_get_num_procs();
    g_threads = std::max(1, std::min(n, 256));
    return g_threads;
}
void setHostThreads(int n) { if (n > 0) g_threads = std::min(n, 256); }

static bool fileExists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

static bool slurp(const std::string &p, std::vector<char> &out, size_t at) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(at + (size_t) sz);
    const size_t got = sz ? fread(out.data() + at, 1, (size_t) sz, f) : 0;
    fclose(f);
    return got == (size_t) sz;
}

void Reader::unmapAll() {
    if (mapped) { munmap(mapped, mappedBytes); mapped = nullptr; }
    for (const Seg &sg : segs) if (sg.p && sg.len) munmap(const_cast<char *>(sg.p), sg.len);
    segs.clear();
}

Reader::~Reader() { unmapAll(); }

bool Reader::open(const std::string &path, std::string &err, bool contiguous) {
    unmapAll();
    base = nullptr; bytes = 0;
    if (fileExists(path)) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) ::close(fd); err = "cannot read " + path; return false; }
        bytes = (size_t) st.st_size;
        if (bytes) {
            mapped = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
            if (mapped == MAP_FAILED) { mapped = nullptr; ::close(fd); err = "cannot map " + path; return false; }
            mappedBytes = bytes;
            madvise(mapped, bytes, MADV_WILLNEED);
            base = (const char *) mapped;
        }
        ::close(fd);
    } else {
        // split data files X.0 .. X.k (DBWriter::close without merge): one anonymous mapping, filled by all host threads
        std::vector<std::string> files;
        std::vector<size_t> at(1, 0);
        while (fileExists(path + "." + std::to_string(files.size()))) {
            struct stat st;
            files.push_back(path + "." + std::to_string(files.size()));
            if (stat(files.back().c_str(), &st) != 0) { err = "cannot read split " + path; return false; }
            at.push_back(at.back() + (size_t) st.st_size);
        }
        if (files.empty()) { err = "database " + path + " not found"; return false; }
        bytes = at.back();
        if (bytes && !contiguous) {
            // result DBs are only read entry by entry: map every file where it is, no copy
            for (size_t f = 0; f < files.size(); f++) {
                const size_t len = at[f + 1] - at[f];
                Seg sg{nullptr, (uint64_t) at[f], len};
                if (len) {
                    const int fd = ::open(files[f].c_str(), O_RDONLY);
                    void *m = fd < 0 ? MAP_FAILED : mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
                    if (fd >= 0) ::close(fd);
                    if (m == MAP_FAILED) { err = "cannot map " + files[f]; return false; }
                    madvise(m, len, MADV_WILLNEED);
                    sg.p = (const char *) m;
                }
                if (len) segs.push_back(sg);
            }
        } else if (bytes) {
            mapped = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (mapped == MAP_FAILED) { mapped = nullptr; err = "cannot allocate " + std::to_string(bytes) + " bytes for " + path; return false; }
            mappedBytes = bytes;
            base = (const char *) mapped;
            // pieces of 8 MB over all files
            struct Piece { size_t file; size_t off; size_t len; };
            std::vector<Piece> pieces;
            for (size_t f = 0; f < files.size(); f++)
                for (size_t o = 0; o < at[f + 1] - at[f]; o += (8u << 20)) pieces.push_back({f, o, std::min<size_t>(8u << 20, at[f + 1] - at[f] - o)});
            std::vector<int> fds(files.size());
            for (size_t f = 0; f < files.size(); f++) fds[f] = ::open(files[f].c_str(), O_RDONLY);
            bool ok = true;
#pragma omp parallel for num_threads(hostThreads()) schedule(dynamic, 1)
            for (size_t k = 0; k < pieces.size(); k++) {
                const Piece &pc = pieces[k];
                size_t done = 0;
                while (done < pc.len) {
                    const ssize_t r = fds[pc.file] < 0 ? -1 : pread(fds[pc.file], (char *) mapped + at[pc.file] + pc.off + done, pc.len - done, (off_t) (pc.off + done));
                    if (r <= 0) { ok = false; break; }
                    done += (size_t) r;
                }
            }
            for (int fd : fds) if (fd >= 0) ::close(fd);
            if (!ok) { err = "cannot read split " + path; return false; }
        }
    }
    std::vector<char> idx;
    if (!slurp(path + ".index", idx, 0)) { err = "cannot read " + path + ".index"; return false; }
    // index parse by all host threads: the text is cut at line ends, every thread counts its lines, then parses them
    // into their final positions
    const int T = hostThreads();
    const char *ib = idx.data();
    const size_t isz = idx.size();
    std::vector<size_t> cut((size_t) T + 1, isz), cnt((size_t) T + 1, 0);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        size_t p = std::max(cut[t - 1], isz * (size_t) t / (size_t) T);
        while (p < isz && p > 0 && ib[p - 1] != '\n') p++;
        cut[t] = p;
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        size_t c = 0;
        const char *p = ib + cut[t], *e = ib + cut[t + 1];
        while (p < e) {
            const char *nl = (const char *) memchr(p, '\n', (size_t) (e - p));
            const char *le = nl ? nl : e;
            if (le > p) c++;                        // blank lines are skipped
            p = nl ? nl + 1 : e;
        }
        cnt[(size_t) t + 1] = c;
    }
    for (int t = 0; t < T; t++) cnt[(size_t) t + 1] += cnt[t];
    const size_t n = cnt[T];
    keys.assign(n, 0); offsets.assign(n, 0); lens.assign(n, 0);
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        size_t i = cnt[t];
        const char *p = ib + cut[t], *e = ib + cut[t + 1];
        while (p < e) {
            if (*p == '\n') { p++; continue; }
            uint64_t v[3] = {0, 0, 0};
            for (int c = 0; c < 3; c++) {
                while (p < e && (*p < '0' || *p > '9')) { if (*p == '\n') break; p++; }
                while (p < e && *p >= '0' && *p <= '9') { v[c] = v[c] * 10 + (uint64_t) (*p - '0'); p++; }
            }
            while (p < e && *p != '\n') p++;
            if (p < e) p++;
            keys[i] = (uint32_t) v[0]; offsets[i] = v[1]; lens[i] = (uint32_t) v[2];
            i++;
        }
    }
    bool sorted = true;
#pragma omp parallel for num_threads(T) reduction(&& : sorted)
    for (size_t i = 1; i < n; i++) sorted = sorted && keys[i - 1] <= keys[i];
    if (!sorted) {
        // kmermatcher's close(false, false) leaves pref_N.index unsorted (DBReader::open re-sorts by id, DBReader.cpp:238-253)
        std::vector<uint64_t> order(n);
#pragma omp parallel for num_threads(T)
        for (size_t i = 0; i < n; i++) order[i] = ((uint64_t) keys[i] << 32) | (uint64_t) i;
        // (key, original position) pairs are distinct, so the order equals a stable sort by key
        __gnu_parallel::sort(order.begin(), order.end());
        std::vector<uint32_t> k2(n), l2(n);
        std::vector<uint64_t> o2(n);
#pragma omp parallel for num_threads(T)
        for (size_t i = 0; i < n; i++) {
            const size_t j = (size_t) (order[i] & 0xFFFFFFFFull);
            k2[i] = keys[j]; o2[i] = offsets[j]; l2[i] = lens[j];
        }
        keys.swap(k2); offsets.swap(o2); lens.swap(l2);
    }
    bool inside = true;
#pragma omp parallel for num_threads(T) reduction(&& : inside)
    for (size_t i = 0; i < n; i++) inside = inside && (offsets[i] + lens[i] <= bytes);
    if (!inside) { err = "index of " + path + " points outside the data file"; return false; }
    FILE *ft = fopen((path + ".dbtype").c_str(), "rb");
    if (!ft) { err = "cannot read " + path + ".dbtype"; return false; }
    int32_t t = 0;
    if (fread(&t, 4, 1, ft) != 1) { fclose(ft); err = "short dbtype file"; return false; }
    fclose(ft);
    if (t & (1 << 31)) { err = path + " is zstd-compressed (--compressed 1): not supported by the GPU commands"; return false; }
    dbtype = t & 0xFFFF;
    return true;
}

bool Writer::open(const std::string &p, int dbtype, std::string &err, bool splitData) {
    path = p;
    split = splitData;
    // data files of an earlier DB of this name (either layout) must not survive next to the new one
    ::unlink(p.c_str());
    for (int k = 0; fileExists(p + "." + std::to_string(k)); k++) ::unlink((p + "." + std::to_string(k)).c_str());
    fds.clear(); fileBytes.clear(); ents.clear();
    if (split) {
        fds.assign((size_t) hostThreads(), -1);
        fileBytes.assign(fds.size(), 0);
        for (size_t k = 0; k < fds.size(); k++) {
            fds[k] = ::open((p + "." + std::to_string(k)).c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fds[k] < 0) { err = "cannot open " + p + "." + std::to_string(k) + " for writing"; return false; }
        }
        fd = -1;
    } else {
        fd = ::open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    }
    fi = ::open((p + ".index").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if ((!split && fd < 0) || fi < 0) { err = "cannot open " + p + " for writing"; return false; }
    FILE *ft = fopen((p + ".dbtype").c_str(), "wb");
    if (!ft) { err = "cannot write dbtype"; return false; }
    const int32_t t = dbtype;
    fwrite(&t, 4, 1, ft);
    fclose(ft);
    offset = 0; indexOffset = 0; failed = false;
    pendData.clear(); pendIndex.clear();
    return true;
}

static bool pwriteAll(int fd, const char *p, size_t n, uint64_t at) {
    while (n) {
        const ssize_t w = pwrite(fd, p, n, (off_t) at);
        if (w <= 0) return false;
        p += w; n -= (size_t) w; at += (uint64_t) w;
    }
    return true;
}

static inline char *putU64(char *b, unsigned long long v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
    while (n) *b++ = tmp[--n];
    return b;
}

static inline void indexLine(std::string &out, uint32_t key, uint64_t off, uint64_t len) {
    char line[72];
    char *b = putU64(line, key); *b++ = '\t';
    b = putU64(b, off); *b++ = '\t';
    b = putU64(b, len); *b++ = '\n';
    out.append(line, (size_t) (b - line));
}

static void flushPending(Writer &w) {
    if (!w.pendData.empty()) {
        if (w.split) {
            if (!pwriteAll(w.fds[0], w.pendData.data(), w.pendData.size(), w.fileBytes[0])) w.failed = true;
            w.fileBytes[0] += w.pendData.size();
        } else if (!pwriteAll(w.fd, w.pendData.data(), w.pendData.size(), w.offset - w.pendData.size())) w.failed = true;
        w.pendData.clear();
    }
    if (!w.pendIndex.empty()) {
        if (!pwriteAll(w.fi, w.pendIndex.data(), w.pendIndex.size(), w.indexOffset)) w.failed = true;
        w.indexOffset += w.pendIndex.size();
        w.pendIndex.clear();
    }
}

void Writer::write(uint32_t key, const char *bytes, size_t n) {
    if (split) ents.push_back({key, (uint32_t) (n + 1), 0u, fileBytes[0] + pendData.size()});
    else indexLine(pendIndex, key, offset, n + 1);
    pendData.append(bytes, n);
    pendData.push_back('\0');
    offset += n + 1;
    if (pendData.size() > (8u << 20)) flushPending(*this);
}

void Writer::writeAll(size_t n, const std::function<uint32_t(size_t)> &keyOf, const std::function<void(size_t, std::string &)> &format,
                      const std::function<bool(size_t)> &skip) {
    flushPending(*this);
    if (n == 0) return;
    const int T = hostThreads();
    size_t chunk = n / ((size_t) T * 8) + 1;
    chunk = std::max<size_t>(1024, std::min<size_t>(chunk, 65536));
    if (split) {
        // a data file per thread: every thread formats its chunks and appends them to ITS file -- no shared offsets, no barriers,
        // no two threads on one inode.  The index is written by close().
        const size_t entBase = ents.size();
        ents.resize(entBase + n);
        const size_t nChunks = (n + chunk - 1) / chunk;
        bool badS = false;
#pragma omp parallel num_threads(std::min<int>(T, (int) fds.size()))
        {
            const size_t t = (size_t) omp_get_thread_num();
            std::string d;
#pragma omp for schedule(static, 1)
            for (size_t c = 0; c < nChunks; c++) {
                d.clear();
                const size_t lo = c * chunk, hi = std::min(n, lo + chunk);
                uint64_t o = fileBytes[t];
                for (size_t i = lo; i < hi; i++) {
                    Ent &e = ents[entBase + i];
                    if (skip && skip(i)) { e.len = 0xFFFFFFFFu; continue; }
                    const size_t before = d.size();
                    format(i, d);
                    d.push_back('\0');
                    e.key = keyOf(i); e.len = (uint32_t) (d.size() - before); e.file = (uint32_t) t; e.off = o + before;
                }
                if (!d.empty()) {
                    if (!pwriteAll(fds[t], d.data(), d.size(), o)) {
#pragma omp atomic write
                        badS = true;
                    }
                    fileBytes[t] = o + d.size();
                }
            }
        }
        if (badS) failed = true;
        return;
    }
    const size_t wave = chunk * (size_t) T;
    std::vector<std::string> dbuf((size_t) T), ibuf((size_t) T);
    std::vector<std::vector<uint32_t>> elen((size_t) T);
    std::vector<uint64_t> dBase((size_t) T + 1), iBase((size_t) T + 1);
    bool bad = false;
#pragma omp parallel num_threads(T)
    {
        for (size_t w0 = 0; w0 < n; w0 += wave) {
#pragma omp for schedule(static, 1)
            for (int c = 0; c < T; c++) {
                std::string &d = dbuf[(size_t) c];
                std::vector<uint32_t> &el = elen[(size_t) c];
                d.clear(); el.clear();
                const size_t lo = std::min(n, w0 + (size_t) c * chunk), hi = std::min(n, lo + chunk);
                for (size_t i = lo; i < hi; i++) {
                    if (skip && skip(i)) { el.push_back(0xFFFFFFFFu); continue; }
                    const size_t before = d.size();
                    format(i, d);
                    d.push_back('\0');
                    el.push_back((uint32_t) (d.size() - before));
                }
            }
#pragma omp single
            {
                uint64_t o = offset;
                for (int c = 0; c < T; c++) { dBase[(size_t) c] = o; o += dbuf[(size_t) c].size(); }
                offset = o;
            }
#pragma omp for schedule(static, 1)
            for (int c = 0; c < T; c++) {
                const std::string &d = dbuf[(size_t) c];
                if (!d.empty() && !pwriteAll(fd, d.data(), d.size(), dBase[(size_t) c])) {
#pragma omp atomic write
                    bad = true;
                }
                std::string &ix = ibuf[(size_t) c];
                ix.clear();
                const size_t lo = std::min(n, w0 + (size_t) c * chunk);
                uint64_t o = dBase[(size_t) c];
                const std::vector<uint32_t> &el = elen[(size_t) c];
                for (size_t j = 0; j < el.size(); j++) {
                    if (el[j] == 0xFFFFFFFFu) continue;
                    indexLine(ix, keyOf(lo + j), o, el[j]);
                    o += el[j];
                }
            }
#pragma omp single
            {
                uint64_t o = indexOffset;
                for (int c = 0; c < T; c++) { iBase[(size_t) c] = o; o += ibuf[(size_t) c].size(); }
                indexOffset = o;
            }
#pragma omp for schedule(static, 1)
            for (int c = 0; c < T; c++) {
                const std::string &ix = ibuf[(size_t) c];
                if (!ix.empty() && !pwriteAll(fi, ix.data(), ix.size(), iBase[(size_t) c])) {
#pragma omp atomic write
                    bad = true;
                }
            }
        }
    }
    if (bad) failed = true;
}

void Writer::writeContiguous(const char *data, uint64_t bytes, size_t n, const uint32_t *keys, const uint64_t *offsets, const uint32_t *lens) {
    flushPending(*this);
    if (split || offset != 0 || indexOffset != 0) { failed = true; return; }
    const int T = hostThreads();
    std::vector<std::string> ibuf((size_t) T);
    bool bad = false;
#pragma omp parallel num_threads(T)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        if (t == 0) {
            for (uint64_t o = 0; o < bytes; o += (64u << 20)) {
                if (!pwriteAll(fd, data + o, (size_t) std::min<uint64_t>(64u << 20, bytes - o), o)) {
#pragma omp atomic write
                    bad = true;
                }
            }
        }
        // the index: every thread but the writer takes an equal share (all of it when there is only one thread)
        const int workers = nt > 1 ? nt - 1 : 1, me = nt > 1 ? t - 1 : 0;
        if (me >= 0) {
            std::string &ix = ibuf[(size_t) (nt > 1 ? t : 0)];
            const size_t lo = n * (size_t) me / (size_t) workers, hi = n * ((size_t) me + 1) / (size_t) workers;
            ix.reserve((hi - lo) * 24);
            for (size_t i = lo; i < hi; i++) indexLine(ix, keys[i], offsets[i], lens[i]);
        }
    }
    uint64_t at = 0;
    for (int t = 0; t < T; t++) {
        const std::string &ix = ibuf[(size_t) t];
        if (!ix.empty() && !pwriteAll(fi, ix.data(), ix.size(), at)) bad = true;
        at += ix.size();
    }
    offset = bytes; indexOffset = at;
    if (bad) failed = true;
}

bool Writer::close() {
    flushPending(*this);
    bool ok = !failed;
    if (split) {
        // files without data disappear, the others keep their order under consecutive numbers (a single one is named X);
        // index offsets count through the files in that order
        std::vector<uint64_t> base(fds.size(), 0);
        std::vector<int> newIdx(fds.size(), -1);
        int kept = 0;
        uint64_t run = 0;
        for (size_t k = 0; k < fds.size(); k++) {
            if (fds[k] >= 0) ok &= ::close(fds[k]) == 0;
            if (fileBytes[k]) { newIdx[k] = kept++; base[k] = run; run += fileBytes[k]; }
        }
        for (size_t k = 0; k < fds.size(); k++) {
            const std::string from = path + "." + std::to_string(k);
            if (newIdx[k] < 0) { ::unlink(from.c_str()); continue; }
            const std::string to = kept == 1 ? path : path + "." + std::to_string(newIdx[k]);
            if (to != from) ok &= ::rename(from.c_str(), to.c_str()) == 0;
        }
        if (kept == 0) {
            const int e = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);      // an empty DB is an empty data file
            if (e >= 0) ::close(e); else ok = false;
        }
        fds.clear();
        // the index, in the order the entries were handed over (ascending keys)
        const int T = hostThreads();
        const size_t n = ents.size();
        std::vector<std::string> ibuf((size_t) T);
        std::vector<uint64_t> at((size_t) T + 1, 0);
#pragma omp parallel for num_threads(T) schedule(static, 1)
        for (int t = 0; t < T; t++) {
            std::string &ix = ibuf[(size_t) t];
            const size_t lo = n * (size_t) t / (size_t) T, hi = n * ((size_t) t + 1) / (size_t) T;
            ix.reserve((hi - lo) * 24);
            for (size_t i = lo; i < hi; i++) {
                const Ent &e = ents[i];
                if (e.len == 0xFFFFFFFFu) continue;
                indexLine(ix, e.key, base[e.file] + e.off, e.len);
            }
        }
        for (int t = 0; t < T; t++) at[(size_t) t + 1] = at[(size_t) t] + ibuf[(size_t) t].size();
        bool badI = false;
#pragma omp parallel for num_threads(T) schedule(static, 1)
        for (int t = 0; t < T; t++) {
            const std::string &ix = ibuf[(size_t) t];
            if (!ix.empty() && !pwriteAll(fi, ix.data(), ix.size(), indexOffset + at[(size_t) t])) {
#pragma omp atomic write
                badI = true;
            }
        }
        if (badI) ok = false;
        ents.clear(); ents.shrink_to_fit();
    }
    if (fd >= 0) ok &= ::close(fd) == 0;
    if (fi >= 0) ok &= ::close(fi) == 0;
    fd = fi = -1;
    return ok;
}

}  // namespace mmdb
