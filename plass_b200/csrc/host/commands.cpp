// commands.cpp -- host side of the drop-in commands: flag parsing, MMseqs DB in/out, text formatting.
// All arithmetic of the hot path happens behind the C ABI (libplassgpu.so); there is no CPU fallback.
#include "commands.h"
#include "mmdb.h"
#include "plassgpu.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <omp.h>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

namespace {

bool pathExists(const std::string &p) { struct stat st; return lstat(p.c_str(), &st) == 0; }

// DBReader::softlinkDb for one file: out -> basename-relative link to in, if in exists
void linkIfExists(const std::string &in, const std::string &out) {
    if (!pathExists(in)) return;
    if (pathExists(out)) unlink(out.c_str());
    char *real = realpath(in.c_str(), nullptr);
    if (real) { if (symlink(real, out.c_str()) != 0) { /* best effort, as the reference */ } free(real); }
}

struct Flags {
    std::vector<std::string> positional;
    std::map<std::string, std::string> kv;
};

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "Error: %s\n", msg.c_str());
    fflush(stdout); fflush(stderr);
    // EXIT() of the reference: the process ends here with EXIT_FAILURE.  Not exit(): the CUDA start-up may still be running in
    // its helper thread, and static destructors (a joinable std::thread) would turn the error into an abort.
    _exit(EXIT_FAILURE);
}

Flags parseFlags(int argc, const char **argv, size_t nPositional, const std::set<std::string> &known) {
    Flags f;
    int i = 0;
    for (; i < argc && f.positional.size() < nPositional; i++) f.positional.push_back(argv[i]);
    if (f.positional.size() != nPositional) die("too few database arguments");
    for (; i < argc; i += 2) {
        const std::string k = argv[i];
        if (known.find(k) == known.end()) die("unknown parameter " + k);
        if (i + 1 >= argc) die("missing value for " + k);
        f.kv[k] = argv[i + 1];
    }
    return f;
}

// MultiParam "nucl:0.200,aa:0.000" (lib/mmseqs/src/commons/MultiParam.cpp)
std::string multi(const std::string &v, bool nucl) {
    if (v.find(':') == std::string::npos) return v;
    size_t p = 0;
    while (p < v.size()) {
        size_t c = v.find(',', p);
        if (c == std::string::npos) c = v.size();
        const std::string part = v.substr(p, c - p);
        const size_t col = part.find(':');
        if (col != std::string::npos) {
            const std::string name = part.substr(0, col);
            if ((nucl && name == "nucl") || (!nucl && name == "aa")) return part.substr(col + 1);
        }
        p = c + 1;
    }
    die("cannot parse multi-parameter " + v);
}

std::string get(const Flags &f, const std::string &k, const std::string &dflt) {
    auto it = f.kv.find(k);
    return it == f.kv.end() ? dflt : it->second;
}
int geti(const Flags &f, const std::string &k, int d) { return atoi(get(f, k, std::to_string(d)).c_str()); }
double getd(const Flags &f, const std::string &k, double d) { auto it = f.kv.find(k); return it == f.kv.end() ? d : strtod(it->second.c_str(), nullptr); }

void requireValue(const Flags &f, const std::string &k, const std::string &allowed, const char *why) {
    auto it = f.kv.find(k);
    if (it != f.kv.end() && it->second != allowed) die(k + " " + it->second + " is not supported by the GPU path (" + why + ")");
}

void checkSubMat(const Flags &f) {
    auto it = f.kv.find("--sub-mat");
    if (it == f.kv.end()) return;
    if (multi(it->second, true) != "nucleotide.out" || multi(it->second, false) != "blosum62.out")
        die("--sub-mat " + it->second + ": only the built-in nucleotide.out / blosum62.out tables are available on the GPU path");
}

// wall-clock of the phases of a command, printed next to the reference's "Time for processing" line
struct Phases {
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    std::string text;
    void lap(const char *name) {
        const auto now = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof(b), "%s%s %lld ms", text.empty() ? "" : ", ", name, (long long) std::chrono::duration_cast<std::chrono::milliseconds>(now - last).count());
        text += b;
        last = now;
    }
    void report() const { printf("Phases: %s (%d host threads)\n", text.c_str(), mmdb::hostThreads()); }
};

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void report() const {   // same wording as the reference (Application.cpp:38-43)
        const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
        printf("Time for processing: %lldh %lldm %llds %lldms\n", ms / 3600000, (ms / 60000) % 60, (ms / 1000) % 60, ms % 1000);
    }
};

// The CUDA context (driver start-up, module load: a few hundred ms) is created by a helper thread while the main thread
// opens the DBs and parses their indices; gpu() joins it at the first use.
struct GpuInit {
    std::thread worker;
    pg_context *ctx = nullptr;
    std::string error;
    bool started = false;
    void start() {
        if (started) return;
        started = true;
        // A single-GPU command initialises only its own device: the CUDA start-up grows with the number of devices the process
        // can see (2.5 s for the first process on a two-GPU box against 0.2 - 0.7 s on a one-GPU box).  Nothing has touched CUDA yet.
        const char *dev = getenv("PLASS_B200_DEVICE");
        int index = dev ? atoi(dev) : 0;
        if (!getenv("CUDA_VISIBLE_DEVICES") && index >= 0) {
            setenv("CUDA_VISIBLE_DEVICES", std::to_string(index).c_str(), 1);
            index = 0;
        }
        worker = std::thread([this, index] {
            if (pg_init(index, &ctx) != 0) error = pg_last_error();
        });
    }
    pg_context *get() {
        start();
        if (worker.joinable()) worker.join();
        if (!ctx) die(error.empty() ? "pg_init failed" : error);
        return ctx;
    }
};
GpuInit g_gpuInit;
pg_context *gpu() { return g_gpuInit.get(); }
void gpuWarmUp() { g_gpuInit.start(); }


// --threads N (Parameters.cpp:2124): host threads of the text layer (index parse, entry parse, formatting, pwrite)
void applyThreads(const Flags &f) {
    auto it = f.kv.find("--threads");
    if (it != f.kv.end() && atoi(it->second.c_str()) > 0) mmdb::setHostThreads(atoi(it->second.c_str()));
}

pg_seqdb *uploadSeqDb(const mmdb::Reader &r) {
    if (r.dbtype != mmdb::DBTYPE_AMINO_ACIDS && r.dbtype != mmdb::DBTYPE_NUCLEOTIDES) die("input is not a sequence database");
    if (r.segmented()) die("internal error: a sequence database must be opened as one contiguous buffer");
    pg_seqdb_view v;
    v.data = r.data(); v.data_bytes = r.dataBytes();
    v.offsets = r.offsets.data(); v.lens = r.lens.data(); v.keys = r.keys.data(); v.n = r.size(); v.dbtype = r.dbtype;
    pg_seqdb *db = nullptr;
    if (pg_seqdb_upload(gpu(), &v, &db) != 0) die(pg_last_error());
    return db;
}

inline char *putU(char *b, unsigned long long v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
    while (n) *b++ = tmp[--n];
    return b;
}
inline char *putI(char *b, long long v) {
    if (v < 0) { *b++ = '-'; return putU(b, (unsigned long long) (-v)); }
    return putU(b, (unsigned long long) v);
}
// Util::fastSeqIdToBuffer (Util.cpp:278-307) as it ends up on disk: "1.00" for exactly 1, else 0.ddd truncated
inline char *putSeqId(char *b, float seqId) {
    if (seqId == 1.0) { memcpy(b, "1.00", 4); return b + 4; }
    *b++ = '0'; *b++ = '.';
    if (seqId < 0.10) *b++ = '0';
    if (seqId < 0.01) *b++ = '0';
    return putI(b, (int) (seqId * 1000));
}

const std::set<std::string> KM_FLAGS = {"--sub-mat", "--alph-size", "--min-seq-id", "--kmer-per-seq", "--spaced-kmer-mode", "--spaced-kmer-pattern",
    "--kmer-per-seq-scale", "--adjust-kmer-len", "--mask", "--mask-lower-case", "--cov-mode", "-k", "-c", "--max-seq-len", "--hash-shift",
    "--split-memory-limit", "--include-only-extendable", "--ignore-multi-kmer", "--threads", "--compressed", "-v"};
const std::set<std::string> RS_FLAGS = {"--sub-mat", "--rescore-mode", "--wrapped-scoring", "--filter-hits", "-e", "-c", "-a", "--cov-mode",
    "--min-seq-id", "--min-aln-len", "--seq-id-mode", "--add-self-matches", "--sort-results", "--db-load-mode", "--threads", "--compressed", "-v"};
const std::set<std::string> EX_FLAGS = {"--min-seq-id", "--max-seq-len", "--keep-target", "--threads", "-v", "--rescore-mode", "--sub-mat", "--db-load-mode", "--compressed"};

// POD array without value-initialisation: the parsing threads are the first to touch their part of it
template <class T>
struct PodArray {
    T *p = nullptr;
    size_t n = 0;
    PodArray() = default;
    explicit PodArray(size_t count) : p((T *) malloc(sizeof(T) * (count + 1))), n(count) { if (!p) die("out of host memory"); }
    PodArray(PodArray &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    PodArray(const PodArray &) = delete;
    PodArray &operator=(const PodArray &) = delete;
    ~PodArray() { free(p); }
    const T *data() const { return p; }
    size_t size() const { return n; }
};

// Text entries -> records, by all host threads.  The entries are cut into contiguous ranges of (nearly) equal BYTES; a first
// sweep counts the lines of every range (memchr), so every thread parses straight into its slice of the one result array.
template <class T, class Fn>
PodArray<T> parseParallel(const mmdb::Reader &db, size_t skipLinesPerEntry, Fn parseEntry) {
    const int nT = mmdb::hostThreads();
    const size_t n = db.size();
    std::vector<size_t> cut((size_t) nT + 1, n), cnt((size_t) nT + 1, 0);
    cut[0] = 0;
    if (n) {
        // equal BYTES per range (entries in key order are not in file order: cumulate the lengths, not the offsets)
        const size_t blk = 4096, nBlk = (n + blk - 1) / blk;
        std::vector<uint64_t> blkBytes(nBlk + 1, 0);
#pragma omp parallel for num_threads(nT) schedule(static)
        for (size_t b = 0; b < nBlk; b++) {
            uint64_t sum = 0;
            for (size_t i = b * blk; i < std::min(n, (b + 1) * blk); i++) sum += db.lens[i];
            blkBytes[b + 1] = sum;
        }
        for (size_t b = 0; b < nBlk; b++) blkBytes[b + 1] += blkBytes[b];
        for (int t = 1; t < nT; t++) {
            const uint64_t want = blkBytes[nBlk] * (uint64_t) t / (uint64_t) nT;
            const size_t b = (size_t) (std::lower_bound(blkBytes.begin(), blkBytes.end(), want) - blkBytes.begin());
            cut[(size_t) t] = std::max(cut[(size_t) t - 1], std::min(n, b * blk));
        }
    }
#pragma omp parallel for num_threads(nT) schedule(static, 1)
    for (int t = 0; t < nT; t++) {
        size_t c = 0;
        for (size_t i = cut[(size_t) t]; i < cut[(size_t) t + 1]; i++) {
            const char *p = db.entry(i), *e = p + (db.lens[i] ? db.lens[i] - 1 : 0);
            size_t lines = 0;
            while (p < e) { const char *nl = (const char *) memchr(p, '\n', (size_t) (e - p)); lines++; if (!nl) break; p = nl + 1; }
            c += lines > skipLinesPerEntry ? lines - skipLinesPerEntry : 0;
        }
        cnt[(size_t) t + 1] = c;
    }
    for (int t = 0; t < nT; t++) cnt[(size_t) t + 1] += cnt[(size_t) t];
    PodArray<T> all(cnt[(size_t) nT]);
#pragma omp parallel for num_threads(nT) schedule(static, 1)
    for (int t = 0; t < nT; t++) {
        T *out = all.p + cnt[(size_t) t];
        for (size_t i = cut[(size_t) t]; i < cut[(size_t) t + 1]; i++) out = parseEntry(i, out);
        if (out != all.p + cnt[(size_t) t + 1]) die("malformed result DB: a line count changed between the two parse sweeps");
    }
    return all;
}

// decimal integer / the fields of a result line; faster than strtol for the well-formed lines the reference writes
static inline const char *getU(const char *s, uint32_t &v) {
    while (*s == ' ' || *s == '\t') s++;
    uint32_t x = 0;
    while (*s >= '0' && *s <= '9') { x = x * 10 + (uint32_t) (*s - '0'); s++; }
    v = x;
    return s;
}
static inline const char *getI(const char *s, int32_t &v) {
    while (*s == ' ' || *s == '\t') s++;
    const bool neg = *s == '-';
    if (neg || *s == '+') s++;
    uint32_t x;
    s = getU(s, x);
    v = neg ? -(int32_t) x : (int32_t) x;
    return s;
}

// The sequence identity column as Util::fastSeqIdToBuffer prints it (Util.cpp:278-307): "1.00" or "0." + three digits.  For
// these texts (float) (digits / 10^decimals) IS the correctly rounded float strtof returns -- checked for all 1001 + 101 values by
// tests/test_host_cpu.py -- so the common case needs no strtof; anything else (exponent, more digits) goes to strtof.
static inline const char *getSeqId(const char *s, float &v) {
    while (*s == ' ' || *s == '\t') s++;
    const char *p = s;
    uint32_t num = 0; int ip = 0, dec = 0;
    while (*p >= '0' && *p <= '9' && ip < 2) { num = num * 10 + (uint32_t) (*p - '0'); p++; ip++; }
    if (ip == 1 && *p == '.') {
        p++;
        while (*p >= '0' && *p <= '9' && dec < 4) { num = num * 10 + (uint32_t) (*p - '0'); p++; dec++; }
        if (dec >= 1 && dec <= 3 && (*p == '\t' || *p == ' ')) {
            static const double den[4] = {1.0, 10.0, 100.0, 1000.0};
            v = (float) ((double) num / den[dec]);
            return p;
        }
    }
    char *e;
    v = strtof(s, &e);
    return e;
}

// Matcher::parseAlignmentRecord over a whole alignment DB (Matcher.cpp:190-320), entries in key order.  withEvalue = false: the
// E-value column is skipped (evalue = 0): neither the extension nor findassemblystart reads it, and strtod is most of a line's cost.
PodArray<pg_aln> parseAlnDb(const mmdb::Reader &aln, bool withEvalue, bool checkSeqId = false) {
    return parseParallel<pg_aln>(aln, 0, [&](size_t i, pg_aln *out) {
        const char *s = aln.entry(i);
        while (*s) {
            pg_aln a; a.query = aln.keys[i];
            s = getU(s, a.target);
            s = getI(s, a.bits);
            const char *e = getSeqId(s, a.seq_id);
            if (checkSeqId && a.seq_id != strtof(s, nullptr)) die("iotest: the sequence-identity parser disagrees with strtof");
            if (withEvalue) { char *e2; a.evalue = strtod(e, &e2); e = e2; }
            else {
                a.evalue = 0.0;
                while (*e == ' ' || *e == '\t') e++;
                while (*e && *e != '\t' && *e != ' ' && *e != '\n') e++;
            }
            s = getI(e, a.q_start); s = getI(s, a.q_end); s = getI(s, a.q_len);
            s = getI(s, a.db_start); s = getI(s, a.db_end); s = getI(s, a.db_len);
            *out++ = a;
            while (*s && *s != '\n') s++;
            if (*s) s++;
        }
        return out;
    });
}

// QueryMatcher::parsePrefilterHits (QueryMatcher.h:81-112) over a kmermatcher result; the first line of every entry
// must be the self line "key\t0\t0"
PodArray<pg_hit> parsePrefDb(const mmdb::Reader &pref) {
    return parseParallel<pg_hit>(pref, 1, [&](size_t i, pg_hit *out) {
        const char *s = pref.entry(i);
        bool first = true;
        while (*s) {
            pg_hit h; h.rep = pref.keys[i];
            int32_t d;
            s = getU(s, h.target);
            s = getI(s, h.score);
            s = getI(s, d);
            h.diag = (int32_t) (short) d;
            if (first) {
                if (h.target != h.rep || h.score != 0 || h.diag != 0) die("prefilter entry does not start with its self line (not a kmermatcher result)");
                first = false;
            } else {
                *out++ = h;
            }
            while (*s && *s != '\n') s++;
            if (*s) s++;
        }
        if (first) die("empty prefilter entry");
        return out;
    });
}

// first index of every key's run in an array ordered by key: start[i] .. start[i + 1] are the records of keys[i]
template <class T, class KeyOf>
std::vector<uint64_t> runStarts(const std::vector<uint32_t> &keys, const T *recs, uint64_t n, KeyOf keyOf) {
    std::vector<uint64_t> start(keys.size() + 1);
    const int nT = mmdb::hostThreads();
    // keys and records are both ascending: every thread finds the first record of its share of the keys by one binary search
    // and walks on from there (a binary search per key is 20 cache misses per key)
    const size_t nk = keys.size();
#pragma omp parallel for num_threads(nT) schedule(static, 1)
    for (int t = 0; t < nT; t++) {
        const size_t a = nk * (size_t) t / (size_t) nT, b = nk * ((size_t) t + 1) / (size_t) nT;
        if (a >= b) continue;
        uint64_t lo = 0, hi = n;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (keyOf(recs[mid]) < keys[a]) lo = mid + 1; else hi = mid; }
        uint64_t pos = lo;
        for (size_t i = a; i < b; i++) {
            const uint32_t k = keys[i];
            while (pos < n && keyOf(recs[pos]) < k) pos++;
            start[i] = pos;
        }
    }
    start[nk] = n;
    return start;
}

// prefilter DB: every key gets "key\t0\t0\n" followed by its hit lines (kmermatcher.cpp:809-924, :705-724;
// QueryMatcher::prefilterHitToBuffer, QueryMatcher.h:114-126).  Entries are written in key order.
void writePrefDb(const std::string &path, bool nucl, const std::vector<uint32_t> &keys, const pg_hit *hits, uint64_t nHits) {
    std::string err;
    mmdb::Writer w;
    if (!w.open(path, nucl ? mmdb::DBTYPE_PREFILTER_REV_RES : mmdb::DBTYPE_PREFILTER_RES, err, true)) die(err);     // a data file per thread
    const std::vector<uint64_t> start = runStarts(keys, hits, nHits, [](const pg_hit &h) { return h.rep; });
    w.writeAll(keys.size(), [&](size_t i) { return keys[i]; }, [&](size_t i, std::string &buf) {
        char line[64];
        const uint32_t key = keys[i];
        char *b = putU(line, key); memcpy(b, "\t0\t0\n", 5); buf.append(line, (size_t) (b + 5 - line));
        for (uint64_t h = start[i]; h < start[i + 1] && hits[h].rep == key; h++) {
            b = putU(line, hits[h].target); *b++ = '\t';
            b = putI(b, hits[h].score); *b++ = '\t';
            b = putI(b, (short) hits[h].diag); *b++ = '\n';
            buf.append(line, (size_t) (b - line));
        }
    });
    if (!w.close()) die("write error");
}

// "%.3E" of an E-value.  The value depends only on (raw score, query length) and the DB size, so a DB holds few distinct
// ones: a small per-thread direct-mapped cache in front of sprintf (the exact glibc rounding stays the only formatter).
struct EvalueText {
    struct Slot { uint64_t bits; char text[14]; unsigned char len; unsigned char used; };
    std::vector<Slot> slots;
    EvalueText() : slots(8192) { for (auto &s : slots) s.used = 0; }
    inline char *put(char *b, double v) {
        uint64_t bits;
        memcpy(&bits, &v, 8);
        Slot &s = slots[(size_t) ((bits * 0x9E3779B97F4A7C15ull) >> 51)];
        if (!s.used || s.bits != bits) {
            const int n = snprintf(s.text, sizeof(s.text), "%.3E", v);
            if (n <= 0 || n >= (int) sizeof(s.text)) return b + sprintf(b, "%.3E", v);
            s.bits = bits; s.len = (unsigned char) n; s.used = 1;
        }
        memcpy(b, s.text, s.len);
        return b + s.len;
    }
};

// alignment DB: Matcher::resultToBuffer (Matcher.cpp:323-370), 10 columns per line
void writeAlnDb(const std::string &path, const std::vector<uint32_t> &keys, const pg_aln *alns, uint64_t nAlns) {
    std::string err;
    mmdb::Writer w;
    if (!w.open(path, mmdb::DBTYPE_ALIGNMENT_RES, err, true)) die(err);     // a data file per thread, as rescorediagonal's DBWriter leaves them
    const std::vector<uint64_t> start = runStarts(keys, alns, nAlns, [](const pg_aln &a) { return a.query; });
    w.writeAll(keys.size(), [&](size_t i) { return keys[i]; }, [&](size_t i, std::string &buf) {
        static thread_local EvalueText evText;
        char line[256];
        const uint32_t key = keys[i];
        for (uint64_t a = start[i]; a < start[i + 1] && alns[a].query == key; a++) {
            const pg_aln &r = alns[a];
            char *b = putU(line, r.target); *b++ = '\t';
            b = putI(b, r.bits); *b++ = '\t';
            b = putSeqId(b, r.seq_id); *b++ = '\t';
            b = evText.put(b, r.evalue); *b++ = '\t';
            b = putI(b, r.q_start); *b++ = '\t'; b = putI(b, r.q_end); *b++ = '\t'; b = putI(b, r.q_len); *b++ = '\t';
            b = putI(b, r.db_start); *b++ = '\t'; b = putI(b, r.db_end); *b++ = '\t'; b = putI(b, r.db_len); *b++ = '\n';
            buf.append(line, (size_t) (b - line));
        }
    });
    if (!w.close()) die("write error");
}

// sequence DB from host arrays (entry i = data[offs[i] .. offs[i] + lens[i]), its last byte the '\0')
void writeSeqArrays(const std::string &path, int dbtype, const char *data, uint64_t bytes, const uint64_t *offs, const uint32_t *lens,
                    const uint32_t *keys, size_t n) {
    mmdb::Writer w;
    std::string err;
    if (!w.open(path, dbtype, err)) die(err);
    // the buffer is the data file already when the entries lie back to back and end with their '\0' (every DB the kernels build)
    unsigned long long gaps = 0;
    const int nT = mmdb::hostThreads();
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : gaps)
    for (size_t i = 0; i < n; i++) {
        const uint64_t next = i + 1 < n ? offs[i + 1] : bytes;
        if (lens[i] == 0 || offs[i] + lens[i] != next || data[offs[i] + lens[i] - 1] != '\0') gaps++;
    }
    if (n && gaps == 0 && offs[0] == 0) w.writeContiguous(data, bytes, n, keys, offs, lens);
    else w.writeAll(n, [&](size_t i) { return keys[i]; }, [&](size_t i, std::string &buf) { buf.append(data + offs[i], lens[i] - 1); });
    if (!w.close()) die("write error");
}

void writeSeqDb(pg_seqdb *out, const std::string &path, int dbtype) {
    char *data; uint64_t bytes, *offs, n; uint32_t *lens, *keys;
    if (pg_seqdb_download(gpu(), out, &data, &bytes, &offs, &lens, &keys, &n) != 0) die(pg_last_error());
    writeSeqArrays(path, dbtype, data, bytes, offs, lens, keys, (size_t) n);
    pg_free_host(data); pg_free_host(offs); pg_free_host(lens); pg_free_host(keys);
}

const std::set<std::string> FS_FLAGS = {"--threads", "-v", "--compressed"};
const std::set<std::string> ORF_FLAGS = {"--min-length", "--max-length", "--max-gaps", "--contig-start-mode", "--contig-end-mode", "--orf-start-mode",
    "--forward-frames", "--reverse-frames", "--translation-table", "--translate", "--use-all-table-starts", "--id-offset", "--create-lookup",
    "--threads", "--compressed", "-v"};
const std::set<std::string> TN_FLAGS = {"--translation-table", "--add-orf-stop", "-v", "--compressed", "--threads"};

unsigned frameMask(const std::string &s) {      // Orf::getFrames (Orf.h:17-35)
    unsigned m = 0;
    size_t p = 0;
    while (p <= s.size()) {
        size_t c = s.find(',', p);
        if (c == std::string::npos) c = s.size();
        const std::string t = s.substr(p, c - p);
        if (t == "1") m |= 1u; else if (t == "2") m |= 2u; else if (t == "3") m |= 4u;
        p = c + 1;
    }
    return m;
}
const std::set<std::string> CC_FLAGS = {"--max-seq-len", "--chop-cycle", "--threads", "-v", "--compressed"};

void checkKmFlags(const Flags &f) {
    requireValue(f, "--mask", "0", "tantan masking is not on the assemble path");
    requireValue(f, "--mask-lower-case", "0", "not on the assemble path");
    requireValue(f, "--spaced-kmer-mode", "0", "spaced k-mers are not on the assemble path");
    requireValue(f, "--adjust-kmer-len", "0", "Markov k-mer length adjustment is not on the assemble path");
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    checkSubMat(f);
}
pg_km_params kmParams(const Flags &f, bool nucl) {
    pg_km_params p;
    p.kmer_size = geti(f, "-k", nucl ? 22 : 14);
    p.alph_size = atoi(multi(get(f, "--alph-size", nucl ? "5" : "13"), nucl).c_str());
    p.kmers_per_seq = geti(f, "--kmer-per-seq", 60);
    p.kmers_per_seq_scale = (float) strtod(multi(get(f, "--kmer-per-seq-scale", nucl ? "0.1" : "0.0"), nucl).c_str(), nullptr);
    p.hash_shift = geti(f, "--hash-shift", 67);
    p.include_only_extendable = geti(f, "--include-only-extendable", 0);
    p.ignore_multi_kmer = geti(f, "--ignore-multi-kmer", 1);
    p.cov_mode = geti(f, "--cov-mode", 0);
    p.cov_thr = (float) getd(f, "-c", 0.0);
    p.hash_start = 0; p.hash_end = 65535;
    return p;
}
// --split-memory-limit (Parameters.cpp, ByteParser: plain bytes or K / M / G / T suffix; 0 = all of the device memory)
uint64_t splitMemoryLimit(const Flags &f) {
    const std::string v = get(f, "--split-memory-limit", "0");
    char *e = nullptr;
    const double x = strtod(v.c_str(), &e);
    if (e == v.c_str() || x < 0) die("cannot parse --split-memory-limit " + v);
    double mul = 1.0;
    switch (*e) {
        case 'k': case 'K': mul = 1024.0; break;
        case 'm': case 'M': mul = 1024.0 * 1024.0; break;
        case 'g': case 'G': mul = 1024.0 * 1024.0 * 1024.0; break;
        case 't': case 'T': mul = 1024.0 * 1024.0 * 1024.0 * 1024.0; break;
        case 'b': case 'B': case '\0': break;
        default: die("cannot parse --split-memory-limit " + v);
    }
    return (uint64_t) (x * mul);
}
void checkRsFlags(const Flags &f) {
    requireValue(f, "--rescore-mode", "3", "the assemble workflows use END_TO_END only");
    requireValue(f, "--wrapped-scoring", "0", "not on the assemble path");
    requireValue(f, "--filter-hits", "0", "not on the assemble path");
    requireValue(f, "-a", "0", "backtraces are not produced by ungapped rescoring on the assemble path");
    requireValue(f, "--sort-results", "0", "the assemble workflows keep prefilter order");
    requireValue(f, "--add-self-matches", "0", "query DB == target DB already keeps the self match");
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    checkSubMat(f);
}
pg_rs_params rsParams(const Flags &f) {
    pg_rs_params p;
    p.rescore_mode = 3;
    p.seq_id_thr = (float) getd(f, "--min-seq-id", 0.0);
    p.eval_thr = getd(f, "-e", 0.001);
    p.cov_mode = geti(f, "--cov-mode", 0);
    p.cov_thr = (float) getd(f, "-c", 0.0);
    p.aln_len_thr = geti(f, "--min-aln-len", 0);
    p.seq_id_mode = geti(f, "--seq-id-mode", 0);
    return p;
}
pg_ex_params exParams(const Flags &f, bool nuclCommand) {
    pg_ex_params p;
    p.seq_id_thr = (float) getd(f, "--min-seq-id", nuclCommand ? 0.99 : 0.9);
    p.max_seq_len = geti(f, "--max-seq-len", nuclCommand ? 200000 : 65535);
    p.keep_target = geti(f, "--keep-target", 1);
    p.rescore_mode = geti(f, "--rescore-mode", 3);
    return p;
}

int extendCommand(int argc, const char **argv, bool nuclCommand) {
    Timer timer;
    Phases ph;
    const Flags f = parseFlags(argc, argv, 3, EX_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    checkSubMat(f);
    std::string err;
    mmdb::Reader seq, aln;
    if (!seq.open(f.positional[0], err) || !aln.open(f.positional[1], err, false)) die(err);
    const bool nucl = seq.dbtype == mmdb::DBTYPE_NUCLEOTIDES;
    (void) nuclCommand;   // like the reference, the comparator follows the command, the letters follow the DB type
    const pg_ex_params p = exParams(f, nuclCommand);
    if (nuclCommand != nucl) die("sequence DB type does not match the command (assembleresults = amino acids, nuclassembleresults = nucleotides)");
    ph.lap("open + index parse");
    const PodArray<pg_aln> alns = parseAlnDb(aln, false);
    ph.lap("parse alignments");
    gpu();
    ph.lap("CUDA init");
    pg_seqdb *db = uploadSeqDb(seq), *out = nullptr;
    ph.lap("upload");
    if (pg_extend(gpu(), db, alns.data(), alns.size(), &p, &out, nullptr) != 0) die(pg_last_error());
    ph.lap("kernels");
    writeSeqDb(out, f.positional[2], seq.dbtype);
    ph.lap("download + write");
    pg_seqdb_free(gpu(), out); pg_seqdb_free(gpu(), db);
    printf("\nDone.\n");
    ph.report();
    timer.report();
    return EXIT_SUCCESS;
}

}  // namespace

int kmermatcher(int argc, const char **argv) {
    Timer timer;
    Phases ph;
    const Flags f = parseFlags(argc, argv, 2, KM_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    checkKmFlags(f);
    std::string err;
    mmdb::Reader seq;
    if (!seq.open(f.positional[0], err)) die(err);
    const bool nucl = seq.dbtype == mmdb::DBTYPE_NUCLEOTIDES;
    const pg_km_params p = kmParams(f, nucl);
    ph.lap("open + index parse");
    if (pg_set_split_memory_limit(gpu(), splitMemoryLimit(f)) != 0) die(pg_last_error());
    ph.lap("CUDA init");
    pg_seqdb *db = uploadSeqDb(seq);
    ph.lap("upload");
    pg_hit *hits = nullptr; uint64_t nHits = 0;
    if (pg_kmermatch(gpu(), db, &p, &hits, &nHits) != 0) die(pg_last_error());
    ph.lap("kernels + download");
    writePrefDb(f.positional[1], nucl, seq.keys, hits, nHits);
    ph.lap("format + write");
    pg_free_host(hits);
    pg_seqdb_free(gpu(), db);
    ph.report();
    timer.report();
    return EXIT_SUCCESS;
}

int rescorediagonal(int argc, const char **argv) {
    Timer timer;
    Phases ph;
    const Flags f = parseFlags(argc, argv, 4, RS_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    checkRsFlags(f);
    if (f.positional[0] != f.positional[1]) die("rescorediagonal on the GPU path requires query DB == target DB (as in assemble.sh / nuclassemble.sh)");
    std::string err;
    mmdb::Reader seq, pref;
    if (!seq.open(f.positional[0], err) || !pref.open(f.positional[2], err, false)) die(err);
    const bool nucl = seq.dbtype == mmdb::DBTYPE_NUCLEOTIDES;
    if (pref.dbtype != (nucl ? mmdb::DBTYPE_PREFILTER_REV_RES : mmdb::DBTYPE_PREFILTER_RES)) die("prefilter DB type does not match the sequence DB");
    if (pref.size() != seq.size() || !std::equal(pref.keys.begin(), pref.keys.end(), seq.keys.begin()))
        die("prefilter DB must hold one entry per sequence (kmermatcher output)");
    const pg_rs_params p = rsParams(f);
    ph.lap("open + index parse");
    const PodArray<pg_hit> hits = parsePrefDb(pref);
    ph.lap("parse prefilter hits");
    gpu();
    ph.lap("CUDA init");
    pg_seqdb *db = uploadSeqDb(seq);
    ph.lap("upload");
    pg_aln *alns = nullptr; uint64_t nAlns = 0;
    if (pg_rescore(gpu(), db, hits.data(), hits.size(), &p, &alns, &nAlns) != 0) die(pg_last_error());
    ph.lap("kernels + download");
    writeAlnDb(f.positional[3], seq.keys, alns, nAlns);
    ph.lap("format + write");
    pg_free_host(alns);
    pg_seqdb_free(gpu(), db);
    ph.report();
    timer.report();
    return EXIT_SUCCESS;
}

// findassemblystart <i:sequenceDB> <i:alnResult> <o:sequenceDB>  (src/assembler/findassemblystart.cpp:35-176)
int findassemblystart(int argc, const char **argv) {
    Timer timer;
    const Flags f = parseFlags(argc, argv, 3, FS_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    std::string err;
    mmdb::Reader seq, aln;
    if (!seq.open(f.positional[0], err) || !aln.open(f.positional[1], err, false)) die(err);
    if (seq.dbtype != mmdb::DBTYPE_AMINO_ACIDS) die("findassemblystart expects an amino-acid sequence DB");
    const PodArray<pg_aln> alns = parseAlnDb(aln, false);
    pg_seqdb *db = uploadSeqDb(seq), *out = nullptr;
    if (pg_findassemblystart(gpu(), db, alns.data(), alns.size(), &out, nullptr) != 0) die(pg_last_error());
    writeSeqDb(out, f.positional[2], mmdb::DBTYPE_AMINO_ACIDS);
    pg_seqdb_free(gpu(), out); pg_seqdb_free(gpu(), db);
    timer.report();
    return EXIT_SUCCESS;
}

// extractorfs <i:sequenceDB> <o:sequenceDB>  (lib/mmseqs/src/util/extractorfs.cpp:20-159): the ORF fragments (keys 0..n-1
// in (read, emission) order, as DBWriter::createRenumberedDB leaves them) and their header DB <o>_h
int extractorfs(int argc, const char **argv) {
    Timer timer;
    const Flags f = parseFlags(argc, argv, 2, ORF_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    requireValue(f, "--create-lookup", "0", "lookup files are not written by the GPU path");
    requireValue(f, "--id-offset", "0", "not used by the assemble workflow");
    // the reference's --translate 1 writes the bare translation; the fused mode of pg_extractorfs is extractorfs +
    // translatenucs --add-orf-stop 1 ('*' framing), which is a different output: run translatenucs as the workflow does
    requireValue(f, "--translate", "0", "use translatenucs on the extracted ORFs, as data/assemble.sh does");
    std::string err;
    mmdb::Reader seq;
    if (!seq.open(f.positional[0], err)) die(err);
    if (seq.dbtype != mmdb::DBTYPE_NUCLEOTIDES) die("extractorfs expects a nucleotide sequence DB");
    pg_orf_params p;
    p.min_length = geti(f, "--min-length", 30);
    p.max_length = geti(f, "--max-length", 32734);
    p.max_gaps = geti(f, "--max-gaps", 2147483647);
    p.contig_start_mode = geti(f, "--contig-start-mode", 2);
    p.contig_end_mode = geti(f, "--contig-end-mode", 2);
    p.orf_start_mode = geti(f, "--orf-start-mode", 1);
    p.forward_frames = frameMask(get(f, "--forward-frames", "1,2,3"));
    p.reverse_frames = frameMask(get(f, "--reverse-frames", "1,2,3"));
    p.translation_table = geti(f, "--translation-table", 1);
    p.use_all_table_starts = geti(f, "--use-all-table-starts", 0);
    const int translate = geti(f, "--translate", 0);
    pg_seqdb *db = uploadSeqDb(seq), *out = nullptr;
    uint32_t *info = nullptr;
    if (pg_extractorfs(gpu(), db, &p, translate, &out, &info) != 0) die(pg_last_error());
    const uint64_t n = pg_seqdb_size(out);
    writeSeqDb(out, f.positional[1], translate ? mmdb::DBTYPE_AMINO_ACIDS : mmdb::DBTYPE_NUCLEOTIDES);
    // header DB: Orf::writeOrfHeader (Orf.cpp:445-462): "readKey \t from (+|-) len [\t flags] \n"
    mmdb::Writer hw;
    if (!hw.open(f.positional[1] + "_h", 12 /* DBTYPE_GENERIC_DB */, err)) die(err);
    char line[96];
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t key = info[4 * i], from = info[4 * i + 1], to = info[4 * i + 2], fl = info[4 * i + 3];
        char *b = putU(line, key); *b++ = '\t';
        b = putU(b, from); *b++ = (from < to) ? '+' : '-';
        b = putU(b, from < to ? to - from : from - to);
        if (fl) { *b++ = '\t'; b = putU(b, fl); }
        *b++ = '\n';
        hw.write((uint32_t) i, line, (size_t) (b - line));
    }
    if (!hw.close()) die("write error");
    linkIfExists(f.positional[0] + ".source", f.positional[1] + ".source");
    pg_free_host(info);
    pg_seqdb_free(gpu(), out); pg_seqdb_free(gpu(), db);
    timer.report();
    return EXIT_SUCCESS;
}

// translatenucs <i:sequenceDB> <o:sequenceDB>  (lib/mmseqs/src/util/translatenucs.cpp:14-128)
int translatenucs(int argc, const char **argv) {
    Timer timer;
    const Flags f = parseFlags(argc, argv, 2, TN_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    std::string err;
    mmdb::Reader seq;
    if (!seq.open(f.positional[0], err)) die(err);
    if (seq.dbtype != mmdb::DBTYPE_NUCLEOTIDES) die("translatenucs expects a nucleotide sequence DB");
    const bool addOrfStop = geti(f, "--add-orf-stop", 0) != 0;
    std::vector<uint8_t> flags;
    if (addOrfStop) {
        // Orf::parseOrfHeader (Orf.cpp:333-443) of the header entry with the same key: third column = incomplete start | end << 1
        mmdb::Reader hdr;
        if (!hdr.open(f.positional[0] + "_h", err)) die(err);
        flags.resize(seq.size());
        for (size_t i = 0; i < seq.size(); i++) {
            const auto it = std::lower_bound(hdr.keys.begin(), hdr.keys.end(), seq.keys[i]);
            if (it == hdr.keys.end() || *it != seq.keys[i]) die("translatenucs --add-orf-stop 1: no ORF header for key " + std::to_string(seq.keys[i]));
            const char *s = hdr.entry((size_t) (it - hdr.keys.begin()));
            int col = 0; unsigned long complete = 0;
            const char *q = s;
            while (*q && *q != '\n') {
                while (*q == ' ' || *q == '\t') q++;
                if (!*q || *q == '\n') break;
                if (col == 2) complete = strtoul(q, nullptr, 10);
                col++;
                while (*q && *q != ' ' && *q != '\t' && *q != '\n') q++;
            }
            if (col < 2) die("translatenucs --add-orf-stop 1: header of key " + std::to_string(seq.keys[i]) + " is not an ORF header");
            flags[i] = (uint8_t) (((complete & 1) ? 0 : 1) | ((complete & 2) ? 0 : 2));
        }
    }
    pg_seqdb *db = uploadSeqDb(seq), *out = nullptr;
    if (pg_translatenucs(gpu(), db, addOrfStop ? flags.data() : nullptr, geti(f, "--translation-table", 1), &out) != 0) die(pg_last_error());
    writeSeqDb(out, f.positional[1], mmdb::DBTYPE_AMINO_ACIDS);
    // DBReader::softlinkDb(db1, db2, SEQUENCE_ANCILLARY): header DB, lookup and source travel with the translated DB
    for (const char *ext : {"_h", "_h.index", "_h.dbtype", ".lookup", ".source"}) linkIfExists(f.positional[0] + ext, f.positional[1] + ext);
    pg_seqdb_free(gpu(), out); pg_seqdb_free(gpu(), db);
    timer.report();
    return EXIT_SUCCESS;
}

// cyclecheck <i:sequenceDB> <o:sequenceDBcycle>  (src/assembler/cyclecheck.cpp:31-274): the output holds only the
// sequences reported as circular, cut at the split diagonal with --chop-cycle 1
int cyclecheck(int argc, const char **argv) {
    Timer timer;
    const Flags f = parseFlags(argc, argv, 2, CC_FLAGS);
    applyThreads(f);
    gpuWarmUp();
    requireValue(f, "--compressed", "0", "uncompressed DBs only");
    std::string err;
    mmdb::Reader seq;
    if (!seq.open(f.positional[0], err)) die(err);
    if (seq.dbtype != mmdb::DBTYPE_NUCLEOTIDES) die("Module cyclecheck only supports nucleotide input database");
    const int maxSeqLen = geti(f, "--max-seq-len", 65535);
    const bool chop = geti(f, "--chop-cycle", 0) != 0;
    pg_seqdb *db = uploadSeqDb(seq);
    uint32_t *split = nullptr;
    if (pg_cyclecheck(gpu(), db, maxSeqLen, &split) != 0) die(pg_last_error());
    mmdb::Writer w;
    if (!w.open(f.positional[1], mmdb::DBTYPE_NUCLEOTIDES, err)) die(err);
    std::string buf;
    for (size_t i = 0; i < seq.size(); i++) {
        if (split[i] == 0) continue;
        if (chop) {                                    // :251-256
            buf.assign(seq.entry(i), split[i]);
            buf.push_back('\n');
            w.write(seq.keys[i], buf.data(), buf.size());
        } else {
            w.write(seq.keys[i], seq.entry(i), seq.lens[i] - 1);
        }
    }
    if (!w.close()) die("write error");
    pg_free_host(split);
    pg_seqdb_free(gpu(), db);
    timer.report();
    return EXIT_SUCCESS;
}

int assembleresults(int argc, const char **argv) { return extendCommand(argc, argv, false); }
int nuclassembleresults(int argc, const char **argv) { return extendCommand(argc, argv, true); }

// assembleiteration <i:sequenceDB> <o:prefDB> <o:alnDB> <o:sequenceDB> [flags of the three steps]
// One whole iteration of data/assemble.sh:88-151 / data/nuclassemble.sh:100-128 in ONE process: the sequence DB is
// uploaded once, kmermatcher -> rescorediagonal -> (nucl)assembleresults stay in HBM (pg_assemble_iteration), and
// pref_N / aln_N / assembly_N are written as the three separate commands would write them.  Saves two process starts,
// two CUDA initialisations, two uploads and the parse of pref_N and aln_N.
int assembleiteration(int argc, const char **argv) {
    Timer timer;
    std::set<std::string> known = KM_FLAGS;
    known.insert(RS_FLAGS.begin(), RS_FLAGS.end());
    known.insert(EX_FLAGS.begin(), EX_FLAGS.end());
    const Flags f = parseFlags(argc, argv, 4, known);
    applyThreads(f);
    gpuWarmUp();
    checkKmFlags(f);
    checkRsFlags(f);
    std::string err;
    mmdb::Reader seq;
    if (!seq.open(f.positional[0], err)) die(err);
    const bool nucl = seq.dbtype == mmdb::DBTYPE_NUCLEOTIDES;
    const pg_km_params kp = kmParams(f, nucl);
    pg_rs_params rp = rsParams(f);
    const pg_ex_params ep = exParams(f, nucl);
    if (pg_set_split_memory_limit(gpu(), splitMemoryLimit(f)) != 0) die(pg_last_error());
    const auto tOpen = std::chrono::steady_clock::now();
    pg_seqdb *db = uploadSeqDb(seq), *out = nullptr;
    pg_hit *hits = nullptr; pg_aln *alns = nullptr; uint64_t nHits = 0, nAlns = 0;
    if (pg_assemble_iteration(gpu(), db, &kp, &rp, &ep, &out, &hits, &nHits, &alns, &nAlns) != 0) die(pg_last_error());
    const auto tGpu = std::chrono::steady_clock::now();
    writePrefDb(f.positional[1], nucl, seq.keys, hits, nHits);
    writeAlnDb(f.positional[2], seq.keys, alns, nAlns);
    writeSeqDb(out, f.positional[3], seq.dbtype);
    pg_free_host(hits); pg_free_host(alns);
    pg_seqdb_free(gpu(), out); pg_seqdb_free(gpu(), db);
    const auto tEnd = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (long long) std::chrono::duration_cast<std::chrono::milliseconds>(b - a).count(); };
    printf("open + index parse %lld ms, upload + GPU iteration + result download %lld ms, format + write %lld ms (%d host threads)\n",
           ms(timer.t0, tOpen), ms(tOpen, tGpu), ms(tGpu, tEnd), mmdb::hostThreads());
    timer.report();
    return EXIT_SUCCESS;
}

// dbdiff <DB a> <DB b> [--mode exact|aln|pref]: logical comparison key -> entry bytes of two MMseqs DBs (physical
// offsets and index order are not a contract).  Prints one JSON line; exit status 0 iff nothing
// mismatches.  --mode aln: an entry whose lines differ only in the E-value column by at most one unit of its last
// printed digit counts as `tolerated` (north-star tolerance 1e-6 on the value, "%.3E" on disk); --mode pref: entries
// whose lines differ only in the sign of the score (nucleotide strand flag, DESIGN.md section 4 hazard 6) are `tolerated`.
int dbdiff(int argc, const char **argv) {
    const Flags f = parseFlags(argc, argv, 2, {"--mode", "--threads"});
    applyThreads(f);
    const std::string mode = get(f, "--mode", "exact");
    if (mode != "exact" && mode != "aln" && mode != "pref") die("dbdiff --mode must be exact, aln or pref");
    std::string err;
    mmdb::Reader a, b;
    if (!a.open(f.positional[0], err, false) || !b.open(f.positional[1], err, false)) die(err);
    // keys of b looked up in a (both ascending)
    unsigned long long missing = 0, mismatching = 0, tolerated = 0, toleratedLines = 0, same = 0;
    long long firstBad = -1;
    const size_t nb = b.size(), na = a.size();
    const int nT = mmdb::hostThreads();
    auto splitCols = [](const char *s, const char *e, const char **col, int maxCols) {
        int c = 0;
        const char *p = s;
        while (c < maxCols) { col[c++] = p; while (p < e && *p != '\t') p++; if (p >= e) break; p++; }
        col[c] = e + 1;
        return c;
    };
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : missing, mismatching, tolerated, toleratedLines, same)
    for (size_t i = 0; i < nb; i++) {
        const auto it = std::lower_bound(a.keys.begin(), a.keys.end(), b.keys[i]);
        if (it == a.keys.end() || *it != b.keys[i]) { missing++; continue; }
        const size_t j = (size_t) (it - a.keys.begin());
        const char *ea = a.entry(j), *eb = b.entry(i);
        const size_t la = a.lens[j], lb = b.lens[i];
        if (la == lb && memcmp(ea, eb, la) == 0) { same++; continue; }
        bool ok = mode != "exact";
        unsigned long long lines = 0;
        if (ok) {
            const char *pa = ea, *pb = eb, *enda = ea + (la ? la - 1 : 0), *endb = eb + (lb ? lb - 1 : 0);
            while (ok && pa < enda && pb < endb) {
                const char *la_e = (const char *) memchr(pa, '\n', (size_t) (enda - pa)); if (!la_e) la_e = enda;
                const char *lb_e = (const char *) memchr(pb, '\n', (size_t) (endb - pb)); if (!lb_e) lb_e = endb;
                if ((la_e - pa) != (lb_e - pb) || memcmp(pa, pb, (size_t) (la_e - pa)) != 0) {
                    const char *ca[12], *cb[12];
                    const int nca = splitCols(pa, la_e, ca, 10), ncb = splitCols(pb, lb_e, cb, 10);
                    if (nca != ncb) { ok = false; break; }
                    for (int c = 0; c < nca && ok; c++) {
                        const size_t wa = (size_t) (ca[c + 1] - ca[c] - 1), wb = (size_t) (cb[c + 1] - cb[c] - 1);
                        if (wa == wb && memcmp(ca[c], cb[c], wa) == 0) continue;
                        if (mode == "aln" && c == 3) {
                            const double x = strtod(ca[c], nullptr), y = strtod(cb[c], nullptr);
                            ok = std::fabs(x - y) <= 1.0005e-3 * std::fabs(y);
                        } else if (mode == "pref" && c == 1) {
                            ok = strtol(ca[c], nullptr, 10) == -strtol(cb[c], nullptr, 10);
                        } else ok = false;
                    }
                    lines++;
                }
                pa = la_e < enda ? la_e + 1 : enda;
                pb = lb_e < endb ? lb_e + 1 : endb;
            }
            if (ok && (pa < enda || pb < endb)) ok = false;      // different number of lines
        }
        if (ok) { tolerated++; toleratedLines += lines; }
        else {
            mismatching++;
#pragma omp critical
            if (firstBad < 0 || (long long) b.keys[i] < firstBad) firstBad = (long long) b.keys[i];
        }
    }
    unsigned long long onlyA = 0;
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : onlyA)
    for (size_t j = 0; j < na; j++) {
        const auto it = std::lower_bound(b.keys.begin(), b.keys.end(), a.keys[j]);
        if (it == b.keys.end() || *it != a.keys[j]) onlyA++;
    }
    printf("{\"entries_a\": %zu, \"entries_b\": %zu, \"identical\": %llu, \"tolerated\": %llu, \"tolerated_lines\": %llu, \"mismatching\": %llu, "
           "\"only_in_b\": %llu, \"only_in_a\": %llu, \"dbtype_equal\": %s, \"first_mismatching_key\": %lld, \"mode\": \"%s\"}\n",
           na, nb, same, tolerated, toleratedLines, mismatching, missing, onlyA, a.dbtype == b.dbtype ? "true" : "false", firstBad, mode.c_str());
    return (mismatching == 0 && missing == 0 && onlyA == 0 && a.dbtype == b.dbtype) ? EXIT_SUCCESS : EXIT_FAILURE;
}

// iotest <pref|aln|seq> <i:DB> <o:DB>: the text layer alone -- parse a prefilter / alignment DB into the records the kernels
// consume and write them back (no GPU involved).  The round trip must reproduce the input (dbdiff); the printed phase
// times are the host-side cost of a drop-in step.
int iotest(int argc, const char **argv) {
    Timer timer;
    Phases ph;
    const Flags f = parseFlags(argc, argv, 3, {"--threads"});
    applyThreads(f);
    std::string err;
    mmdb::Reader in;
    if (!in.open(f.positional[1], err, false)) die(err);
    ph.lap("open + index parse");
    if (f.positional[0] == "pref") {
        const PodArray<pg_hit> hits = parsePrefDb(in);
        ph.lap("parse prefilter hits");
        writePrefDb(f.positional[2], in.dbtype == mmdb::DBTYPE_PREFILTER_REV_RES, in.keys, hits.data(), hits.size());
    } else if (f.positional[0] == "aln") {
        const PodArray<pg_aln> alns = parseAlnDb(in, true, true);      // every identity column cross-checked against strtof
        ph.lap("parse alignments");
        {
            // the parse the extension commands use (E-value column skipped) must agree in every other field
            const PodArray<pg_aln> lean = parseAlnDb(in, false);
            if (lean.size() != alns.size()) die("iotest: the two alignment parsers disagree on the number of lines");
            for (size_t i = 0; i < alns.size(); i++) {
                pg_aln a = alns.data()[i];
                a.evalue = 0.0;
                if (memcmp(&a, lean.data() + i, sizeof(pg_aln)) != 0 && !(a.query == lean.data()[i].query && a.target == lean.data()[i].target && a.bits == lean.data()[i].bits &&
                        a.seq_id == lean.data()[i].seq_id && a.q_start == lean.data()[i].q_start && a.q_end == lean.data()[i].q_end && a.q_len == lean.data()[i].q_len &&
                        a.db_start == lean.data()[i].db_start && a.db_end == lean.data()[i].db_end && a.db_len == lean.data()[i].db_len))
                    die("iotest: the two alignment parsers disagree");
            }
            ph.lap("parse without E-values + compare");
        }
        writeAlnDb(f.positional[2], in.keys, alns.data(), alns.size());
    } else if (f.positional[0] == "seq") {
        // a sequence DB through the writer of the GPU commands' results (contiguous fast path or entry by entry)
        mmdb::Reader seq;
        if (!seq.open(f.positional[1], err)) die(err);
        ph.lap("open sequence DB");
        writeSeqArrays(f.positional[2], seq.dbtype, seq.data(), seq.dataBytes(), seq.offsets.data(), seq.lens.data(), seq.keys.data(), seq.size());
    } else die("iotest: first argument must be pref, aln or seq");
    ph.lap("format + write");
    ph.report();
    timer.report();
    return EXIT_SUCCESS;
}
