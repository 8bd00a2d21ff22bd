#include "mmdb.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <sys/stat.h>

namespace mmdb {

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

bool Reader::open(const std::string &path, std::string &err) {
    data.clear();
    if (fileExists(path)) {
        if (!slurp(path, data, 0)) { err = "cannot read " + path; return false; }
    } else {
        int i = 0;
        while (fileExists(path + "." + std::to_string(i))) {
            if (!slurp(path + "." + std::to_string(i), data, data.size())) { err = "cannot read split " + path; return false; }
            i++;
        }
        if (i == 0) { err = "database " + path + " not found"; return false; }
    }
    std::vector<char> idx;
    if (!slurp(path + ".index", idx, 0)) { err = "cannot read " + path + ".index"; return false; }
    std::vector<uint32_t> k; std::vector<uint64_t> o; std::vector<uint32_t> l;
    const char *p = idx.data(), *e = idx.data() + idx.size();
    while (p < e) {
        uint64_t v[3] = {0, 0, 0};
        for (int c = 0; c < 3; c++) {
            while (p < e && (*p < '0' || *p > '9')) { if (*p == '\n') break; p++; }
            while (p < e && *p >= '0' && *p <= '9') { v[c] = v[c] * 10 + (uint64_t) (*p - '0'); p++; }
        }
        while (p < e && *p != '\n') p++;
        if (p < e) p++;
        k.push_back((uint32_t) v[0]); o.push_back(v[1]); l.push_back((uint32_t) v[2]);
    }
    std::vector<size_t> order(k.size());
    std::iota(order.begin(), order.end(), 0);
    if (!std::is_sorted(k.begin(), k.end())) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return k[a] < k[b]; });
    keys.resize(k.size()); offsets.resize(k.size()); lens.resize(k.size());
    for (size_t i = 0; i < order.size(); i++) {
        keys[i] = k[order[i]]; offsets[i] = o[order[i]]; lens[i] = l[order[i]];
        if (offsets[i] + lens[i] > data.size()) { err = "index of " + path + " points outside the data file"; return false; }
    }
    FILE *ft = fopen((path + ".dbtype").c_str(), "rb");
    if (!ft) { err = "cannot read " + path + ".dbtype"; return false; }
    int32_t t = 0;
    if (fread(&t, 4, 1, ft) != 1) { fclose(ft); err = "short dbtype file"; return false; }
    fclose(ft);
    if (t & (1 << 31)) { err = path + " is zstd-compressed (--compressed 1): not supported by the GPU commands"; return false; }
    dbtype = t & 0xFFFF;
    return true;
}

bool Writer::open(const std::string &p, int dbtype, std::string &err) {
    path = p;
    fd = fopen(p.c_str(), "wb");
    fi = fopen((p + ".index").c_str(), "wb");
    if (!fd || !fi) { err = "cannot open " + p + " for writing"; return false; }
    setvbuf(fd, nullptr, _IOFBF, 1 << 22);
    setvbuf(fi, nullptr, _IOFBF, 1 << 22);
    FILE *ft = fopen((p + ".dbtype").c_str(), "wb");
    if (!ft) { err = "cannot write dbtype"; return false; }
    const int32_t t = dbtype;
    fwrite(&t, 4, 1, ft);
    fclose(ft);
    offset = 0;
    return true;
}

void Writer::write(uint32_t key, const char *bytes, size_t n) {
    fwrite(bytes, 1, n, fd);
    fputc('\0', fd);
    fprintf(fi, "%u\t%llu\t%llu\n", key, (unsigned long long) offset, (unsigned long long) (n + 1));
    offset += n + 1;
}

bool Writer::close() {
    bool ok = true;
    if (fd) ok &= fclose(fd) == 0;
    if (fi) ok &= fclose(fi) == 0;
    fd = fi = nullptr;
    return ok;
}

}  // namespace mmdb
