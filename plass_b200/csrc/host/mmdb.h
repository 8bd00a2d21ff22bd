// mmdb.h -- MMseqs2 on-disk DB reader/writer for the drop-in commands (own implementation of the on-disk format;
// reference: lib/mmseqs/src/commons/DBReader.cpp:173-253,770-831 (open / index parse, capped at 4 threads there),
// DBWriter.cpp:193-252,522-614 (per-thread files + merge)).
//   X | X.0..X.k  entry bytes, each entry ends with '\0'
//   X.index       "key \t offset \t length \n" (length includes the '\0'; not guaranteed key-sorted)
//   X.dbtype      4-byte little-endian int
// The text layer is what decides the wall-clock of a drop-in step once the kernels take milliseconds:
// the data file is mmap'ed instead of read, the index is parsed by all host threads, and entries are formatted by all
// host threads into per-chunk buffers that are written with pwrite at their final offsets.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace mmdb {

enum { DBTYPE_AMINO_ACIDS = 0, DBTYPE_NUCLEOTIDES = 1, DBTYPE_ALIGNMENT_RES = 5, DBTYPE_PREFILTER_RES = 7, DBTYPE_GENERIC_DB = 12, DBTYPE_PREFILTER_REV_RES = 14 };

int hostThreads();                   // --threads / MMSEQS_NUM_THREADS / all cores
void setHostThreads(int n);

struct Reader {
    std::vector<uint32_t> keys;      // ascending
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> lens;
    int dbtype = 0;
    Reader() = default;
    Reader(const Reader &) = delete;
    Reader &operator=(const Reader &) = delete;
    ~Reader();
    bool open(const std::string &path, std::string &err);
    size_t size() const { return keys.size(); }
    const char *data() const { return base; }
    size_t dataBytes() const { return bytes; }
    const char *entry(size_t i) const { return base + offsets[i]; }

private:
    const char *base = nullptr;      // mmap of the single data file, or `owned` (split data files X.0 .. X.k)
    size_t bytes = 0;
    void *mapped = nullptr;
    size_t mappedBytes = 0;
    std::vector<char> owned;
};

// Writes entries (in ascending key order) as one data file + index + dbtype.
struct Writer {
    std::string path;
    int fd = -1, fi = -1;
    uint64_t offset = 0, indexOffset = 0;
    std::string pendData, pendIndex;     // sequential interface: buffered
    bool failed = false;
    bool open(const std::string &path, int dbtype, std::string &err);
    void write(uint32_t key, const char *bytes, size_t n);   // appends '\0'
    // Entries i = 0 .. n-1 with key keyOf(i): format(i, out) APPENDS the entry's bytes (without the trailing '\0') to out and
    // may be called from any host thread, for any i, in any order; an entry for which `skip(i)` holds is not written.
    void writeAll(size_t n, const std::function<uint32_t(size_t)> &keyOf, const std::function<void(size_t, std::string &)> &format,
                  const std::function<bool(size_t)> &skip = nullptr);
    bool close();
};

}  // namespace mmdb
