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
    // contiguous = false: split data files X.0 .. X.k are mapped one by one (no copy) and only entry(i) may be used --
    // enough for result DBs, which are parsed entry by entry; sequence DBs are uploaded as one buffer and need data().
    bool open(const std::string &path, std::string &err, bool contiguous = true);
    size_t size() const { return keys.size(); }
    const char *data() const { return base; }            // nullptr for a segmented reader
    size_t dataBytes() const { return bytes; }
    bool segmented() const { return !segs.empty(); }
    const char *entry(size_t i) const {
        const uint64_t o = offsets[i];
        if (segs.empty()) return base + o;
        size_t k = 0;
        while (k + 1 < segs.size() && o >= segs[k + 1].start) k++;      // at most one file per writer thread
        return segs[k].p + (o - segs[k].start);
    }

private:
    struct Seg { const char *p; uint64_t start; size_t len; };
    const char *base = nullptr;      // mmap of the single data file, or one anonymous mapping filled from the split files
    size_t bytes = 0;
    void *mapped = nullptr;
    size_t mappedBytes = 0;
    std::vector<Seg> segs;           // split data files mapped one by one
    void unmapAll();
};

// Writes entries (in ascending key order) as data file(s) + index + dbtype.
// splitData = false: one data file X (DBWriter::close(merge = true): what the assembler leaves).
// splitData = true:  one data file per host thread, X.0 .. X.k, index offsets into their concatenation (DBWriter::close without
//   merge: what rescorediagonal leaves) -- buffered writes to ONE file serialise on its inode lock, so a single data file caps the
//   writer at one core's page-cache copy rate however many threads format; with a file per thread the writes run in parallel.
//   Files that received no data are not left behind, and a single non-empty file is named X.
struct Writer {
    std::string path;
    int fd = -1, fi = -1;
    uint64_t offset = 0, indexOffset = 0;
    std::string pendData, pendIndex;     // sequential interface: buffered
    bool failed = false;
    bool split = false;
    std::vector<int> fds;                // split mode: the per-thread data files
    std::vector<uint64_t> fileBytes;
    struct Ent { uint32_t key; uint32_t len; uint32_t file; uint64_t off; };
    std::vector<Ent> ents;               // split mode: the index is written at close(), once the files' final sizes are known
    bool open(const std::string &path, int dbtype, std::string &err, bool splitData = false);
    void write(uint32_t key, const char *bytes, size_t n);   // appends '\0'
    // Entries i = 0 .. n-1 with key keyOf(i): format(i, out) APPENDS the entry's bytes (without the trailing '\0') to out and
    // may be called from any host thread, for any i, in any order; an entry for which `skip(i)` holds is not written.
    void writeAll(size_t n, const std::function<uint32_t(size_t)> &keyOf, const std::function<void(size_t, std::string &)> &format,
                  const std::function<bool(size_t)> &skip = nullptr);
    // The entries already lie back to back, each with its trailing '\0', in `data` (offsets[i] + lens[i] == offsets[i + 1],
    // offsets[0] == 0): the data file is that buffer, written by one thread in large pieces (one writer per inode is the fast
    // way to fill a file through the page cache) while the others format the index.  Single-file mode, nothing written before.
    void writeContiguous(const char *data, uint64_t bytes, size_t n, const uint32_t *keys, const uint64_t *offsets, const uint32_t *lens);
    bool close();
};

}  // namespace mmdb
