// mmdb.h -- minimal MMseqs2 DB reader/writer for the drop-in commands (own implementation of the
// on-disk format; reference: lib/mmseqs/src/commons/DBReader.cpp:173-253,770-831, DBWriter.cpp:193-252,522-614).
//   X | X.0..X.k  entry bytes, each entry ends with '\0'
//   X.index       "key \t offset \t length \n" (length includes the '\0'; not guaranteed key-sorted)
//   X.dbtype      4-byte little-endian int
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mmdb {

enum { DBTYPE_AMINO_ACIDS = 0, DBTYPE_NUCLEOTIDES = 1, DBTYPE_ALIGNMENT_RES = 5, DBTYPE_PREFILTER_RES = 7, DBTYPE_PREFILTER_REV_RES = 14 };

struct Reader {
    std::vector<char> data;          // concatenation of the data files
    std::vector<uint32_t> keys;      // ascending
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> lens;
    int dbtype = 0;
    bool open(const std::string &path, std::string &err);
    size_t size() const { return keys.size(); }
    const char *entry(size_t i) const { return data.data() + offsets[i]; }
};

// Writes entries (already in ascending key order) as one data file + index + dbtype.
struct Writer {
    std::string path;
    FILE *fd = nullptr, *fi = nullptr;
    uint64_t offset = 0;
    bool open(const std::string &path, int dbtype, std::string &err);
    void write(uint32_t key, const char *bytes, size_t n);   // appends '\0'
    bool close();
};

}  // namespace mmdb
