// pg_common.cuh -- shared device/host helpers of the B200 hot path (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/plassgpu.h"

namespace pg {

void set_error(const std::string &msg);

#define PG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            pg::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define PG_CHECK(cond, msg)                                                                        \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            pg::set_error(std::string(msg) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define PG_TRY(call)                    \
    do {                                \
        int rc_ = (call);               \
        if (rc_ != 0) return rc_;       \
    } while (0)

// 16-byte record that every sort in the pipeline moves (KmerPosition<short> is 16 B too,
// reference lib/mmseqs/src/linclust/kmermatcher.h:49-54).
//   k-mer record  (sort #1): w0 = kmer (nt: strand flag in bit 63), w1 = id<<32 | seqLen<<16 | pos
//   pair record   (sort #2): w0 = rep<<32 | target, w1 = strand<<17 | qrev... see pg_kmermatch.cu
struct __align__(16) Rec {
    unsigned long long w0;
    unsigned long long w1;
};

constexpr int NUM_SMS = 148;   // B200

// XXH64 of one little-endian u64 with seed (xxhash.h XXH64, len == 8 path); reference call site
// kmermatcher.cpp:33-38 hashUInt64.
__host__ __device__ __forceinline__ unsigned long long xxh64_u64(unsigned long long v, unsigned long long seed) {
    const unsigned long long P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL,
                             P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    unsigned long long h = seed + P5 + 8ULL;
    unsigned long long k = v * P2;
    k = (k << 31) | (k >> 33);
    k *= P1;
    h ^= k;
    h = ((h << 27) | (h >> 37)) * P1 + P4;
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

// cudaFuncSetAttribute is a per-DEVICE setting: call sites remember the devices they have configured, not a process-wide flag
// (contexts on several GPUs in one process).  True the first time the site runs on the calling thread's current device.
inline bool first_use_on_device(std::atomic<unsigned long long> &seen) {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ULL << (d & 63);
    return (seen.fetch_or(bit) & bit) == 0;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

// key -> index in the key-sorted DB (DBReader::getId).  Dense DBs (key == index) take the fast path.
__device__ __forceinline__ unsigned find_id(const unsigned *__restrict__ keys, unsigned n, unsigned key) {
    if (key < n && __ldg(keys + key) == key) return key;
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && __ldg(keys + lo) == key) ? lo : 0xFFFFFFFFu;
}

// Simple device buffer with capacity reuse.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 16 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_error(std::string("cudaMalloc of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e)); return 1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *) p; }
};

}  // namespace pg

// Device view of a sequence DB.
struct pg_seqdb {
    char *data = nullptr;                 // entry bytes (residues + "\n\0")
    unsigned long long *offsets = nullptr;
    unsigned *lens = nullptr;             // entry lengths (seqLen = len - 2)
    unsigned *keys = nullptr;
    uint64_t n = 0;
    uint64_t data_bytes = 0;
    int dbtype = 0;
    unsigned max_seq_len = 0;             // max(len) - 2
    unsigned max_key = 0;
    double residues = 0;                  // getAminoAcidDBSize = sum(len) - 2n
    bool dense_keys = false;
    bool contiguous = false;              // offsets[i] + lens[i] == offsets[i + 1] for every i < n - 1 (entries back to back in index order)
    bool borrowed = false;                // pg_seqdb_adopt: the arrays belong to the caller
    bool downloadPending = false;         // an asynchronous pg_seqdb_download is (or was) in flight on the copy stream
    // pg_seqdb_upload_async: the host -> device copies run on the context's upload stream; evReady marks their end and the
    // per-DB statistics (max length, residues, ...) are computed at first use (pg::db_ready)
    cudaEvent_t evReady = nullptr;
    bool uploadPending = false;
};

namespace pg {
// key -> index with the DB's own shortcuts: a dense DB (key == index everywhere, the usual sequence DB) needs no look at the
// keys array at all -- one 32-byte sector less per random look-up
__device__ __forceinline__ unsigned find_id_db(const pg_seqdb &db, unsigned key) {
    if (db.dense_keys) return key < (unsigned) db.n ? key : 0xFFFFFFFFu;
    return find_id(db.keys, (unsigned) db.n, key);
}
// entry i: first byte and entry length (residues + 2).  Entries that lie back to back give their length as the difference of
// two neighbouring offsets (the same sector three times out of four) instead of a load from a third array.
__device__ __forceinline__ const char *seq_entry(const pg_seqdb &db, unsigned i, unsigned *entryLen) {
    const unsigned long long o = db.offsets[i];
    if (db.contiguous && (unsigned long long) i + 1 < db.n) *entryLen = (unsigned) (db.offsets[i + 1] - o);
    else *entryLen = db.lens[i];
    return db.data + o;
}
}  // namespace pg

