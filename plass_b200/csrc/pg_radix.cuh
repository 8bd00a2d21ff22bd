// pg_radix.cuh -- hand-written least-significant-digit radix sort for 16-byte records (sm_100a).
//
// Replaces the reference's SORT_PARALLEL (ips4o / omptl) calls, lib/mmseqs/src/commons/FastSort.h:1-24,
// used at kmermatcher.cpp:408-412 (sort #1) and :427-431 (sort #2).
//
// Design (HBM-bound: every pass is one coalesced read + one coalesced write of the records):
//   * one histogram kernel computes the 256-bin histograms of ALL passes in a single read -- or none at all when the producer of
//     the records counted them on the way out (preHist: the extraction kernels do, for sort #1),
//   * per pass one "onesweep" kernel: a tile of 3072 records is ranked in registers with
//     warp match_any + warp-private counters, reordered through shared memory so that every digit
//     run leaves the SM as contiguous 16-byte stores, and the tile's global offsets come from a
//     decoupled look-back over per-tile status words (single pass, no second read of the data),
//   * stable, so passes compose into a multi-word key sort.
//   * the last pass can leave by-products of the final order it alone knows: bucket bounds + smallest key (sort #1), segment
//     bounds + smallest target per representative (sort #2) -- RadixBounds,
//   * a pass can scatter straight into OTHER GPUs' memory (radix_scatter_peer: the multi-GPU exchanges).
// Round 2 also built the pass as a PERSISTENT kernel whose tiles arrive by bulk asynchronous copy (cp.async.bulk global -> shared,
// completion on an mbarrier; SASS UBLKCP.S.G) and whose digit runs leave as bulk shared -> global copies (UBLKCP.G.S).  Measured
// slower than the register-tile kernel inside the iteration (3.5 vs 2.75 ms per pass): it stays selectable
// (radix_set_mode 1..3 / PLASS_B200_RADIX_MODE) and parity-tested, the default is mode 0.
#pragma once
#include "pg_common.cuh"

namespace pg {

struct DigitPass {
    int word;            // 0 -> Rec::w0, 1 -> Rec::w1
    int shift;           // digit = (w >> shift) & mask
    unsigned mask;       // <= 255 (8-bit digits) or <= 1023 (wide digits: 9 / 10 bits, see plan_add_bits_w)
    int hashed;          // 1: w is first replaced by mix64(w & hashMask) -- partition by hash bits instead of key bits
                         // 2: digit = index of the interval of aux[0..auxN] (ascending) that holds the high half of w0 (owner rank of a representative)
                         // 3: w is replaced by (w >> 32) - hashMask (keys of a rank's contiguous range, rebased to 0)
    unsigned long long hashMask;
    const unsigned *aux; // mode 2: device array of interval bounds, aux[0] = 0
    unsigned auxN;       // mode 2: number of intervals
};

// murmur3 fmix64: the bucket hash of the partial-key partition (any fixed bijective mixer would do)
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

constexpr int RADIX_MAX_PASSES = 12;
constexpr int RADIX_THREADS = 256;
// records per thread of the scatter kernel: 16 (4096-record tile, 2 CTAs/SM), 12 (3 CTAs/SM) or 8 (4 CTAs/SM)
void radix_set_items(int items);

struct RadixPlan {
    DigitPass pass[RADIX_MAX_PASSES];
    int npasses;
};

// Optional by-product of the LAST pass of a hashed-bucket partition (sort #1): the first / one-past-last output index of
// every bucket (bucket = mix64(w0 & hashMask) & bucketMask, all of whose bits the plan sorts by) and the smallest
// w0 & hashMask of the input.  The last pass knows every record's final position, so the separate sweep over the sorted
// records that used to find the bucket boundaries (one more read of every record) is not needed.  start[] must be preset
// to 0xFF.., end[] to 0, *minKey to 0xFF.. by the caller.
// kind 1 (sort #2): the segments are the values of the high half of w0 (the representative); start / end are indexed by it
// (start preset to 0xFF.., end to 0) and minLow[value] (preset to 0xFF..) receives the smallest low half of w0 (the smallest target).
struct RadixBounds {
    unsigned long long *start = nullptr;
    unsigned long long *end = nullptr;
    unsigned long long *minKey = nullptr;
    unsigned long long hashMask = 0;
    unsigned bucketMask = 0;
    int kind = 0;
    unsigned *minLow = nullptr;
};

// Workspace owned by the caller (sized by radix_workspace_bytes); maxDigitBits = the widest digit of the plan (8..10).
size_t radix_workspace_bytes(uint64_t n, int maxDigitBits = 8);

// Sorts n records by the digits of plan (pass[0] least significant).  `a` holds the input, `b` is a
// scratch buffer of the same size; *sorted points to whichever of the two holds the result.
int radix_sort(Rec *a, Rec *b, uint64_t n, const RadixPlan &plan, void *workspace, size_t workspace_bytes,
               cudaStream_t stream, Rec **sorted, uint64_t *launches,
               cudaEvent_t evScatterBegin = nullptr, cudaEvent_t evScatterEnd = nullptr, const RadixBounds *bounds = nullptr,
               const unsigned long long *preHist = nullptr);
// preHist (device, 3 x 256 counters): the histograms of the 8-bit digits of mix64(w0 & hashMask) at shifts 0, 8, 16 over exactly
// these n records, counted by the producer of the records.  If the plan consists of such digits (plan_add_hash_bits from bit 0,
// at most three passes; a narrower last digit is folded) the histogram sweep over the records is skipped.
// true if radix_sort would honour `bounds` (kind 0) for this plan
bool radix_emits_bounds(const RadixPlan &plan);
// true if radix_sort would honour kind-1 bounds (segments by the high half of w0) for this plan
bool radix_emits_segments(const RadixPlan &plan);
// 0: register-tile kernel (round 1), 1: persistent bulk-copy (TMA) kernel, 2 records-per-thread / stage variants; see pg_radix.cu
void radix_set_mode(int mode);
int radix_get_mode();

// Fused partition + exchange (multi-GPU): one pass over `a` whose digit runs are written to dstBase[digit] + rank-within-digit
// x 16 bytes -- dstBase = device array of 256 byte addresses, local or in a peer GPU's memory (CUDA IPC mapping, NVLink stores).
// radix_pass_histogram first (device histogram -> host, bases left in the workspace), then radix_scatter_peer.
int radix_pass_histogram(const Rec *a, uint64_t n, const DigitPass &pass, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                         unsigned long long *hostHist /* [256] */, uint64_t *launches);
int radix_scatter_peer(const Rec *a, uint64_t n, const DigitPass &pass, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                       const unsigned long long *d_dstBase, uint64_t *launches);

// Helper to build a plan over bit ranges: appends 8-bit digits covering bits [lo, hi) of word w.
void plan_add_bits(RadixPlan &plan, int word, int lo, int hi);
// same over bits [lo, hi) of mix64(w0 & hashMask)
void plan_add_hash_bits(RadixPlan &plan, unsigned long long hashMask, int lo, int hi);
// one pass whose digit is the interval index of (w0 >> 32) in bounds[0..n) (device array, ascending, bounds[0] = 0), n <= 256
void plan_add_interval(RadixPlan &plan, const unsigned *deviceBounds, unsigned n);
// 8-bit digits over bits [0, nbits) of (w0 >> 32) - sub
void plan_add_rebased_high_bits(RadixPlan &plan, unsigned long long sub, int nbits);
// Wide digits: the bit range is cut into ceil((hi-lo)/digitBits) digits of (almost) equal width <= digitBits (8..10).
// A 512- or 1024-bin pass costs a little more than a 256-bin pass (status words, shorter store runs) but a 20-bit key
// takes 2 passes instead of 3.
void plan_add_bits_w(RadixPlan &plan, int word, int lo, int hi, int digitBits);
void plan_add_hash_bits_w(RadixPlan &plan, unsigned long long hashMask, int lo, int hi, int digitBits);

}  // namespace pg
