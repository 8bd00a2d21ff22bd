// pg_kmermatch.cu -- GPU kmermatcher: k-mer extraction + bottom-m selection, group/join, best diagonal.
//
// Replaces (reference lib/mmseqs/src/linclust/kmermatcher.cpp):
//   fillKmerPositionArray :77-385   -> classify_kernel + extract_warp_kernel<NMAX> + extract_block_kernel
//   SORT_PARALLEL #1      :408-412  -> pg::radix_sort on the k-mer word
//   assignGroup           :450-559  -> group_kernel (rep = min by (seqLen desc, id, pos), diagonal, filter)
//   SORT_PARALLEL #2      :427-431  -> pg::radix_sort on (rep, target, diagonal)
//   writeKmerMatcherResult:809-924  -> reduce_count_kernel / reduce_emit_kernel (best diagonal per rep,target)
// Bit-exactness notes refer to SURVEY.md Appendix A.1-A.4.
#include "pg_internal.cuh"
#include "pg_scan.cuh"
#include "pg_tables.h"

#include <cstring>
#include <type_traits>

namespace pg {

__constant__ unsigned char c_aa2num[256];   // ASCII -> code of the active k-mer alphabet
struct KmConst {
    int k;
    int nt;               // 1 = nucleotides
    int xCode;            // code of X (12 reduced / 20 full / 4 nt)
    unsigned base;        // aa index base = alphabetSize-1 (kmermatcher.cpp:113)
    unsigned long long seed;
    int kmersPerSeq;
    float scale;
    int ignoreMulti;
    unsigned hashStart, hashEnd;
    int includeOnlyExtendable;
    int covMode;
    float covThr;
    // wide records (KmerPosition<int>, kmermatcher.cpp:797-802: any sequence of >= SHRT_MAX residues): the 16-byte record
    // cannot hold id + seqLen + pos at full width, so it carries the sequence's RANK in (seqLen desc, id asc) order --
    // exactly the part of sort #1's comparator between the k-mer and the position (kmermatcher.h:56-96) -- and a 32-bit
    // position; id and seqLen are looked up from the rank where assignGroup needs them.
    int wide;
    const unsigned *rankOf;   // sequence index -> rank
    const unsigned *byRank;   // rank -> sequence index
    const unsigned *keys;     // DB keys / entry lengths by sequence index
    const unsigned *lens;
};

// w1 of a k-mer record: narrow id<<32 | seqLen<<16 | pos; wide rank<<32 | pos
__device__ __forceinline__ unsigned long long kmer_w1(const KmConst &c, unsigned id, unsigned si, unsigned seqLen, unsigned pos) {
    if (c.wide) return ((unsigned long long) __ldg(c.rankOf + si) << 32) | (unsigned long long) pos;
    return ((unsigned long long) id << 32) | ((unsigned long long) (seqLen & 0xFFFFu) << 16) | (pos & 0xFFFFu);
}

// One k-mer of a sequence while it is being selected (SequencePosition, kmermatcher.h:10-46).
struct __align__(16) Cand {
    unsigned long long kmer;   // as stored (nt: strand flag in bit 63)
    unsigned score;            // 16-bit hash; 0xFFFFFFFF = padding
    unsigned pos;
};

__device__ __forceinline__ unsigned long long cmp_kmer(unsigned long long k, int nt) { return nt ? (k | (1ULL << 63)) : k; }

// (score, kmer[|bit63], pos) lexicographic -- SequencePosition::compareByScore[Reverse]
__device__ __forceinline__ bool cand_less(const Cand &a, const Cand &b, int nt) {
    if (a.score != b.score) return a.score < b.score;
    const unsigned long long ka = cmp_kmer(a.kmer, nt), kb = cmp_kmer(b.kmer, nt);
    if (ka != kb) return ka < kb;
    return a.pos < b.pos;
}

// k-mer index + score at position pos of the code array; returns false if the window holds X
// (Sequence::kmerContainsX) or is a reverse-complement palindrome (kmermatcher.cpp:156-158).
// KT > 0: k is the compile-time constant KT (loops fully unrolled); KT == 0: k = c.k.  NTM: 1 nucleotides, 0 amino acids,
// -1 decide at run time from c.nt.
template <int KT, int NTM>
__device__ __forceinline__ bool make_kmer_t(const unsigned char *codes, int pos, int L, const KmConst &c, Cand &out) {
    const int k = KT > 0 ? KT : c.k;
    const bool nt = NTM < 0 ? (c.nt != 0) : (NTM != 0);
    bool hasX = false;
    if (nt) {
        unsigned long long idx = 0;
#pragma unroll
        for (int j = 0; j < k; j++) { const unsigned v = codes[pos + j]; hasX |= (v == (unsigned) c.xCode); idx = (idx << 2) | (v & 3u); }
        if (hasX) return false;
        unsigned long long rev = 0, t = idx;     // Util::revComplement (Util.cpp:601-638)
#pragma unroll
        for (int j = 0; j < k; j++) { rev = (rev << 2) | ((t & 3ULL) ^ 2ULL); t >>= 2; }
        if (rev == idx) return false;
        const bool pickRev = rev < idx;
        idx = pickRev ? rev : idx;
        out.score = (unsigned) (xxh64_u64(idx, c.seed) & 0xFFFFULL);
        out.kmer = pickRev ? (idx & ~(1ULL << 63)) : (idx | (1ULL << 63));
        out.pos = pickRev ? (unsigned) (L - pos - k) : (unsigned) pos;
    } else {
        unsigned long long idx = 0, pw = 1;      // Indexer::int2index (Indexer.h:20-83)
#pragma unroll
        for (int j = 0; j < k; j++) { const unsigned v = codes[pos + j]; hasX |= (v == (unsigned) c.xCode); idx += (unsigned long long) v * pw; pw *= c.base; }
        if (hasX) return false;
        out.kmer = idx;
        out.pos = (unsigned) pos;
        out.score = (unsigned) (xxh64_u64(idx, c.seed) & 0xFFFFULL);
    }
    return true;
}
__device__ __forceinline__ bool make_kmer(const unsigned char *codes, int pos, int L, const KmConst &c, Cand &out) {
    return make_kmer_t<0, -1>(codes, pos, L, c, out);
}

// The selection loop of kmermatcher.cpp:274-347 over candidates sorted by cand_less, run by ONE thread.
// sorted[0..cnt) holds all k-mers of the sequence (or at least every k-mer with score <= t);
// emits into outRecs (capacity >= kmerConsidered), returns the number emitted.
__device__ int select_sequential(const Cand *sorted, int cnt, unsigned long long kmerConsidered, unsigned threshold, int tooMuch,
                                 const KmConst &c, unsigned id, unsigned si, unsigned seqLen, Rec *outRecs) {
    int nOut = 0;
    unsigned long long selected = 0;
    for (int i = 0; i < cnt && selected < kmerConsidered; i++) {
        if (c.ignoreMulti) {
            const unsigned long long kmer = cmp_kmer(sorted[i].kmer, c.nt);
            if (i + 1 < cnt) {
                unsigned long long next = cmp_kmer(sorted[i + 1].kmer, c.nt);
                if (kmer == next) {
                    while (kmer == next && i < cnt) {
                        i++;
                        if (i >= cnt) break;
                        next = cmp_kmer(sorted[i].kmer, c.nt);
                    }
                }
            }
            if (i >= cnt) break;
        }
        const unsigned sc = sorted[i].score;
        if (sc < threshold) {
            if (sc == (threshold - 1) && tooMuch) {
                tooMuch--;
                threshold -= (tooMuch == 0) ? 1 : 0;
            }
            selected++;
            if (sc >= c.hashStart && sc <= c.hashEnd) {
                Rec r;
                r.w0 = sorted[i].kmer;
                r.w1 = kmer_w1(c, id, si, seqLen, sorted[i].pos);
                outRecs[nOut++] = r;
            }
        }
    }
    return nOut;
}

// ------------------------------------------------------------------------------------------------
// classify: bin sequences by the number of k-mer windows so that each class runs in a kernel whose
// shared-memory tile fits it.  class 0: <=64, 1: <=256, 2: <=1024, 3: larger (block kernel).
// ------------------------------------------------------------------------------------------------
__global__ void classify_kernel(const unsigned *__restrict__ lens, unsigned lo, unsigned hi, unsigned n, int k, unsigned *__restrict__ lists /*4 x n*/,
                                unsigned *__restrict__ counts /*4*/) {
    const unsigned i = lo + blockIdx.x * blockDim.x + threadIdx.x;   // sequences [lo, hi) of n
    int cls = -1;
    if (i < hi) {
        const int L = (int) lens[i] - 2;
        const int nk = L - k + 1;
        cls = nk <= 64 ? 0 : (nk <= 256 ? 1 : (nk <= 1024 ? 2 : 3));
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, cls == q);
        if (m) {
            unsigned base = 0;
            const int leader = __ffs(m) - 1;
            if ((int) lane_id() == leader) base = atomicAdd(&counts[q], __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (cls == q) lists[(size_t) q * n + base + __popc(m & ((1u << lane_id()) - 1u))] = i;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// extract: one warp per sequence, at most NMAX k-mer windows.
// ------------------------------------------------------------------------------------------------
// Sort key of a k-mer inside one sequence, packed so that (hi, lo) lexicographic == (score, k-mer[|bit63], pos)
// of SequencePosition::compareByScore[Reverse]:  hi = score<<48 | kmer62..15,  lo = kmer14..0<<33 | pos<<1 | strand.
struct __align__(16) PCand { unsigned long long hi, lo; };
__device__ __forceinline__ PCand pack_cand(const Cand &c) {
    const unsigned long long k63 = c.kmer & ~(1ULL << 63);
    PCand p;
    p.hi = ((unsigned long long) c.score << 48) | (k63 >> 15);
    p.lo = ((k63 & 0x7FFFULL) << 33) | ((unsigned long long) c.pos << 1) | (c.kmer >> 63);
    return p;
}
__device__ __forceinline__ unsigned pc_score(const PCand &p) { return (unsigned) (p.hi >> 48); }
__device__ __forceinline__ unsigned long long pc_kmer63(const PCand &p) { return ((p.hi & 0xFFFFFFFFFFFFULL) << 15) | ((p.lo >> 33) & 0x7FFFULL); }
__device__ __forceinline__ unsigned long long pc_kmer_stored(const PCand &p, int nt) { return nt ? (pc_kmer63(p) | ((p.lo & 1ULL) << 63)) : pc_kmer63(p); }
__device__ __forceinline__ unsigned pc_pos(const PCand &p) { return (unsigned) ((p.lo >> 1) & 0xFFFFFFFFULL); }
__device__ __forceinline__ bool pc_less(const PCand &a, const PCand &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ bool pc_same_kmer(const PCand &a, const PCand &b) { return ((a.hi ^ b.hi) & 0xFFFFFFFFFFFFULL) == 0 && ((a.lo ^ b.lo) >> 33) == 0; }

// The selection loop of kmermatcher.cpp:274-347 over packed candidates (slow path: duplicates present).
__device__ int select_sequential_packed(const PCand *sorted, int cnt, unsigned long long kmerConsidered, unsigned threshold, int tooMuch,
                                        const KmConst &c, unsigned id, unsigned si, unsigned seqLen, Rec *outRecs) {
    int nOut = 0;
    unsigned long long selected = 0;
    for (int i = 0; i < cnt && selected < kmerConsidered; i++) {
        if (c.ignoreMulti) {
            if (i + 1 < cnt && pc_same_kmer(sorted[i], sorted[i + 1])) {
                const PCand cur = sorted[i];
                bool same = true;
                while (same && i < cnt) {
                    i++;
                    if (i >= cnt) break;
                    same = pc_same_kmer(cur, sorted[i]);
                }
            }
            if (i >= cnt) break;
        }
        const unsigned sc = pc_score(sorted[i]);
        if (sc < threshold) {
            if (sc == (threshold - 1) && tooMuch) {
                tooMuch--;
                threshold -= (tooMuch == 0) ? 1 : 0;
            }
            selected++;
            if (sc >= c.hashStart && sc <= c.hashEnd) {
                Rec r;
                r.w0 = pc_kmer_stored(sorted[i], c.nt);
                r.w1 = kmer_w1(c, id, si, seqLen, pc_pos(sorted[i]));
                outRecs[nOut++] = r;
            }
        }
    }
    return nOut;
}

// By-product of the extraction: the 256-bin histograms of the first three 8-bit digits of mix64(k-mer) -- the digits sort #1's
// partition passes use (plan_add_hash_bits: shifts 0, 8, 16; the last pass's narrower digit is a fold of its 256 bins).  The
// records are in shared memory when they leave, so counting them here saves the separate histogram sweep (one more read of
// every record) in front of the partition.
constexpr int EXTRACT_HIST_BINS = 3 * 256;
__device__ __forceinline__ void hist_partition_digits(unsigned *sHist, unsigned long long maskedKmer) {
    const unsigned long long h = mix64(maskedKmer);
    atomicAdd(&sHist[(unsigned) h & 255u], 1u);
    atomicAdd(&sHist[256u + ((unsigned) (h >> 8) & 255u)], 1u);
    atomicAdd(&sHist[512u + ((unsigned) (h >> 16) & 255u)], 1u);
}

template <int NMAX, int KT, int NTM>
__global__ void __launch_bounds__(128, 8) extract_warp_kernel(const pg_seqdb db, const unsigned *__restrict__ list,
                                                           const unsigned *__restrict__ listCount, const KmConst c,
                                                           Rec *__restrict__ out, unsigned long long *__restrict__ outCount,
                                                           unsigned long long outCap, unsigned long long *__restrict__ partHist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sHist[EXTRACT_HIST_BINS];
    for (int i = threadIdx.x; i < EXTRACT_HIST_BINS; i += blockDim.x) sHist[i] = 0;
    __syncthreads();
    const unsigned long long histMask = c.nt ? ~(1ULL << 63) : ~0ULL;
    constexpr int WARPS = 4;
    constexpr int CODES = NMAX + 40;   // k <= 32
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    PCand *cand = reinterpret_cast<PCand *>(smem_raw) + (size_t) w * NMAX;
    Rec *outRecs = reinterpret_cast<Rec *>(smem_raw + (size_t) WARPS * NMAX * sizeof(PCand)) + (size_t) w * (NMAX + 1);
    unsigned char *codes = smem_raw + (size_t) WARPS * NMAX * sizeof(PCand) + (size_t) WARPS * (NMAX + 1) * sizeof(Rec) + (size_t) w * CODES;

    // powers of 31 for the whole-sequence hash of short sequences (Util::hash: h = h * 31 + code), one table per CTA
    __shared__ unsigned long long sPow31[NMAX <= 64 ? 128 : 1];
    if (NMAX <= 64) {
        if (threadIdx.x < 128) { unsigned long long pw = 1; for (int j = 0; j < (int) threadIdx.x; j++) pw *= 31ULL; sPow31[threadIdx.x] = pw; }
        __syncthreads();
    }
    const unsigned nList = *listCount;
    // Software pipeline over the warp's sequences, three stages deep: while sequence li is processed, the first 64 residues of
    // li + s, the index entry (offset / length / key) of li + 2s and the list entry of li + 3s are in flight -- every load of an
    // iteration depends only on values requested an iteration earlier.  (Straight code sits through three dependent memory
    // latencies -- list -> offset / length -> residues -- per ~50-residue sequence.)
    const unsigned liStride = gridDim.x * WARPS;
    unsigned li = blockIdx.x * WARPS + w;
    unsigned si0 = 0, si1 = 0, si2 = 0; unsigned long long off0 = 0, off1 = 0; int len0 = 0, len1 = 0;
    unsigned char pre0 = 0, pre1 = 0;
    if (li < nList) {
        si0 = list[li]; off0 = db.offsets[si0]; len0 = (int) db.lens[si0] - 2;
        if (lane < len0) pre0 = (unsigned char) db.data[off0 + lane];
        if (lane + 32 < len0) pre1 = (unsigned char) db.data[off0 + lane + 32];
    }
    if ((unsigned long long) li + liStride < nList) { si1 = list[li + liStride]; off1 = db.offsets[si1]; len1 = (int) db.lens[si1] - 2; }
    if ((unsigned long long) li + 2ull * liStride < nList) si2 = list[li + 2 * liStride];
    for (; li < nList; li += liStride) {
        const unsigned si = si0;
        const char *seq = db.data + off0;
        const int entryLen = len0;
        const unsigned char cur0 = pre0, cur1 = pre1;
        const unsigned long long rest = (unsigned long long) nList - li;      // > 0; sequence li + a*s exists iff a * s < rest
        // stage 1 -> 0: the next sequence's residues are requested (its index entry arrived during the last iteration)
        si0 = si1; off0 = off1; len0 = len1;
        pre0 = 0; pre1 = 0;
        if (liStride < rest) {
            if (lane < len0) pre0 = (unsigned char) db.data[off0 + lane];
            if (lane + 32 < len0) pre1 = (unsigned char) db.data[off0 + lane + 32];
        }
        // stage 2 -> 1: index entry of the sequence after next; stage 3 -> 2: list entry of the one after that
        if (2ull * liStride < rest) { si1 = si2; off1 = db.offsets[si1]; len1 = (int) db.lens[si1] - 2; }
        if (3ull * liStride < rest) si2 = list[li + 3 * liStride];
        const unsigned id = db.keys[si];         // needed only when the records are written
        // Sequence::mapSequence (Sequence.cpp:476-489): map until '\n' / '\0'
        int L = entryLen;
        if (lane < entryLen) { codes[lane] = c_aa2num[cur0]; if (cur0 == '\n' || cur0 == 0) L = min(L, (int) lane); }
        if (lane + 32 < entryLen) { codes[lane + 32] = c_aa2num[cur1]; if (cur1 == '\n' || cur1 == 0) L = min(L, (int) lane + 32); }
        for (int i = lane + 64; i < entryLen; i += 32) {
            const unsigned char ch = (unsigned char) seq[i];
            codes[i] = c_aa2num[ch];
            if (ch == '\n' || ch == 0) L = min(L, i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) L = min(L, __shfl_xor_sync(0xFFFFFFFFu, L, o));
        __syncwarp();
        // whole-sequence hash: Util::hash (poly 31) then XXH64 (kmermatcher.cpp:133-138).  Each lane hashes a
        // chunk; H(AB) = H(A) * 31^|B| + H(B) is associative, so the chunks fold in log2(32) shuffle steps.
        unsigned long long seqHash;
        if (NMAX <= 64 && L <= 128) {
            // sum(code[i] * 31^(L-1-i)) == the Horner form, modulo 2^64: every lane takes the positions lane, lane + 32, ...
            unsigned long long h = 0;
            for (int i = lane; i < L; i += 32) h += (unsigned long long) codes[i] * sPow31[L - 1 - i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
            seqHash = xxh64_u64(h, c.seed);
        } else {
            const int chunk = (L + 31) / 32;
            const int b = min(L, lane * chunk), e = min(L, b + chunk);
            unsigned long long h = 0, pw = 1;
            for (int i = b; i < e; i++) { h = h * 31ULL + codes[i]; pw *= 31ULL; }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long hr = __shfl_down_sync(0xFFFFFFFFu, h, o);
                const unsigned long long pr = __shfl_down_sync(0xFFFFFFFFu, pw, o);
                if ((lane & (2 * o - 1)) == 0) { h = h * pr + hr; pw *= pr; }
            }
            seqHash = xxh64_u64(__shfl_sync(0xFFFFFFFFu, h, 0), c.seed);
        }
        // all k-mer windows, compacted in position order
        int cnt = 0;
        bool scoreless = false;
        const int nWin = L - (KT > 0 ? KT : c.k) + 1;
        if constexpr (KT > 0 && NTM == 0 && (KT & 1) == 0) {
            // amino acids, even compile-time k: Indexer::int2index (Indexer.h:20-83) is sum(code[pos+j] * base^j).  The two
            // halves of a window are the half-sums of positions pos and pos + k/2, so every position computes ONE k/2-term
            // sum in 32-bit arithmetic (base^(k/2) <= 20^7 < 2^32) and a window is half[pos] + half[pos+k/2] * base^(k/2):
            // ~k/2 narrow multiplies per window instead of k wide ones.  X windows (Sequence::kmerContainsX) are found
            // from a bit mask of the X positions.  Both arrays alias the output staging area, which is not in use yet.
            constexpr int H = KT / 2;
            unsigned *half = reinterpret_cast<unsigned *>(outRecs);
            unsigned *xmask = half + CODES;
            unsigned baseH = 1;
#pragma unroll
            for (int j = 0; j < H; j++) baseH *= c.base;
            for (int p0 = 0; p0 < L; p0 += 32) {
                const int pos = p0 + lane;
                const bool isX = pos < L && codes[pos] == (unsigned char) c.xCode;
                const unsigned xm = __ballot_sync(0xFFFFFFFFu, isX);
                if (lane == 0) xmask[p0 >> 5] = xm;
                if (pos + H <= L) {
                    unsigned hs = 0, pw = 1;
#pragma unroll
                    for (int j = 0; j < H; j++) { hs += (unsigned) codes[pos + j] * pw; pw *= c.base; }
                    half[pos] = hs;
                }
            }
            if (lane == 0) xmask[(L + 31) >> 5] = 0;
            __syncwarp();
            // A read has fewer windows than the bottom-m budget (kmersPerSeq - 1 + scale * L >= number of windows): every k-mer
            // is taken whatever its score, and with the whole hash range wanted the score is not needed at all -- no XXH64 per
            // window.  (If the sequence turns out to repeat a k-mer, the scores are computed after all, below.)
            scoreless = c.ignoreMulti && c.hashStart == 0 && c.hashEnd >= 65535u && nWin > 0 &&
                        (unsigned long long) ((float) (c.kmersPerSeq - 1) + (c.scale * (float) L)) >= (unsigned long long) nWin;
            if (NMAX <= 64 && scoreless) {
                // Fused path of a read: window -> k-mer -> duplicate check -> record, no candidate list, no score.  At most two
                // windows per lane (NMAX = 64), held in registers while the half sums (which alias the staging area) are alive.
                unsigned long long km[2]; bool okk[2];
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int pos = r * 32 + lane;
                    bool ok = pos < nWin;
                    if (ok) {
                        const unsigned long long xb = ((unsigned long long) xmask[pos >> 5] | ((unsigned long long) xmask[(pos >> 5) + 1] << 32)) >> (pos & 31);
                        ok = (xb & ((1ULL << KT) - 1ULL)) == 0;
                    }
                    km[r] = ok ? (unsigned long long) half[pos] + (unsigned long long) half[pos + H] * (unsigned long long) baseH : 0ULL;
                    okk[r] = ok;
                }
                __syncwarp();                                 // half / xmask are dead: the staging area takes the records
                unsigned long long *set = reinterpret_cast<unsigned long long *>(cand);        // 2 * NMAX slots in the (unused) candidate area
                constexpr unsigned SLOTS = 2 * NMAX;
                for (int i = lane; i < (int) SLOTS; i += 32) set[i] = ~0ULL;
                __syncwarp();
                bool dup = false;
                int n = 0;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, okk[r]);
                    if (okk[r]) {
                        unsigned slot = (unsigned) ((km[r] * 0x9E3779B97F4A7C15ULL) >> 40) & (SLOTS - 1);
                        while (true) {
                            const unsigned long long old = atomicCAS(&set[slot], ~0ULL, km[r]);
                            if (old == ~0ULL) break;
                            if (old == km[r]) { dup = true; break; }
                            slot = (slot + 1) & (SLOTS - 1);
                        }
                        Rec rr; rr.w0 = km[r]; rr.w1 = kmer_w1(c, id, si, (unsigned) L, (unsigned) (r * 32 + lane));
                        outRecs[1 + n + __popc(m & ltMask)] = rr;
                    }
                    n += __popc(m);
                }
                if (__ballot_sync(0xFFFFFFFFu, dup) == 0) {
                    // every k-mer is taken (kmermatcher.cpp:274-347 with kmerConsidered >= the number of k-mers, all distinct);
                    // the sequence-identity record goes first (:241-246)
                    if (lane == 0) { Rec r0; r0.w0 = seqHash; r0.w1 = kmer_w1(c, id, si, (unsigned) L, 0u); outRecs[0] = r0; }
                    const int nOutF = n + 1;
                    __syncwarp();
                    unsigned long long baseF = 0;
                    if (lane == 0) baseF = atomicAdd(outCount, (unsigned long long) nOutF);
                    baseF = __shfl_sync(0xFFFFFFFFu, baseF, 0);
                    if (baseF + nOutF <= outCap)
                        for (int i = lane; i < nOutF; i += 32) { const Rec r = outRecs[i]; out[baseF + i] = r; hist_partition_digits(sHist, r.w0 & histMask); }
                    __syncwarp();
                    continue;
                }
                // a repeated k-mer (rare): the sorted walk decides, with scores, from the candidate list
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, okk[r]);
                    if (okk[r]) {
                        Cand cd; cd.kmer = km[r]; cd.pos = (unsigned) (r * 32 + lane);
                        cd.score = (unsigned) (xxh64_u64(km[r], c.seed) & 0xFFFFULL);
                        cand[cnt + __popc(m & ltMask)] = pack_cand(cd);
                    }
                    cnt += __popc(m);
                }
                scoreless = false;
                __syncwarp();
            } else
            for (int p0 = 0; p0 < nWin; p0 += 32) {
                const int pos = p0 + lane;
                bool ok = pos < nWin;
                Cand cd; cd.kmer = 0; cd.score = 0; cd.pos = (unsigned) pos;
                if (ok) {
                    const unsigned long long xb = ((unsigned long long) xmask[pos >> 5] | ((unsigned long long) xmask[(pos >> 5) + 1] << 32)) >> (pos & 31);
                    ok = (xb & ((1ULL << KT) - 1ULL)) == 0;
                }
                if (ok) {
                    cd.kmer = (unsigned long long) half[pos] + (unsigned long long) half[pos + H] * (unsigned long long) baseH;
                    if (!scoreless) cd.score = (unsigned) (xxh64_u64(cd.kmer, c.seed) & 0xFFFFULL);
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) cand[cnt + __popc(m & ltMask)] = pack_cand(cd);
                cnt += __popc(m);
            }
            __syncwarp();                                     // half / xmask are dead from here on: the staging area is reused
        } else {
            for (int p0 = 0; p0 < nWin; p0 += 32) {
                const int pos = p0 + lane;
                Cand cd;
                const bool ok = (pos < nWin) && make_kmer_t<KT, NTM>(codes, pos, L, c, cd);
                const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) cand[cnt + __popc(m & ltMask)] = pack_cand(cd);
                cnt += __popc(m);
            }
        }
        // threshold of the bottom-m sketch (kmermatcher.cpp:223-238)
        const unsigned long long want = (unsigned long long) ((float) (c.kmersPerSeq - 1) + (c.scale * (float) L));
        const unsigned long long kmerConsidered = min(want, (unsigned long long) cnt);
        // Short-cut for the common case of reads: every k-mer is selected (cnt <= kmerConsidered) and no k-mer occurs
        // twice in the sequence.  Then the sorted walk of kmermatcher.cpp:274-347 selects all of them and their order is
        // irrelevant (the records are re-ordered globally afterwards), so neither the sort nor the walk is needed.
        // Duplicates are detected with a small open-addressing set in shared memory (the output staging area).
        bool noDup = false;
        if (c.ignoreMulti && cnt > 0) {
            unsigned long long *set = reinterpret_cast<unsigned long long *>(outRecs);
            constexpr unsigned SLOTS = 2 * NMAX;
            for (int i = lane; i < (int) SLOTS; i += 32) set[i] = ~0ULL;
            __syncwarp();
            bool dup = false;
            for (int i = lane; i < cnt; i += 32) {
                const unsigned long long k63 = pc_kmer63(cand[i]);
                unsigned slot = (unsigned) ((k63 * 0x9E3779B97F4A7C15ULL) >> 40) & (SLOTS - 1);     // multiplicative hash: the set only has to spread
                while (true) {
                    const unsigned long long old = atomicCAS(&set[slot], ~0ULL, k63);
                    if (old == ~0ULL) break;
                    if (old == k63) { dup = true; break; }
                    slot = (slot + 1) & (SLOTS - 1);
                }
            }
            noDup = __ballot_sync(0xFFFFFFFFu, dup) == 0;
            __syncwarp();
            if (scoreless && !noDup) {
                // rare: a repeated k-mer sends the sequence through the sorted walk, which orders by score
                for (int i = lane; i < cnt; i += 32) {
                    Cand cd; cd.kmer = pc_kmer_stored(cand[i], c.nt); cd.pos = pc_pos(cand[i]);
                    cd.score = (unsigned) (xxh64_u64(pc_kmer63(cand[i]), c.seed) & 0xFFFFULL);
                    cand[i] = pack_cand(cd);
                }
                __syncwarp();
                scoreless = false;
            }
        }
        const bool allDistinct = noDup && kmerConsidered >= (unsigned long long) cnt;
        // Sort-free selection when only a part of the (distinct) k-mers is taken -- the nucleotide workflow keeps
        // 59 + 0.1 L of the L - 21 k-mers of a read.  The sorted walk of kmermatcher.cpp:274-347 then selects every k-mer
        // whose score lies below the score t of the kmerConsidered-th smallest, plus the first `tooMuch` (all, if
        // tooMuch == 0) of the k-mers with score t in (k-mer, position) order, capped at kmerConsidered.  t comes from a
        // 16-step radix select over the scores; only the (usually one or two) k-mers of the last bin need an order.
        bool quick = false, keepMine = false;
        unsigned qT = 0;
        PCand mine; mine.hi = 0; mine.lo = 0;
        if (noDup && !allDistinct && kmerConsidered > 0) {
            const int k = (int) kmerConsidered;
            unsigned t = 0;
            for (int bit = 15; bit >= 0; bit--) {
                const unsigned trial = t | (1u << bit);
                int below = 0;
                for (int p0 = 0; p0 < cnt; p0 += 32) {
                    const int i = p0 + lane;
                    below += __popc(__ballot_sync(0xFFFFFFFFu, i < cnt && pc_score(cand[i]) < trial));
                }
                if (below < k) t = trial;                    // the k-th smallest score is >= trial
            }
            int below = 0, inBins = 0;
            for (int p0 = 0; p0 < cnt; p0 += 32) {
                const int i = p0 + lane;
                const unsigned sc = i < cnt ? pc_score(cand[i]) : 0xFFFFFFFFu;
                below += __popc(__ballot_sync(0xFFFFFFFFu, sc < t));
                inBins += __popc(__ballot_sync(0xFFFFFFFFu, sc <= t));
            }
            const int lastBin = inBins - below, tooMuch = inBins - k;
            if (lastBin <= 32) {
                quick = true;
                qT = t;
                int nLast = tooMuch > 0 ? min(tooMuch, lastBin) : lastBin;
                nLast = min(nLast, k - below);
                PCand *lastList = reinterpret_cast<PCand *>(outRecs);     // the set is dead, the staging area not yet in use
                int base = 0;
                for (int p0 = 0; p0 < cnt; p0 += 32) {
                    const int i = p0 + lane;
                    const bool is = i < cnt && pc_score(cand[i]) == t;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, is);
                    if (is) lastList[base + __popc(m & ltMask)] = cand[i];
                    base += __popc(m);
                }
                __syncwarp();
                if (lane < lastBin) {
                    mine = lastList[lane];
                    int rank = 0;
                    for (int j = 0; j < lastBin; j++) rank += pc_less(lastList[j], mine) ? 1 : 0;
                    keepMine = rank < nLast;
                }
                __syncwarp();
            }
        }
        if (!allDistinct && !quick) {
            int n2 = 1;
            while (n2 < cnt) n2 <<= 1;
            for (int i = cnt + lane; i < n2; i += 32) { cand[i].hi = ~0ULL; cand[i].lo = ~0ULL; }
            __syncwarp();
            // bitonic sort by (score, kmer, pos)   [std::sort at kmermatcher.cpp:266-272; total order => same result]
            if (c.ignoreMulti) {
                for (int kk = 2; kk <= n2; kk <<= 1) {
                    for (int j = kk >> 1; j > 0; j >>= 1) {
                        for (int t = lane; t < (n2 >> 1); t += 32) {
                            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int ix = i | j;
                            const bool up = ((i & kk) == 0);
                            const PCand a = cand[i], b = cand[ix];
                            if (pc_less(b, a) == up) { cand[i] = b; cand[ix] = a; }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        int nOut = 0;
        // sequence-identity record first (:241-246)
        const unsigned sh16 = (unsigned) (seqHash & 0xFFFFULL);
        if (sh16 >= c.hashStart && sh16 <= c.hashEnd) {
            if (lane == 0) { Rec r; r.w0 = seqHash; r.w1 = kmer_w1(c, id, si, (unsigned) L, 0u); outRecs[0] = r; }
            nOut = 1;
        }
        if (allDistinct) {
            // all cnt k-mers are selected (unsorted), subject only to the shard's hash range
            for (int p0 = 0; p0 < cnt; p0 += 32) {
                const int i = p0 + lane;
                bool emit = false; PCand pc; pc.hi = 0; pc.lo = 0;
                if (i < cnt) { pc = cand[i]; const unsigned sc = pc_score(pc); emit = sc >= c.hashStart && sc <= c.hashEnd; }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
                if (emit) {
                    Rec r;
                    r.w0 = pc_kmer_stored(pc, c.nt);
                    r.w1 = kmer_w1(c, id, si, (unsigned) L, pc_pos(pc));
                    outRecs[nOut + __popc(m & ltMask)] = r;
                }
                nOut += __popc(m);
            }
        } else if (quick) {
            for (int p0 = 0; p0 < cnt; p0 += 32) {
                const int i = p0 + lane;
                bool emit = false; PCand pc; pc.hi = 0; pc.lo = 0;
                if (i < cnt) { pc = cand[i]; const unsigned sc = pc_score(pc); emit = sc < qT && sc >= c.hashStart && sc <= c.hashEnd; }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
                if (emit) {
                    Rec r;
                    r.w0 = pc_kmer_stored(pc, c.nt);
                    r.w1 = kmer_w1(c, id, si, (unsigned) L, pc_pos(pc));
                    outRecs[nOut + __popc(m & ltMask)] = r;
                }
                nOut += __popc(m);
            }
            {
                const bool emit = keepMine && qT >= c.hashStart && qT <= c.hashEnd;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
                if (emit) {
                    Rec r;
                    r.w0 = pc_kmer_stored(mine, c.nt);
                    r.w1 = kmer_w1(c, id, si, (unsigned) L, pc_pos(mine));
                    outRecs[nOut + __popc(m & ltMask)] = r;
                }
                nOut += __popc(m);
            }
        } else if (cnt > 0 && kmerConsidered > 0) {
            if (c.ignoreMulti) {
                // t = score of the kmerConsidered-th smallest; candidates = scores <= t; ties of the last bin = tooMuch
                const unsigned t = pc_score(cand[kmerConsidered - 1]);
                int below = 0, inBins = 0, dups = 0;
                for (int p0 = 0; p0 < cnt; p0 += 32) {
                    const int i = p0 + lane;
                    const bool in = i < cnt;
                    const unsigned sc = in ? pc_score(cand[i]) : 0xFFFFFFFFu;
                    below += __popc(__ballot_sync(0xFFFFFFFFu, in && sc < t));
                    inBins += __popc(__ballot_sync(0xFFFFFFFFu, in && sc <= t));
                    dups += __popc(__ballot_sync(0xFFFFFFFFu, in && i + 1 < cnt && pc_same_kmer(cand[i], cand[i + 1])));
                }
                const unsigned threshold = t + 1;
                const int tooMuch = inBins - (int) kmerConsidered;
                if (dups == 0) {
                    // no repeated k-mer: every candidate is visited, so the loop selects the elements below the last
                    // bin and then the first `tooMuch` (all, if tooMuch == 0) of the last bin, capped at kmerConsidered
                    const int lastBin = inBins - below;
                    int nSel = below + (tooMuch > 0 ? min(tooMuch, lastBin) : lastBin);
                    if ((unsigned long long) nSel > kmerConsidered) nSel = (int) kmerConsidered;
                    for (int p0 = 0; p0 < nSel; p0 += 32) {
                        const int i = p0 + lane;
                        bool emit = false; PCand pc; pc.hi = 0; pc.lo = 0;
                        if (i < nSel) { pc = cand[i]; const unsigned sc = pc_score(pc); emit = sc >= c.hashStart && sc <= c.hashEnd; }
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
                        if (emit) {
                            Rec r;
                            r.w0 = pc_kmer_stored(pc, c.nt);
                            r.w1 = kmer_w1(c, id, si, (unsigned) L, pc_pos(pc));
                            outRecs[nOut + __popc(m & ltMask)] = r;
                        }
                        nOut += __popc(m);
                    }
                } else {
                    int add = 0;
                    if (lane == 0) add = select_sequential_packed(cand, cnt, kmerConsidered, threshold, tooMuch, c, id, si, (unsigned) L, outRecs + nOut);
                    nOut += __shfl_sync(0xFFFFFFFFu, add, 0);
                }
            } else {
                // positional order kept (--ignore-multi-kmer 0): the kmerConsidered-th smallest score by counting
                int add = 0;
                if (lane == 0) {
                    unsigned lo = 0, hi = 65535;
                    while (lo < hi) {
                        const unsigned mid = (lo + hi) >> 1;
                        int le = 0;
                        for (int i = 0; i < cnt; i++) le += (pc_score(cand[i]) <= mid);
                        if ((unsigned long long) le >= kmerConsidered) hi = mid; else lo = mid + 1;
                    }
                    int le = 0;
                    for (int i = 0; i < cnt; i++) le += (pc_score(cand[i]) <= lo);
                    add = select_sequential_packed(cand, cnt, kmerConsidered, lo + 1, le - (int) kmerConsidered, c, id, si, (unsigned) L, outRecs + nOut);
                }
                nOut += __shfl_sync(0xFFFFFFFFu, add, 0);
            }
        }
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0 && nOut) base = atomicAdd(outCount, (unsigned long long) nOut);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base + nOut <= outCap)
            for (int i = lane; i < nOut; i += 32) { const Rec r = outRecs[i]; out[base + i] = r; hist_partition_digits(sHist, r.w0 & histMask); }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < EXTRACT_HIST_BINS; i += blockDim.x)
        if (sHist[i]) atomicAdd(partHist + i, (unsigned long long) sHist[i]);
}

// ------------------------------------------------------------------------------------------------
// extract: one block per long sequence (more than 1024 windows); candidates live in global scratch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) extract_block_kernel(const pg_seqdb db, const unsigned *__restrict__ list,
                                                            const unsigned *__restrict__ listCount, const KmConst c,
                                                            Rec *__restrict__ out, unsigned long long *__restrict__ outCount,
                                                            unsigned long long outCap, unsigned char *__restrict__ scratch,
                                                            size_t scratchPerBlock, unsigned long long *__restrict__ partHist) {
    __shared__ unsigned sHist[EXTRACT_HIST_BINS];
    for (int i = threadIdx.x; i < EXTRACT_HIST_BINS; i += 256) sHist[i] = 0;
    __syncthreads();
    const unsigned long long histMask = c.nt ? ~(1ULL << 63) : ~0ULL;
    __shared__ unsigned hier[128];
    __shared__ unsigned fine[512];
    __shared__ unsigned sCnt;
    __shared__ unsigned long long sPart[256], sPow[256];
    __shared__ unsigned sThreshold, sCoarse;
    __shared__ int sTooMuch, sInBins, sNOut;
    __shared__ unsigned long long sBase;
    const int tid = threadIdx.x;
    const unsigned nList = *listCount;
    unsigned char *myScratch = scratch + (size_t) blockIdx.x * scratchPerBlock;
    for (unsigned li = blockIdx.x; li < nList; li += gridDim.x) {
        const unsigned si = list[li];
        const char *seq = db.data + db.offsets[si];
        const int entryLen = (int) db.lens[si] - 2;
        // scratch layout: codes[entryLen+1 rounded to 16] | Cand cand[n2] | Rec outRecs[...]
        unsigned char *codes = myScratch;
        const size_t codesBytes = ((size_t) entryLen + 16) & ~(size_t) 15;
        Cand *cand = reinterpret_cast<Cand *>(myScratch + codesBytes);
        __shared__ int sL;
        if (tid == 0) sL = entryLen;
        __syncthreads();
        for (int i = tid; i < entryLen; i += 256) {
            const unsigned char ch = (unsigned char) seq[i];
            codes[i] = c_aa2num[ch];
            if (ch == '\n' || ch == 0) atomicMin(&sL, i);
        }
        __syncthreads();
        const int L = sL;
        const unsigned id = db.keys[si];
        // sequence hash, chunked polynomial
        {
            const int chunk = (L + 255) / 256;
            const int b = min(L, tid * chunk), e = min(L, b + chunk);
            unsigned long long h = 0, pw = 1;
            for (int i = b; i < e; i++) { h = h * 31ULL + codes[i]; pw *= 31ULL; }
            sPart[tid] = h; sPow[tid] = pw;
        }
        for (int i = tid; i < 128; i += 256) hier[i] = 0;
        for (int i = tid; i < 512; i += 256) fine[i] = 0;
        if (tid == 0) sCnt = 0;
        __syncthreads();
        const int nWin = L - c.k + 1;
        // pass 1: coarse histogram + count
        for (int pos = tid; pos < nWin; pos += 256) {
            Cand cd;
            if (make_kmer(codes, pos, L, c, cd)) { atomicAdd(&hier[cd.score >> 9], 1u); atomicAdd(&sCnt, 1u); }
        }
        __syncthreads();
        const int cnt = (int) sCnt;
        const unsigned long long want = (unsigned long long) ((float) (c.kmersPerSeq - 1) + (c.scale * (float) L));
        const unsigned long long kmerConsidered = min(want, (unsigned long long) cnt);
        if (tid == 0) {
            unsigned long long acc = 0;
            for (int l = 0; l < 256; l++) acc = acc * sPow[l] + sPart[l];
            const unsigned long long seqHash = xxh64_u64(acc, c.seed);
            sNOut = 0;
            Rec *outRecs = reinterpret_cast<Rec *>(cand);   // reused later; identity record kept in registers instead
            (void) outRecs;
            const unsigned sh16 = (unsigned) (seqHash & 0xFFFFULL);
            sBase = seqHash;
            sInBins = (sh16 >= c.hashStart && sh16 <= c.hashEnd) ? 1 : 0;   // temporarily: "emit identity record"
            // coarse walk (kmermatcher.cpp:227-232)
            unsigned long long inBins = 0; unsigned ht = 0;
            if (cnt > 0) {
                for (ht = 0; ht < 128 && inBins < kmerConsidered; ht++) inBins += hier[ht];
                ht -= (ht > 0) ? 1 : 0;
            }
            sCoarse = ht;
        }
        __syncthreads();
        const bool emitIdentity = sInBins != 0;
        const unsigned long long seqHash = sBase;
        const unsigned coarse = sCoarse;
        __syncthreads();
        // pass 2: fine histogram of the coarse bin that holds the threshold
        for (int pos = tid; pos < nWin; pos += 256) {
            Cand cd;
            if (make_kmer(codes, pos, L, c, cd) && (cd.score >> 9) == coarse) atomicAdd(&fine[cd.score & 511u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long inBins = 0;
            for (unsigned h = 0; h < coarse; h++) inBins += hier[h];
            unsigned threshold = coarse * 512;
            if (cnt > 0 && kmerConsidered > 0) {
                for (; threshold <= 65535u && inBins < kmerConsidered; threshold++) inBins += fine[threshold - coarse * 512];
            } else {
                threshold = 0; inBins = kmerConsidered;
            }
            sThreshold = threshold;
            sInBins = (int) inBins;
            sTooMuch = (int) (inBins - kmerConsidered);
            sCnt = 0;
        }
        __syncthreads();
        const unsigned threshold = sThreshold;
        const int inBins = sInBins;
        // pass 3: gather candidates (score < threshold; without ignoreMulti positional order is needed,
        // which this block path does not provide -> handled by requiring ignoreMulti, checked on the host)
        for (int pos = tid; pos < nWin; pos += 256) {
            Cand cd;
            if (make_kmer(codes, pos, L, c, cd) && cd.score < threshold) cand[atomicAdd(&sCnt, 1u)] = cd;
        }
        __syncthreads();
        int n2 = 1;
        while (n2 < inBins) n2 <<= 1;
        for (int i = inBins + tid; i < n2; i += 256) { cand[i].score = 0xFFFFFFFFu; cand[i].kmer = ~0ULL; cand[i].pos = 0xFFFFFFFFu; }
        __syncthreads();
        for (int kk = 2; kk <= n2; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (n2 >> 1); t += 256) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int ix = i | j;
                    const bool up = ((i & kk) == 0);
                    const Cand a = cand[i], b = cand[ix];
                    if (cand_less(b, a, c.nt) == up) { cand[i] = b; cand[ix] = a; }
                }
                __syncthreads();
            }
        }
        Rec *outRecs = reinterpret_cast<Rec *>(cand + n2);
        if (tid == 0) {
            int nOut = 0;
            if (emitIdentity) {
                Rec r; r.w0 = seqHash; r.w1 = kmer_w1(c, id, si, (unsigned) L, 0u);
                outRecs[nOut++] = r;
            }
            if (cnt > 0 && kmerConsidered > 0)
                nOut += select_sequential(cand, inBins, kmerConsidered, threshold, sTooMuch, c, id, si, (unsigned) L, outRecs + nOut);
            sNOut = nOut;
            sBase = nOut ? atomicAdd(outCount, (unsigned long long) nOut) : 0ULL;
        }
        __syncthreads();
        const int nOut = sNOut;
        const unsigned long long base = sBase;
        if (base + nOut <= outCap)
            for (int i = tid; i < nOut; i += 256) { const Rec r = outRecs[i]; out[base + i] = r; hist_partition_digits(sHist, r.w0 & histMask); }
        __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < EXTRACT_HIST_BINS; i += 256)
        if (sHist[i]) atomicAdd(partHist + i, (unsigned long long) sHist[i]);
}

// ------------------------------------------------------------------------------------------------
// group: assignGroup (kmermatcher.cpp:450-559) on records sorted by k-mer.
// ------------------------------------------------------------------------------------------------
constexpr int GROUP_THREADS = 256;
constexpr int GROUP_ITEMS = 4;
constexpr int GROUP_TILE = GROUP_THREADS * GROUP_ITEMS;   // 1024 records (static shared memory stays below 48 KiB)

struct GroupAcc {            // partial reduction of one k-mer group
    unsigned long long key;  // min of rep_key(w1)
    unsigned strand;         // bit 63 of the k-mer word of the record holding `key`
    unsigned count;
};
__device__ __forceinline__ unsigned long long rep_key(unsigned long long w1) {
    // order (seqLen desc, id asc, pos asc) -- compareRepSequenceAndIdAndPos (kmermatcher.h:56-96)
    const unsigned long long id = w1 >> 32, len = (w1 >> 16) & 0xFFFFULL, pos = w1 & 0xFFFFULL;
    return ((0xFFFFULL ^ len) << 48) | (id << 16) | pos;
}
__device__ __forceinline__ void acc_add(GroupAcc &a, unsigned long long key, unsigned strand, unsigned count) {
    if (count == 0) return;
    if (a.count == 0 || key < a.key || (key == a.key && strand < a.strand)) { a.key = key; a.strand = strand; }
    a.count += count;
}

// Util::canBeCovered (Util.cpp:533-551)
__device__ __forceinline__ bool can_be_covered(float covThr, int covMode, float q, float t) {
    // -c 0 (the assemble workflows): every ratio of two positive lengths passes, no division needed.  Modes 3 / 4 also
    // bound the ratio from above and zero lengths give NaN / inf ratios: those keep the reference's arithmetic.
    if (covThr <= 0.0f && q > 0.0f && t > 0.0f && covMode != 3 && covMode != 4) return true;
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (t / q) >= covThr;
        case 2: return (q / t) >= covThr;
        case 3: return (t / q) >= covThr && (t / q) <= 1.0f;
        case 4: return (q / t) >= covThr && (q / t) <= 1.0f;
        case 5: return (fminf(t, q) / fmaxf(t, q)) >= covThr;
        default: return true;
    }
}

// WIDE: records carry rank<<32 | pos (see KmConst); w1 itself is then the representative order, ids / lengths come from the
// rank tables, nothing is truncated to 16 bits and the pair record holds a 32-bit biased diagonal with the strand in bit 32.
template <bool WIDE>
__device__ __forceinline__ unsigned long long rep_key_t(unsigned long long w1) { return WIDE ? w1 : rep_key(w1); }

// firstKmer (optional): the records are a SUBSET of the job's k-mer records (spill list of the bucketed path); the group of
// assignGroup's first-group quirk is then named by its k-mer (strand bit cleared) instead of by its position.
template <bool WIDE>
__global__ void __launch_bounds__(GROUP_THREADS) group_kernel(const Rec *__restrict__ in, unsigned long long n, const KmConst c,
                                                              Rec *__restrict__ out, unsigned long long *__restrict__ outCount,
                                                              const unsigned long long *__restrict__ firstKmer = nullptr) {
    __shared__ Rec tile[GROUP_TILE];
    __shared__ int headIdx[GROUP_TILE];       // index (in tile) of the first record of the group of record i
    __shared__ GroupAcc gacc[GROUP_TILE];     // valid at head indices: reduction of the whole group
    __shared__ GroupAcc sBack, sFwd;
    __shared__ int sBackAtZero;
    __shared__ unsigned sWarpOut[GROUP_THREADS / 32];
    __shared__ unsigned long long sOutBase;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nt = c.nt;
    const unsigned long long tileBase = (unsigned long long) blockIdx.x * GROUP_TILE;
    const int count = (int) min((unsigned long long) GROUP_TILE, n - tileBase);
    for (int i = tid; i < count; i += GROUP_THREADS) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(in + tileBase) + i);
        Rec r; r.w0 = ((unsigned long long) raw.y << 32) | raw.x; r.w1 = ((unsigned long long) raw.w << 32) | raw.z;
        tile[i] = r;
    }
    __syncthreads();
    // warp 0 walks backwards from the tile start, warp 1 forwards from the tile end, over the
    // records of the groups cut by the tile boundary.
    if (w == 0) {
        GroupAcc a; a.key = 0; a.strand = 0; a.count = 0;
        const unsigned long long k0 = cmp_kmer(tile[0].w0, nt);
        long long p = (long long) tileBase - 1 - lane;
        int atZero = (tileBase == 0);
        bool go = true;
        while (go) {
            bool same = false; unsigned long long key = 0; unsigned st = 0;
            if (p >= 0) {
                const Rec r = in[p];
                same = cmp_kmer(r.w0, nt) == k0;
                key = rep_key_t<WIDE>(r.w1); st = (unsigned) (r.w0 >> 63);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, same);
            // records are contiguous: count matches from lane 0 upwards until the first mismatch
            const unsigned firstMiss = __ffs(~m);     // 1-based lane of first non-match, 0 if all match
            const int take = firstMiss ? (int) firstMiss - 1 : 32;
            for (int l = 0; l < take; l++) {
                const unsigned long long kl = __shfl_sync(0xFFFFFFFFu, key, l);
                const unsigned sl = __shfl_sync(0xFFFFFFFFu, st, l);
                acc_add(a, kl, sl, 1);
            }
            if (take == 32) {
                p -= 32;
                if (__shfl_sync(0xFFFFFFFFu, p, 0) < 0) { go = false; atZero = 1; }
            } else {
                go = false;
                // group start is global index 0 iff we ran off the front
                const long long lastTaken = __shfl_sync(0xFFFFFFFFu, p, 0) - (take - 1);
                atZero = (take > 0) ? (lastTaken == 0) : (tileBase == 0);
            }
        }
        if (lane == 0) { sBack = a; sBackAtZero = atZero; }
    } else if (w == 1) {
        GroupAcc a; a.key = 0; a.strand = 0; a.count = 0;
        const unsigned long long k1 = cmp_kmer(tile[count - 1].w0, nt);
        unsigned long long p = tileBase + count + lane;
        bool go = true;
        while (go) {
            bool same = false; unsigned long long key = 0; unsigned st = 0;
            if (p < n) {
                const Rec r = in[p];
                same = cmp_kmer(r.w0, nt) == k1;
                key = rep_key_t<WIDE>(r.w1); st = (unsigned) (r.w0 >> 63);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, same);
            const unsigned firstMiss = __ffs(~m);
            const int take = firstMiss ? (int) firstMiss - 1 : 32;
            for (int l = 0; l < take; l++) {
                const unsigned long long kl = __shfl_sync(0xFFFFFFFFu, key, l);
                const unsigned sl = __shfl_sync(0xFFFFFFFFu, st, l);
                acc_add(a, kl, sl, 1);
            }
            if (take == 32) p += 32; else go = false;
        }
        if (lane == 0) sFwd = a;
    }
    // head flags -> head index by a running max inside each thread's contiguous chunk + block scan
    {
        const int b = tid * GROUP_ITEMS;
        int last = -1;
        for (int i = b; i < b + GROUP_ITEMS && i < count; i++) {
            const bool head = (i == 0) || cmp_kmer(tile[i].w0, nt) != cmp_kmer(tile[i - 1].w0, nt);
            if (head) last = i;
            headIdx[i] = last;   // -1 = group started in an earlier chunk
        }
        // block inclusive max-scan of `last`
        int v = last;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int nb = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v = max(v, nb); }
        __shared__ int sWarpMax[GROUP_THREADS / 32];
        if (lane == 31) sWarpMax[w] = v;
        __syncthreads();
        int carry = -1;
        for (int ww = 0; ww < w; ww++) carry = max(carry, sWarpMax[ww]);
        const int prevThreads = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        const int before = max(carry, lane > 0 ? prevThreads : -1);   // max head index of all earlier chunks
        for (int i = b; i < b + GROUP_ITEMS && i < count; i++)
            if (headIdx[i] < 0) headIdx[i] = before;
    }
    __syncthreads();
    // each head reduces its in-tile group (groups are short; heavy hitters loop)
    for (int i = tid; i < count; i += GROUP_THREADS) {
        if (headIdx[i] == i) {
            GroupAcc a; a.key = 0; a.strand = 0; a.count = 0;
            const unsigned long long km = cmp_kmer(tile[i].w0, nt);
            int j = i;
            while (j < count && cmp_kmer(tile[j].w0, nt) == km) {
                acc_add(a, rep_key_t<WIDE>(tile[j].w1), (unsigned) (tile[j].w0 >> 63), 1);
                j++;
            }
            if (i == 0) acc_add(a, sBack.key, sBack.strand, sBack.count);
            if (j == count) acc_add(a, sFwd.key, sFwd.strand, sFwd.count);
            gacc[i] = a;
        }
    }
    __syncthreads();
    // emit
    unsigned long long outBase = 0;
    Rec outRec[GROUP_ITEMS];
    unsigned keepMask = 0;
#pragma unroll
    for (int it = 0; it < GROUP_ITEMS; it++) {
        const int i = it * GROUP_THREADS + tid;
        bool keep = false;
        if (i < count) {
            const int h = headIdx[i];
            const GroupAcc a = gacc[h];
            if (a.count >= 2) {
                const Rec r = tile[i];
                unsigned repId, tId; int queryLen, repPos, tLen, tPos;
                if constexpr (WIDE) {
                    const unsigned qi = __ldg(c.byRank + (unsigned) (a.key >> 32)), ti = __ldg(c.byRank + (unsigned) (r.w1 >> 32));
                    repId = __ldg(c.keys + qi); queryLen = (int) __ldg(c.lens + qi) - 2; repPos = (int) (unsigned) (a.key & 0xFFFFFFFFULL);
                    tId = __ldg(c.keys + ti); tLen = (int) __ldg(c.lens + ti) - 2; tPos = (int) (unsigned) (r.w1 & 0xFFFFFFFFULL);
                } else {
                    repId = (unsigned) ((a.key >> 16) & 0xFFFFFFFFULL);
                    queryLen = (int) (0xFFFFULL ^ (a.key >> 48));
                    repPos = (int) (short) (a.key & 0xFFFFULL);
                    tId = (unsigned) (r.w1 >> 32);
                    tLen = (int) (short) ((r.w1 >> 16) & 0xFFFFULL);
                    tPos = (int) (short) (r.w1 & 0xFFFFULL);
                }
                int diagonal = repPos - tPos;
                unsigned qRev = 0;
                if (nt) {
                    // the reference initialises repIsReverse = false for the very first group (kmermatcher.cpp:463)
                    const bool firstGroup = firstKmer ? ((r.w0 & ~(1ULL << 63)) == *firstKmer) : ((h == 0) && sBackAtZero);
                    const bool repIsReverse = firstGroup ? false : (a.strand == 0);
                    const bool targetIsReverse = ((r.w0 >> 63) == 0);
                    int queryPos, targetPos;
                    if (repIsReverse && !targetIsReverse) { queryPos = repPos; targetPos = tPos; qRev = 1; }
                    else if (repIsReverse && targetIsReverse) { queryPos = (queryLen - 1) - repPos; targetPos = (tLen - 1) - tPos; qRev = 0; }
                    else if (!repIsReverse && targetIsReverse) { queryPos = (queryLen - 1) - repPos; targetPos = (tLen - 1) - tPos; qRev = 1; }
                    else { queryPos = repPos; targetPos = tPos; qRev = 0; }
                    if (!WIDE) { queryPos = (short) queryPos; targetPos = (short) targetPos; }   // T = short arithmetic (a no-op for in-range values)
                    diagonal = queryPos - targetPos;
                }
                const bool canBeExtended = diagonal < 0 || (diagonal > (queryLen - tLen));
                keep = c.includeOnlyExtendable == 0 ? can_be_covered(c.covThr, c.covMode, (float) queryLen, (float) tLen) : canBeExtended;
                if (keep) {
                    outRec[it].w0 = ((unsigned long long) repId << 32) | tId;
                    if constexpr (WIDE) {
                        outRec[it].w1 = ((unsigned long long) qRev << 32) | (unsigned long long) ((unsigned) diagonal + 0x80000000u);
                    } else {
                        const unsigned biased = (unsigned) (((int) (short) diagonal) + 32768) & 0xFFFFu;
                        outRec[it].w1 = ((unsigned long long) qRev << 16) | biased;
                    }
                }
            }
        }
        if (keep) keepMask |= 1u << it;
    }
    // block compaction (order inside the output is irrelevant: sort #2 follows)
    const unsigned mine = __popc(keepMask);
    unsigned v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += nb; }
    if (lane == 31) sWarpOut[w] = v;
    __syncthreads();
    unsigned woff = 0, total = 0;
    for (int ww = 0; ww < GROUP_THREADS / 32; ww++) { if (ww < w) woff += sWarpOut[ww]; total += sWarpOut[ww]; }
    if (tid == 0) sOutBase = total ? atomicAdd(outCount, (unsigned long long) total) : 0ULL;
    __syncthreads();
    outBase = sOutBase + woff + v - mine;
#pragma unroll
    for (int it = 0; it < GROUP_ITEMS; it++)
        if (keepMask & (1u << it)) {
            uint4 raw;
            raw.x = (unsigned) outRec[it].w0; raw.y = (unsigned) (outRec[it].w0 >> 32);
            raw.z = (unsigned) outRec[it].w1; raw.w = (unsigned) (outRec[it].w1 >> 32);
            reinterpret_cast<uint4 *>(out)[outBase++] = raw;
        }
}

// ------------------------------------------------------------------------------------------------
// group, fast path: partial-key partition + shared-memory hash join.
// Instead of fully sorting the 64-bit k-mers (8 radix passes) the records are partitioned by B bits of a hash
// of the k-mer (ceil(B/8) passes, B chosen so that a bucket holds ~500 records).  Equal k-mers share a bucket, so
// one CTA per bucket can build the groups in a shared-memory hash table: per k-mer the minimum of the packed
// (seqLen desc, id, pos, strand) key = the representative of assignGroup, and the group size.  A second sweep
// over the bucket (held in registers) emits the (rep, target, diagonal) pairs.  Buckets that do not fit the
// table raise `overflow` and the caller falls back to the full sort + group_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int HG_THREADS = 128;
constexpr int HG_ITEMS_SMALL = 6, HG_TABLE_SMALL = 1024;    // buckets of up to 768 records: 20 KB of shared memory, 8 CTAs per SM
constexpr int HG_ITEMS_BIG = 12, HG_TABLE_BIG = 2048;       // the rare larger ones (deferred to a second launch)
constexpr int HG_MAX_BUCKET = HG_THREADS * HG_ITEMS_BIG;    // 1536 records

__global__ void bucket_bounds_kernel(const Rec *__restrict__ in, unsigned long long n, unsigned long long hashMask, unsigned bucketMask,
                                     unsigned long long *__restrict__ start, unsigned long long *__restrict__ end,
                                     unsigned long long *__restrict__ minKmer) {
    unsigned long long localMin = ~0ULL;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned long long k = in[i].w0 & hashMask;
        const unsigned b = (unsigned) mix64(k) & bucketMask;
        localMin = min(localMin, k);
        if (i == 0) start[b] = 0;
        else {
            const unsigned pb = (unsigned) mix64(in[i - 1].w0 & hashMask) & bucketMask;
            if (pb != b) { start[b] = i; end[pb] = i; }
        }
        if (i == n - 1) end[b] = n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) localMin = min(localMin, __shfl_xor_sync(0xFFFFFFFFu, localMin, o));
    if ((threadIdx.x & 31) == 0) atomicMin(minKmer, localMin);
}

// spill list of the bucketed path: buckets beyond the largest shared-memory instance (a k-mer that occurs in thousands of
// sequences, or an unlucky pile-up) are copied out, fully sorted and grouped by group_kernel -- only these buckets, the
// rest of the iteration keeps the fast path.
__global__ void huge_offsets_kernel(const unsigned *__restrict__ list, unsigned n, const unsigned long long *__restrict__ start,
                                    const unsigned long long *__restrict__ end, unsigned long long *__restrict__ off /* n + 1 */) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long run = 0;
        for (unsigned i = 0; i < n; i++) { off[i] = run; run += end[list[i]] - start[list[i]]; }
        off[n] = run;
    }
}
__global__ void huge_gather_kernel(const Rec *__restrict__ in, const unsigned *__restrict__ list, unsigned n, const unsigned long long *__restrict__ start,
                                   const unsigned long long *__restrict__ end, const unsigned long long *__restrict__ off, Rec *__restrict__ out) {
    for (unsigned i = blockIdx.x; i < n; i += gridDim.x) {
        const unsigned b = list[i];
        const unsigned long long s0 = start[b], cnt = end[b] - s0, o = off[i];
        for (unsigned long long j = threadIdx.x; j < cnt; j += blockDim.x)
            reinterpret_cast<uint4 *>(out)[o + j] = __ldg(reinterpret_cast<const uint4 *>(in) + s0 + j);
    }
}

// packed representative key: (seqLen desc, id asc, pos asc, strand asc); seqLen < 32768 (narrow records)
__device__ __forceinline__ unsigned long long hg_pack(const Rec &r) {
    const unsigned long long id = r.w1 >> 32, len = (r.w1 >> 16) & 0x7FFFULL, pos = r.w1 & 0xFFFFULL;
    return ((0x7FFFULL ^ len) << 49) | (id << 17) | (pos << 1) | (r.w0 >> 63);
}

// pair record of member r in a group whose representative is described by (repId, queryLen, repPos, repStrand)
__device__ __forceinline__ bool make_pair_record(const Rec &r, unsigned repId, int queryLen, int repPos, unsigned repStrand, bool firstGroup,
                                                 const KmConst &c, Rec &out) {
    const unsigned tId = (unsigned) (r.w1 >> 32);
    const int tLen = (int) (short) ((r.w1 >> 16) & 0xFFFFULL);
    const int tPos = (int) (short) (r.w1 & 0xFFFFULL);
    int diagonal = repPos - tPos;
    unsigned qRev = 0;
    if (c.nt) {
        // the reference initialises repIsReverse = false for the very first group (kmermatcher.cpp:463)
        const bool repIsReverse = firstGroup ? false : (repStrand == 0);
        const bool targetIsReverse = ((r.w0 >> 63) == 0);
        int queryPos, targetPos;
        if (repIsReverse && !targetIsReverse) { queryPos = repPos; targetPos = tPos; qRev = 1; }
        else if (repIsReverse && targetIsReverse) { queryPos = (short) ((queryLen - 1) - repPos); targetPos = (short) ((tLen - 1) - tPos); qRev = 0; }
        else if (!repIsReverse && targetIsReverse) { queryPos = (short) ((queryLen - 1) - repPos); targetPos = (short) ((tLen - 1) - tPos); qRev = 1; }
        else { queryPos = repPos; targetPos = tPos; qRev = 0; }
        diagonal = queryPos - targetPos;
    }
    const bool canBeExtended = diagonal < 0 || (diagonal > (queryLen - tLen));
    const bool keep = c.includeOnlyExtendable == 0 ? can_be_covered(c.covThr, c.covMode, (float) queryLen, (float) tLen) : canBeExtended;
    if (keep) {
        out.w0 = ((unsigned long long) repId << 32) | tId;
        const unsigned biased = (unsigned) (((int) (short) diagonal) + 32768) & 0xFFFFu;
        out.w1 = ((unsigned long long) qRev << 16) | biased;
    }
    return keep;
}

// MODE 0: sweep over all buckets, handle those of up to HG_THREADS * ITEMS records and queue the larger ones in bigList;
// MODE 1: process bigList.
template <int TABLE, int ITEMS, int MODE>
__global__ void __launch_bounds__(HG_THREADS) hash_group_kernel(const Rec *__restrict__ in, const unsigned long long *__restrict__ start,
                                                                const unsigned long long *__restrict__ end, unsigned nBuckets,
                                                                unsigned long long hashMask, const unsigned long long *__restrict__ minKmer,
                                                                const KmConst c, Rec *__restrict__ out, unsigned long long *__restrict__ outCount,
                                                                unsigned *__restrict__ bigList, unsigned *__restrict__ bigCount,
                                                                unsigned *__restrict__ hugeList, unsigned *__restrict__ hugeCount) {
    __shared__ unsigned long long sKey[TABLE];
    __shared__ unsigned long long sMin[TABLE];
    __shared__ unsigned sCnt[TABLE];
    __shared__ unsigned sItemWarp[ITEMS][HG_THREADS / 32];   // pairs emitted per (item round, warp), then their exclusive offsets
    __shared__ unsigned long long sOutBase;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const unsigned long long firstKmer = *minKmer;
    const unsigned nWork = MODE == 0 ? nBuckets : *bigCount;
    for (unsigned wi = blockIdx.x; wi < nWork; wi += gridDim.x) {
        const unsigned b = MODE == 0 ? wi : bigList[wi];
        const unsigned long long s0 = start[b], e0 = end[b];
        if (e0 <= s0) continue;
        const unsigned count = (unsigned) (e0 - s0);
        if (count > (unsigned) (HG_THREADS * ITEMS)) {
            // too large for this instance: the next larger one, or the spill list (sorted and grouped apart, only these buckets)
            if (tid == 0) {
                if (MODE == 0 && count <= (unsigned) HG_MAX_BUCKET) bigList[atomicAdd(bigCount, 1u)] = b;
                else hugeList[atomicAdd(hugeCount, 1u)] = b;
            }
            continue;
        }
        // table sized to the bucket (load factor <= 0.5, <= 0.75 for the very largest): clearing it is most of the
        // shared-memory traffic of an average bucket
        unsigned tsize = 64;
        while (tsize < 2 * count && tsize < (unsigned) TABLE) tsize <<= 1;
        const unsigned tmask = tsize - 1;
        for (unsigned i = tid; i < tsize; i += HG_THREADS) { sKey[i] = ~0ULL; sMin[i] = ~0ULL; sCnt[i] = 0; }
        __syncthreads();
        Rec rec[ITEMS];
        unsigned slotOf[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const unsigned i = it * HG_THREADS + tid;
            slotOf[it] = 0xFFFFFFFFu;
            if (i < count) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(in + s0) + i);
                rec[it].w0 = ((unsigned long long) raw.y << 32) | raw.x;
                rec[it].w1 = ((unsigned long long) raw.w << 32) | raw.z;
            }
        }
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const unsigned i = it * HG_THREADS + tid;
            if (i < count) {
                const unsigned long long k = rec[it].w0 & hashMask;
                // aa: the reference drops k-mer == SIZE_T_MAX, its own array sentinel (kmermatcher.cpp:475); ~0 is also the
                // empty marker of the table.  nt keys have 63 bits and never collide with it.
                if (!(hashMask == ~0ULL && k == ~0ULL)) {
                    unsigned slot = (unsigned) (mix64(k) >> 40) & tmask;
                    while (true) {
                        const unsigned long long old = atomicCAS(&sKey[slot], ~0ULL, k);
                        if (old == ~0ULL || old == k) break;
                        slot = (slot + 1) & tmask;
                    }
                    atomicMin(&sMin[slot], hg_pack(rec[it]));
                    atomicAdd(&sCnt[slot], 1u);
                    slotOf[it] = slot;
                }
            }
        }
        __syncthreads();
        // sweep 1: which members produce a pair (the record itself is rebuilt in sweep 2 to keep registers low)
        unsigned keepMask = 0;
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            bool keep = false;
            if (slotOf[it] != 0xFFFFFFFFu && sCnt[slotOf[it]] >= 2) {
                const unsigned long long m = sMin[slotOf[it]];
                Rec tmp;
                keep = make_pair_record(rec[it], (unsigned) ((m >> 17) & 0xFFFFFFFFULL), (int) (0x7FFFULL ^ (m >> 49)), (int) (short) ((m >> 1) & 0xFFFFULL),
                                        (unsigned) (m & 1ULL), (rec[it].w0 & hashMask) == firstKmer, c, tmp);
            }
            if (keep) keepMask |= 1u << it;
            const unsigned bm = __ballot_sync(0xFFFFFFFFu, keep);
            if (lane == 0) sItemWarp[it][w] = __popc(bm);
        }
        __syncthreads();
        // output positions in (item round, warp, lane) order: the 32 records a warp writes per round are contiguous
        if (tid == 0) {
            unsigned run = 0;
#pragma unroll
            for (int it = 0; it < ITEMS; it++)
#pragma unroll
                for (int ww = 0; ww < HG_THREADS / 32; ww++) { const unsigned v = sItemWarp[it][ww]; sItemWarp[it][ww] = run; run += v; }
            sOutBase = run ? atomicAdd(outCount, (unsigned long long) run) : 0ULL;
        }
        __syncthreads();
        const unsigned long long ob = sOutBase;
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const bool keep = (keepMask >> it) & 1u;
            const unsigned bm = __ballot_sync(0xFFFFFFFFu, keep);
            if (keep) {
                const unsigned long long m = sMin[slotOf[it]];
                Rec o;
                make_pair_record(rec[it], (unsigned) ((m >> 17) & 0xFFFFFFFFULL), (int) (0x7FFFULL ^ (m >> 49)), (int) (short) ((m >> 1) & 0xFFFFULL),
                                 (unsigned) (m & 1ULL), (rec[it].w0 & hashMask) == firstKmer, c, o);
                uint4 raw;
                raw.x = (unsigned) o.w0; raw.y = (unsigned) (o.w0 >> 32); raw.z = (unsigned) o.w1; raw.w = (unsigned) (o.w1 >> 32);
                reinterpret_cast<uint4 *>(out)[ob + sItemWarp[it][w] + __popc(bm & ltMask)] = raw;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// reduce: writeKmerMatcherResult (kmermatcher.cpp:809-924) on pairs sorted by (rep, target, diagonal).
// A run starts where rep or target changes; the scan over the run continues while the TARGET id stays
// the same even across a rep boundary (the reference's while-loop at :880 only tests the id).
// ------------------------------------------------------------------------------------------------
// WIDE pair records: 32-bit biased diagonal in w1 bits 0..31, strand in bit 32 (narrow: 16 bits, strand in bit 16)
template <bool WIDE> __device__ __forceinline__ unsigned pair_diag(unsigned long long w1) { return WIDE ? (unsigned) w1 : (unsigned) (w1 & 0xFFFFu); }
template <bool WIDE> __device__ __forceinline__ unsigned pair_rev(unsigned long long w1) { return (unsigned) (w1 >> (WIDE ? 32 : 16)) & 1u; }

template <bool WIDE>
__device__ __forceinline__ bool reduce_one(const Rec *__restrict__ in, unsigned long long n, unsigned long long i, pg_hit &h) {
    const Rec r = in[i];
    if (i > 0 && in[i - 1].w0 == r.w0) return false;            // not the first record of (rep, target)
    const unsigned rep = (unsigned) (r.w0 >> 32), target = (unsigned) r.w0;
    unsigned diagB = pair_diag<WIDE>(r.w1), prevDiag = diagB;
    unsigned best = diagB, bestRev = pair_rev<WIDE>(r.w1);
    unsigned maxDiag = 0, diagCnt = 0, top = 0, revCnt = 0;
    unsigned long long j = i;
    while (j < n) {
        const Rec q = in[j];
        if ((unsigned) q.w0 != target) break;
        const unsigned d = pair_diag<WIDE>(q.w1);
        const unsigned rv = pair_rev<WIDE>(q.w1);
        if (prevDiag == d) { diagCnt++; revCnt += rv; } else { diagCnt = 1; revCnt = rv; }
        // nt: records that agree on (rep, target, diagonal) but not on the strand flag have no defined order in the
        // reference (unstable ips4o sort): it reports the flag of whichever record ends up last.
        // They arise from the first-group quirk of assignGroup (one odd record per target of that group) and from
        // reverse-palindromic repeats.  Here the majority of the winning diagonal decides (a tie -> forward), which is
        // independent of the emission order of the pairs and agrees with the reference unless its odd record is last.
        if (diagCnt >= maxDiag) { best = d; maxDiag = diagCnt; bestRev = (2u * revCnt > diagCnt) ? 1u : 0u; }
        prevDiag = d;
        j++; top++;
    }
    if (target == rep) return false;                             // :899-904
    h.rep = rep; h.target = target;
    h.score = bestRev ? -(int) top : (int) top;
    // hit_t::diagonal is an unsigned short printed as a short (QueryMatcher.h:35-51): the low 16 bits of the diagonal
    h.diag = (int) (short) (unsigned short) (WIDE ? (best - 0x80000000u) : (best - 32768u));
    return true;
}

template <bool WIDE>
__global__ void __launch_bounds__(256) reduce_count_kernel(const Rec *__restrict__ in, unsigned long long n, unsigned *__restrict__ blockCounts) {
    const unsigned long long i = (unsigned long long) blockIdx.x * 256 + threadIdx.x;
    pg_hit h;
    const bool emit = (i < n) && reduce_one<WIDE>(in, n, i, h);
    const unsigned c = __syncthreads_count(emit);
    if (threadIdx.x == 0) blockCounts[blockIdx.x] = c;
}

template <bool WIDE>
__global__ void __launch_bounds__(256) reduce_emit_kernel(const Rec *__restrict__ in, unsigned long long n,
                                                          const unsigned long long *__restrict__ blockOffsets, pg_hit *__restrict__ hits) {
    __shared__ unsigned sWarp[8];
    const unsigned long long i = (unsigned long long) blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    pg_hit h;
    const bool emit = (i < n) && reduce_one<WIDE>(in, n, i, h);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
    if (lane == 0) sWarp[w] = __popc(m);
    __syncthreads();
    unsigned off = 0;
    for (int ww = 0; ww < w; ww++) off += sWarp[ww];
    if (emit) hits[blockOffsets[blockIdx.x] + off + __popc(m & ((1u << lane) - 1u))] = h;
}

// one step of the reference's run scan (kmermatcher.cpp:880-893) with the majority strand rule of reduce_one
__device__ __forceinline__ void run_step(unsigned d, unsigned rv, unsigned &prevDiag, unsigned &diagCnt, unsigned &revCnt,
                                         unsigned &maxDiag, unsigned &best, unsigned &bestRev, unsigned &top) {
    if (prevDiag == d) { diagCnt++; revCnt += rv; } else { diagCnt = 1; revCnt = rv; }
    if (diagCnt >= maxDiag) { best = d; maxDiag = diagCnt; bestRev = (2u * revCnt > diagCnt) ? 1u : 0u; }
    prevDiag = d;
    top++;
}

// ------------------------------------------------------------------------------------------------
// reduce, fast path: sort the pairs by the representative only (ceil(keyBits/8) radix passes), then one WARP per
// representative sorts that representative's pairs by (target, diagonal, strand) in registers (bitonic network over
// warp shuffles, up to 128 pairs) and runs the writeKmerMatcherResult scan on them.  Representatives with more pairs go
// to a CTA-wide shared-memory sort.  The reference's scan does not stop at a change of representative when the next
// block starts with the same target id; the warp therefore peeks at the following representative(s).
// ------------------------------------------------------------------------------------------------
constexpr int SEG_WARP_MAX = 512;     // pairs one warp sorts in registers (16 per lane)
constexpr int SEG_BLOCK_MAX = 16384;   // pairs of one representative sorted by a CTA in shared memory (128 KB)

__global__ void seg_bounds_kernel(const Rec *__restrict__ in, unsigned long long n, unsigned long long *__restrict__ start,
                                  unsigned long long *__restrict__ end, unsigned *__restrict__ minTarget /* preset to 0xFFFFFFFF */) {
    const unsigned long long nRound = (n + 31) & ~31ULL;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < nRound; i += (unsigned long long) gridDim.x * blockDim.x) {
        const bool live = i < n;
        unsigned rep = 0xFFFFFFFFu, target = 0xFFFFFFFFu;
        if (live) {
            const unsigned long long w0 = in[i].w0;
            rep = (unsigned) (w0 >> 32); target = (unsigned) w0;
            if (i == 0) start[rep] = 0;
            else {
                const unsigned prev = (unsigned) (in[i - 1].w0 >> 32);
                if (prev != rep) { start[rep] = i; end[prev] = i; }
            }
            if (i == n - 1) end[rep] = n;
        }
        // smallest target of every representative: the lanes that share a representative combine before the atomic
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, rep);
        const unsigned mt = __reduce_min_sync(peers, target);
        if (live && (threadIdx.x & 31) == (unsigned) (__ffs(peers) - 1)) atomicMin(&minTarget[rep], mt);
    }
}

template <int SLOTS>
__device__ __forceinline__ void warp_bitonic(unsigned long long (&r)[SLOTS], int n2, unsigned lane) {
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                // partner lives in the same lane, another register: js is resolved at compile time so that r[] stays in registers
#pragma unroll
                for (int js = 1; js < SLOTS; js <<= 1) {
                    if (j == js * 32) {
#pragma unroll
                        for (int sl = 0; sl < SLOTS; sl++) {
                            if ((sl & js) == 0) {
                                const bool up = (((sl * 32 + (int) lane) & kk) == 0);
                                const unsigned long long a = r[sl], b = r[sl | js];
                                if ((b < a) == up) { r[sl] = b; r[sl | js] = a; }
                            }
                        }
                    }
                }
            } else {
                const bool lower = ((lane & (unsigned) j) == 0);
#pragma unroll
                for (int sl = 0; sl < SLOTS; sl++) {
                    const unsigned long long p = __shfl_xor_sync(0xFFFFFFFFu, r[sl], j);
                    const bool up = (((sl * 32 + (int) lane) & kk) == 0);
                    r[sl] = (lower == up) ? min(r[sl], p) : max(r[sl], p);
                }
            }
        }
    }
}

struct ScanState { unsigned prevDiag, diagCnt, revCnt, maxDiag, best, bestRev, top; };

// The reference's scan for (rep A, target X) reached the end of A's block.  If the next representative's smallest
// target is X too, the scan goes on through that run (kmermatcher.cpp:880 tests only the target id), possibly over
// several representatives.  Executed by a whole warp; returns the final state in every lane.  *overflow is raised when
// the continued run is longer than the staging list.
__device__ ScanState continue_run_warp(const Rec *__restrict__ in, unsigned long long n, const unsigned long long *__restrict__ end,
                                       const unsigned *__restrict__ minTarget,
                                       unsigned long long nextPos, unsigned X, ScanState st, unsigned *stage /* >= 256 words */,
                                       unsigned *overflow) {
    const unsigned lane = threadIdx.x & 31;
    while (nextPos < n) {
        const unsigned B = (unsigned) (in[nextPos].w0 >> 32);
        if (minTarget[B] != X) break;                 // the usual case: one load decides
        const unsigned long long be = end[B];
        // collect (diagonal, strand) of B's pairs with target X
        unsigned cnt = 0;
        for (unsigned long long i0 = nextPos; i0 < be; i0 += 32) {
            const unsigned long long i = i0 + lane;
            bool hit = false; unsigned v = 0;
            if (i < be) { const Rec r = in[i]; hit = (unsigned) r.w0 == X; v = (unsigned) ((r.w1 & 0xFFFFULL) << 1) | (unsigned) ((r.w1 >> 16) & 1ULL); }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
            const unsigned pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (hit && pos < 256) stage[pos] = v;
            cnt += __popc(m);
        }
        __syncwarp();
        if (cnt > 256) { if (lane == 0) atomicExch(overflow, 1u); break; }
        if (lane == 0) {
            for (unsigned i = 1; i < cnt; i++) { const unsigned v = stage[i]; int j = (int) i - 1; while (j >= 0 && stage[j] > v) { stage[j + 1] = stage[j]; j--; } stage[j + 1] = v; }
            for (unsigned i = 0; i < cnt; i++) run_step(stage[i] >> 1, stage[i] & 1u, st.prevDiag, st.diagCnt, st.revCnt, st.maxDiag, st.best, st.bestRev, st.top);
        }
        st.prevDiag = __shfl_sync(0xFFFFFFFFu, st.prevDiag, 0); st.diagCnt = __shfl_sync(0xFFFFFFFFu, st.diagCnt, 0);
        st.revCnt = __shfl_sync(0xFFFFFFFFu, st.revCnt, 0); st.maxDiag = __shfl_sync(0xFFFFFFFFu, st.maxDiag, 0);
        st.best = __shfl_sync(0xFFFFFFFFu, st.best, 0); st.bestRev = __shfl_sync(0xFFFFFFFFu, st.bestRev, 0); st.top = __shfl_sync(0xFFFFFFFFu, st.top, 0);
        __syncwarp();
        if ((unsigned long long) cnt != be - nextPos) break;      // B has other targets after X: the scan stops inside B
        nextPos = be;                                              // B consisted of X only: look at the next representative
    }
    return st;
}

// MODE 0: sweep over all keys, handle representatives with <= 128 pairs (4 registers per lane: high occupancy), queue the
//         medium ones (<= SEG_WARP_MAX) in midList and the large ones in bigList;   MODE 1: process midList.
template <int MODE>
__global__ void __launch_bounds__(256) reduce_rep_warp_kernel(const Rec *__restrict__ in, unsigned long long n,
                                                              const unsigned long long *__restrict__ start, const unsigned long long *__restrict__ end,
                                                              const unsigned *__restrict__ minTarget,
                                                              unsigned keyLo, unsigned keyHi, pg_hit *__restrict__ tmpHits, unsigned *__restrict__ hitCount,
                                                              unsigned *__restrict__ bigList, unsigned *__restrict__ bigCount,
                                                              unsigned *__restrict__ midList, unsigned *__restrict__ midCount, unsigned *__restrict__ overflow) {
    constexpr int KEYS = MODE == 0 ? 128 : SEG_WARP_MAX;
    __shared__ unsigned long long sKeys[8][KEYS];
    __shared__ unsigned sStage[8][256];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned warpsTotal = gridDim.x * (blockDim.x >> 5);
    // MODE 0 sweeps the representatives this rank owns ([keyLo, keyHi); everything on a single GPU)
    const unsigned nWork = MODE == 0 ? (keyHi - keyLo) : *midCount;
    for (unsigned wi = blockIdx.x * (blockDim.x >> 5) + w; wi < nWork; wi += warpsTotal) {
        const unsigned rep = MODE == 0 ? keyLo + wi : midList[wi];
        const unsigned long long s0 = start[rep], e0 = end[rep];
        if (e0 <= s0) continue;
        const unsigned count = (unsigned) (e0 - s0);
        if (MODE == 0) {
            if (count > SEG_WARP_MAX) { if (lane == 0) bigList[atomicAdd(bigCount, 1u)] = rep; continue; }
            if (count > 128) { if (lane == 0) midList[atomicAdd(midCount, 1u)] = rep; continue; }
        }
        int n2 = 32;
        while (n2 < (int) count) n2 <<= 1;
        // load + sort + stage in shared memory, with as many registers per lane as this representative needs
        auto sortSegment = [&](auto slotsTag) {
            constexpr int SL = decltype(slotsTag)::value;
            unsigned long long r[SL];
#pragma unroll
            for (int sl = 0; sl < SL; sl++) {
                const unsigned e = sl * 32 + lane;
                r[sl] = ~0ULL;
                if (e < count) { const Rec p = in[s0 + e]; r[sl] = ((unsigned long long) (unsigned) p.w0 << 17) | ((p.w1 & 0xFFFFULL) << 1) | ((p.w1 >> 16) & 1ULL); }
            }
            warp_bitonic<SL>(r, SL * 32, lane);
#pragma unroll
            for (int sl = 0; sl < SL; sl++) sKeys[w][sl * 32 + lane] = r[sl];
        };
        if (MODE == 0) {
            if (n2 <= 32) sortSegment(std::integral_constant<int, 1>());
            else if (n2 <= 64) sortSegment(std::integral_constant<int, 2>());
            else sortSegment(std::integral_constant<int, 4>());
        } else {
            if (n2 <= 256) sortSegment(std::integral_constant<int, 8>());
            else sortSegment(std::integral_constant<int, 16>());
        }
        __syncwarp();
        const unsigned long long *key = sKeys[w];
        // run starts (first element of every target), compacted in order; then one lane per run
        unsigned short *runStart = reinterpret_cast<unsigned short *>(sStage[w]);     // up to 512 entries
        unsigned nRuns = 0;
        for (int sl = 0; sl * 32 < (int) count; sl++) {
            const int i = sl * 32 + (int) lane;
            const bool isStart = (i < (int) count) && (i == 0 || (key[i - 1] >> 17) != (key[i] >> 17));
            const unsigned m = __ballot_sync(0xFFFFFFFFu, isStart);
            if (isStart) runStart[nRuns + __popc(m & ((1u << lane) - 1u))] = (unsigned short) i;
            nRuns += __popc(m);
        }
        __syncwarp();
        unsigned nEmitted = 0;
        // the last run of the block (largest target) is the only one that reaches the block end
        bool ownsLast = false; ScanState lastSt; lastSt.prevDiag = lastSt.diagCnt = lastSt.revCnt = lastSt.maxDiag = lastSt.best = lastSt.bestRev = lastSt.top = 0;
        long long lastSlot = -1; unsigned lastTarget = 0;
        for (unsigned r0 = 0; r0 < nRuns; r0 += 32) {
            const unsigned ri = r0 + lane;
            bool emit = false; pg_hit h; h.rep = rep; h.target = 0; h.score = 0; h.diag = 0;
            ScanState st; st.prevDiag = st.diagCnt = st.revCnt = st.maxDiag = st.best = st.bestRev = st.top = 0;
            if (ri < nRuns) {
                const int i = runStart[ri];
                const int e = (ri + 1 < nRuns) ? (int) runStart[ri + 1] : (int) count;
                const unsigned long long k = key[i];
                const unsigned target = (unsigned) (k >> 17);
                st.prevDiag = (unsigned) ((k >> 1) & 0xFFFFULL); st.best = st.prevDiag; st.bestRev = (unsigned) (k & 1ULL);
                for (int j = i; j < e; j++)
                    run_step((unsigned) ((key[j] >> 1) & 0xFFFFULL), (unsigned) (key[j] & 1ULL), st.prevDiag, st.diagCnt, st.revCnt, st.maxDiag, st.best, st.bestRev, st.top);
                emit = target != rep;
                h.target = target;
                h.score = st.bestRev ? -(int) st.top : (int) st.top;
                h.diag = (int) (short) (unsigned short) (st.best - 32768u);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
            const long long slot = (long long) (s0 + nEmitted + __popc(m & ((1u << lane) - 1u)));
            if (emit) tmpHits[slot] = h;
            if (ri + 1 == nRuns) { ownsLast = true; lastSt = st; lastSlot = emit ? slot : -1; lastTarget = h.target; }
            nEmitted += __popc(m);
        }
        __syncwarp();
        // continuation of the last run into the following representative(s)
        const unsigned ownerMask = __ballot_sync(0xFFFFFFFFu, ownsLast);
        if (ownerMask && e0 < n) {
            const int owner = __ffs(ownerMask) - 1;
            ScanState st;
            st.prevDiag = __shfl_sync(0xFFFFFFFFu, lastSt.prevDiag, owner); st.diagCnt = __shfl_sync(0xFFFFFFFFu, lastSt.diagCnt, owner);
            st.revCnt = __shfl_sync(0xFFFFFFFFu, lastSt.revCnt, owner); st.maxDiag = __shfl_sync(0xFFFFFFFFu, lastSt.maxDiag, owner);
            st.best = __shfl_sync(0xFFFFFFFFu, lastSt.best, owner); st.bestRev = __shfl_sync(0xFFFFFFFFu, lastSt.bestRev, owner);
            st.top = __shfl_sync(0xFFFFFFFFu, lastSt.top, owner);
            const unsigned X = __shfl_sync(0xFFFFFFFFu, lastTarget, owner);
            const long long slot = __shfl_sync(0xFFFFFFFFu, lastSlot, owner);
            const unsigned topBefore = st.top;
            st = continue_run_warp(in, n, end, minTarget, e0, X, st, sStage[w], overflow);
            if (st.top != topBefore && slot >= 0 && lane == 0) {
                pg_hit h; h.rep = rep; h.target = X;
                h.score = st.bestRev ? -(int) st.top : (int) st.top;
                h.diag = (int) (short) (unsigned short) (st.best - 32768u);
                tmpHits[slot] = h;
            }
        }
        if (lane == 0) hitCount[rep] = nEmitted;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// reduce, representatives with more than 128 pairs: aggregate before sorting.  A representative with c pairs has only
// D ~ c/8 distinct (target, diagonal) combinations (a pair per shared k-mer), so the pairs are first counted in a
// shared-memory hash table keyed by (target, diagonal) -> (count, reverse-strand count), the D occupied slots are
// compacted and only those are sorted by (target, diagonal) in registers.  The scan of writeKmerMatcherResult then runs
// over (diagonal, count) entries: a diagonal's last update of the `>=` rule happens at its last record, when its count
// is complete, so the aggregated scan selects the same diagonal, strand majority and score.
//   reduce_rep_hashwarp_kernel   one warp per representative of midList (129 .. RH_WARP_MAX pairs)
//   reduce_rep_hashcta_kernel    one CTA per representative of bigList (more); representatives whose table or sorted
//                                list overflows go to hugeList and are done by reduce_rep_block_kernel (full sort)
// ------------------------------------------------------------------------------------------------
constexpr int RH_WARP_MAX = SEG_WARP_MAX;       // pairs per representative in the warp kernel (table of 2 * that)
constexpr int RH_WARP_TABLE = 2 * RH_WARP_MAX;
constexpr int RH_WARPS = 3;                     // warps per CTA of the warp kernel (39 KB of static shared memory)
constexpr int RH_CTA_TABLE = 2048;              // only representatives with <= RH_SORT_MAX distinct entries finish here anyway
constexpr int RH_SORT_MAX = 512;                // aggregated entries one warp sorts in registers
constexpr int RH_PROBE_MAX = 64;

// entry = target << 32 | diagonal << 16 | table slot
__device__ __forceinline__ void agg_step(unsigned d, unsigned cnt, unsigned rv, ScanState &st) {
    st.prevDiag = d; st.diagCnt = cnt; st.revCnt = rv;
    if (cnt >= st.maxDiag) { st.best = d; st.maxDiag = cnt; st.bestRev = (2u * rv > cnt) ? 1u : 0u; }
    st.top += cnt;
}

// insert pairs in[s0 + first], in[s0 + first + step], ... into the table; returns false if a probe sequence gave up
__device__ __forceinline__ bool rh_insert(const Rec *__restrict__ in, unsigned long long s0, unsigned count, unsigned first, unsigned step,
                                          unsigned long long *sKey, unsigned *sCnt, unsigned tmask) {
    bool okAll = true;
    // four independent loads in flight per lane: the insert loop is otherwise one exposed DRAM latency per record
    for (unsigned i0 = first; i0 < count; i0 += 4 * step) {
        uint4 raws[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned i = i0 + u * step;
            if (i < count) raws[u] = __ldg(reinterpret_cast<const uint4 *>(in + s0) + i);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned i = i0 + u * step;
            if (i >= count) break;
            const uint4 raw = raws[u];
            const unsigned long long key = ((unsigned long long) raw.x << 16) | (unsigned long long) (raw.z & 0xFFFFu);   // target, biased diagonal
            const unsigned strand = (raw.z >> 16) & 1u;
            unsigned slot = (unsigned) (mix64(key) >> 40) & tmask;
            bool placed = false;
            for (int probe = 0; probe < RH_PROBE_MAX; probe++) {
                const unsigned long long old = atomicCAS(&sKey[slot], ~0ULL, key);
                if (old == ~0ULL || old == key) { placed = true; break; }
                slot = (slot + 1) & tmask;
            }
            if (placed) atomicAdd(&sCnt[slot], 1u + (strand << 16));
            else okAll = false;
        }
    }
    return okAll;
}

// One warp: sorts the D <= RH_SORT_MAX entries of ent[] (shared memory), scans the (rep, target) runs, writes the hit
// lines of `rep` to tmpHits[s0 ...] and hitCount[rep]; continues the last run into the following representatives.
__device__ void rh_sort_scan_emit(const Rec *__restrict__ in, unsigned long long n, const unsigned long long *__restrict__ end,
                                  const unsigned *__restrict__ minTarget, unsigned rep, unsigned long long s0, unsigned long long e0,
                                  unsigned long long *ent, unsigned D, const unsigned *sCnt, unsigned *stage /* >= 256 words */,
                                  pg_hit *__restrict__ tmpHits, unsigned *__restrict__ hitCount, unsigned *overflow) {
    const unsigned lane = threadIdx.x & 31;
    int n2 = 32;
    while (n2 < (int) D) n2 <<= 1;
    auto sortEntries = [&](auto slotsTag) {
        constexpr int SL = decltype(slotsTag)::value;
        unsigned long long r[SL];
#pragma unroll
        for (int sl = 0; sl < SL; sl++) { const unsigned e = sl * 32 + lane; r[sl] = e < D ? ent[e] : ~0ULL; }
        __syncwarp();
        warp_bitonic<SL>(r, SL * 32, lane);
#pragma unroll
        for (int sl = 0; sl < SL; sl++) ent[sl * 32 + lane] = r[sl];
    };
    if (n2 <= 32) sortEntries(std::integral_constant<int, 1>());
    else if (n2 <= 64) sortEntries(std::integral_constant<int, 2>());
    else if (n2 <= 128) sortEntries(std::integral_constant<int, 4>());
    else if (n2 <= 256) sortEntries(std::integral_constant<int, 8>());
    else sortEntries(std::integral_constant<int, 16>());
    __syncwarp();
    // run starts (first entry of every target), compacted in order; then one lane per run
    unsigned short *runStart = reinterpret_cast<unsigned short *>(stage);     // up to 512 entries
    unsigned nRuns = 0;
    for (int sl = 0; sl * 32 < (int) D; sl++) {
        const int i = sl * 32 + (int) lane;
        const bool isStart = (i < (int) D) && (i == 0 || (ent[i - 1] >> 32) != (ent[i] >> 32));
        const unsigned m = __ballot_sync(0xFFFFFFFFu, isStart);
        if (isStart) runStart[nRuns + __popc(m & ((1u << lane) - 1u))] = (unsigned short) i;
        nRuns += __popc(m);
    }
    __syncwarp();
    unsigned nEmitted = 0;
    bool ownsLast = false; ScanState lastSt; lastSt.prevDiag = lastSt.diagCnt = lastSt.revCnt = lastSt.maxDiag = lastSt.best = lastSt.bestRev = lastSt.top = 0;
    long long lastSlot = -1; unsigned lastTarget = 0;
    for (unsigned r0 = 0; r0 < nRuns; r0 += 32) {
        const unsigned ri = r0 + lane;
        bool emit = false; pg_hit h; h.rep = rep; h.target = 0; h.score = 0; h.diag = 0;
        ScanState st; st.prevDiag = st.diagCnt = st.revCnt = st.maxDiag = st.best = st.bestRev = st.top = 0;
        if (ri < nRuns) {
            const int i = runStart[ri];
            const int e = (ri + 1 < nRuns) ? (int) runStart[ri + 1] : (int) D;
            const unsigned target = (unsigned) (ent[i] >> 32);
            for (int j = i; j < e; j++) {
                const unsigned long long k = ent[j];
                const unsigned cv = sCnt[(unsigned) (k & 0xFFFFULL)];
                agg_step((unsigned) ((k >> 16) & 0xFFFFULL), cv & 0xFFFFu, cv >> 16, st);
            }
            emit = target != rep;
            h.target = target;
            h.score = st.bestRev ? -(int) st.top : (int) st.top;
            h.diag = (int) (short) (unsigned short) (st.best - 32768u);
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
        const long long slot = (long long) (s0 + nEmitted + __popc(m & ((1u << lane) - 1u)));
        if (emit) tmpHits[slot] = h;
        if (ri + 1 == nRuns) { ownsLast = true; lastSt = st; lastSlot = emit ? slot : -1; lastTarget = h.target; }
        nEmitted += __popc(m);
    }
    __syncwarp();
    // continuation of the last run into the following representative(s)
    const unsigned ownerMask = __ballot_sync(0xFFFFFFFFu, ownsLast);
    if (ownerMask && e0 < n) {
        const int owner = __ffs(ownerMask) - 1;
        ScanState st;
        st.prevDiag = __shfl_sync(0xFFFFFFFFu, lastSt.prevDiag, owner); st.diagCnt = __shfl_sync(0xFFFFFFFFu, lastSt.diagCnt, owner);
        st.revCnt = __shfl_sync(0xFFFFFFFFu, lastSt.revCnt, owner); st.maxDiag = __shfl_sync(0xFFFFFFFFu, lastSt.maxDiag, owner);
        st.best = __shfl_sync(0xFFFFFFFFu, lastSt.best, owner); st.bestRev = __shfl_sync(0xFFFFFFFFu, lastSt.bestRev, owner);
        st.top = __shfl_sync(0xFFFFFFFFu, lastSt.top, owner);
        const unsigned X = __shfl_sync(0xFFFFFFFFu, lastTarget, owner);
        const long long slot = __shfl_sync(0xFFFFFFFFu, lastSlot, owner);
        const unsigned topBefore = st.top;
        st = continue_run_warp(in, n, end, minTarget, e0, X, st, stage, overflow);
        if (st.top != topBefore && slot >= 0 && lane == 0) {
            pg_hit h; h.rep = rep; h.target = X;
            h.score = st.bestRev ? -(int) st.top : (int) st.top;
            h.diag = (int) (short) (unsigned short) (st.best - 32768u);
            tmpHits[slot] = h;
        }
    }
    if (lane == 0) hitCount[rep] = nEmitted;
    __syncwarp();
}

__global__ void __launch_bounds__(RH_WARPS * 32) reduce_rep_hashwarp_kernel(const Rec *__restrict__ in, unsigned long long n,
                                                                            const unsigned long long *__restrict__ start, const unsigned long long *__restrict__ end,
                                                                            const unsigned *__restrict__ minTarget,
                                                                            const unsigned *__restrict__ midList, const unsigned *__restrict__ midCount,
                                                                            pg_hit *__restrict__ tmpHits, unsigned *__restrict__ hitCount, unsigned *__restrict__ overflow) {
    // per warp: table keys (the compacted entries are written over them, in place), counts, staging
    __shared__ unsigned long long sKey[RH_WARPS][RH_WARP_TABLE];
    __shared__ unsigned sCnt[RH_WARPS][RH_WARP_TABLE];
    __shared__ unsigned sStage[RH_WARPS][256];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const unsigned warpsTotal = gridDim.x * RH_WARPS;
    const unsigned nWork = *midCount;
    for (unsigned wi = blockIdx.x * RH_WARPS + w; wi < nWork; wi += warpsTotal) {
        const unsigned rep = midList[wi];
        const unsigned long long s0 = start[rep], e0 = end[rep];
        const unsigned count = (unsigned) (e0 - s0);           // 129 .. RH_WARP_MAX
        unsigned tsize = 256;
        while (tsize < 2 * count) tsize <<= 1;                 // load factor <= 0.5: the probes always end
        for (unsigned i = lane; i < tsize; i += 32) { sKey[w][i] = ~0ULL; sCnt[w][i] = 0; }
        __syncwarp();
        rh_insert(in, s0, count, lane, 32, sKey[w], sCnt[w], tsize - 1);
        __syncwarp();
        // in-place compaction: an entry moves to an index <= its slot, and every chunk is read before it is written
        unsigned D = 0;
        for (unsigned base = 0; base < tsize; base += 32) {
            const unsigned long long key = sKey[w][base + lane];
            const bool occ = key != ~0ULL;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, occ);
            __syncwarp();
            if (occ) sKey[w][D + __popc(m & ltMask)] = ((key >> 16) << 32) | ((key & 0xFFFFULL) << 16) | (unsigned long long) (base + lane);
            D += __popc(m);
            __syncwarp();
        }
        rh_sort_scan_emit(in, n, end, minTarget, rep, s0, e0, sKey[w], D, sCnt[w], sStage[w], tmpHits, hitCount, overflow);
    }
}

__global__ void __launch_bounds__(256) reduce_rep_hashcta_kernel(const Rec *__restrict__ in, unsigned long long n,
                                                                 const unsigned long long *__restrict__ start, const unsigned long long *__restrict__ end,
                                                                 const unsigned *__restrict__ minTarget,
                                                                 const unsigned *__restrict__ bigList, const unsigned *__restrict__ bigCount,
                                                                 unsigned *__restrict__ hugeList, unsigned *__restrict__ hugeCount,
                                                                 pg_hit *__restrict__ tmpHits, unsigned *__restrict__ hitCount, unsigned *__restrict__ overflow) {
    __shared__ unsigned long long sKey[RH_CTA_TABLE];
    __shared__ unsigned sCnt[RH_CTA_TABLE];
    __shared__ unsigned long long sEnt[RH_SORT_MAX];
    __shared__ unsigned sStage[256];
    __shared__ unsigned sD, sFail;
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const unsigned nBig = *bigCount;
    for (unsigned bi = blockIdx.x; bi < nBig; bi += gridDim.x) {
        const unsigned rep = bigList[bi];
        const unsigned long long s0 = start[rep], e0 = end[rep];
        const unsigned long long count64 = e0 - s0;
        unsigned tsize = 1024;
        while (tsize < 2 * count64 && tsize < (unsigned) RH_CTA_TABLE) tsize <<= 1;
        for (unsigned i = tid; i < tsize; i += 256) { sKey[i] = ~0ULL; sCnt[i] = 0; }
        if (tid == 0) { sD = 0; sFail = count64 > 0xFFFFULL ? 1u : 0u; }     // counts are 16-bit fields
        __syncthreads();
        if (!sFail && !rh_insert(in, s0, (unsigned) count64, tid, 256, sKey, sCnt, tsize - 1)) sFail = 1;
        __syncthreads();
        if (!sFail) {
            for (unsigned base = w * 32; base < tsize; base += 256) {
                const unsigned long long key = sKey[base + lane];
                const bool occ = key != ~0ULL;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, occ);
                unsigned pos0 = 0;
                if (lane == 0 && m) pos0 = atomicAdd(&sD, (unsigned) __popc(m));
                pos0 = __shfl_sync(0xFFFFFFFFu, pos0, 0);
                const unsigned pos = pos0 + __popc(m & ltMask);
                if (occ && pos < (unsigned) RH_SORT_MAX) sEnt[pos] = ((key >> 16) << 32) | ((key & 0xFFFFULL) << 16) | (unsigned long long) (base + lane);
            }
        }
        __syncthreads();
        const bool fail = sFail || sD > (unsigned) RH_SORT_MAX;
        if (fail) {
            if (tid == 0) hugeList[atomicAdd(hugeCount, 1u)] = rep;
        } else if (w == 0) {
            rh_sort_scan_emit(in, n, end, minTarget, rep, s0, e0, sEnt, sD, sCnt, sStage, tmpHits, hitCount, overflow);
        }
        __syncthreads();
    }
}

// representatives with more than SEG_WARP_MAX pairs: one CTA each, bitonic sort in shared memory
__global__ void __launch_bounds__(256) reduce_rep_block_kernel(const Rec *__restrict__ in, unsigned long long n,
                                                               const unsigned long long *__restrict__ start, const unsigned long long *__restrict__ end,
                                                               const unsigned *__restrict__ minTarget,
                                                               const unsigned *__restrict__ bigList, const unsigned *__restrict__ bigCount,
                                                               pg_hit *__restrict__ tmpHits, unsigned *__restrict__ hitCount, unsigned *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char rb_smem[];
    unsigned long long *key = reinterpret_cast<unsigned long long *>(rb_smem);
    __shared__ unsigned sWarp[8];
    __shared__ unsigned sStage[256];
    __shared__ ScanState sLast;
    __shared__ long long sLastSlot;
    __shared__ unsigned sLastTarget, sHasLast;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned nBig = *bigCount;
    for (unsigned bi = blockIdx.x; bi < nBig; bi += gridDim.x) {
        const unsigned rep = bigList[bi];
        const unsigned long long s0 = start[rep], e0 = end[rep];
        const unsigned count = (unsigned) (e0 - s0);
        if (count > SEG_BLOCK_MAX) { if (tid == 0) { atomicExch(overflow, 1u); hitCount[rep] = 0; } continue; }
        int n2 = 1;
        while (n2 < (int) count) n2 <<= 1;
        for (int i = tid; i < n2; i += 256) {
            unsigned long long k = ~0ULL;
            if (i < (int) count) { const Rec p = in[s0 + i]; k = ((unsigned long long) (unsigned) p.w0 << 17) | ((p.w1 & 0xFFFFULL) << 1) | ((p.w1 >> 16) & 1ULL); }
            key[i] = k;
        }
        if (tid == 0) sHasLast = 0;
        __syncthreads();
        for (int kk = 2; kk <= n2; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (n2 >> 1); t += 256) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int ix = i | j;
                    const bool up = ((i & kk) == 0);
                    const unsigned long long a = key[i], c2 = key[ix];
                    if ((c2 < a) == up) { key[i] = c2; key[ix] = a; }
                }
                __syncthreads();
            }
        }
        unsigned long long base = 0;
        for (int i0 = 0; i0 < (int) count; i0 += 256) {
            const int i = i0 + tid;
            bool emit = false; pg_hit h; h.rep = rep; h.target = 0; h.score = 0; h.diag = 0;
            bool reached = false; ScanState st; st.prevDiag = st.diagCnt = st.revCnt = st.maxDiag = st.best = st.bestRev = st.top = 0;
            if (i < (int) count) {
                const unsigned long long k = key[i];
                if (i == 0 || (key[i - 1] >> 17) != (k >> 17)) {
                    const unsigned target = (unsigned) (k >> 17);
                    st.prevDiag = (unsigned) ((k >> 1) & 0xFFFFULL); st.best = st.prevDiag; st.bestRev = (unsigned) (k & 1ULL);
                    int j = i;
                    while (j < (int) count && (unsigned) (key[j] >> 17) == target) {
                        run_step((unsigned) ((key[j] >> 1) & 0xFFFFULL), (unsigned) (key[j] & 1ULL), st.prevDiag, st.diagCnt, st.revCnt, st.maxDiag, st.best, st.bestRev, st.top);
                        j++;
                    }
                    emit = target != rep;
                    h.target = target;
                    h.score = st.bestRev ? -(int) st.top : (int) st.top;
                    h.diag = (int) (short) (unsigned short) (st.best - 32768u);
                    reached = (j == (int) count);
                }
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
            if (lane == 0) sWarp[w] = __popc(m);
            __syncthreads();
            unsigned off = 0, total = 0;
            for (int ww = 0; ww < 8; ww++) { if (ww < w) off += sWarp[ww]; total += sWarp[ww]; }
            const long long slot = (long long) (s0 + base + off + __popc(m & ((1u << lane) - 1u)));
            if (emit) tmpHits[slot] = h;
            if (reached) { sLast = st; sLastSlot = emit ? slot : -1; sLastTarget = h.target; sHasLast = 1; }
            base += total;
            __syncthreads();
        }
        if (w == 0 && sHasLast && e0 < n) {
            ScanState st = sLast;
            const unsigned topBefore = st.top;
            st = continue_run_warp(in, n, end, minTarget, e0, sLastTarget, st, sStage, overflow);
            if (st.top != topBefore && sLastSlot >= 0 && lane == 0) {
                pg_hit h; h.rep = rep; h.target = sLastTarget;
                h.score = st.bestRev ? -(int) st.top : (int) st.top;
                h.diag = (int) (short) (unsigned short) (st.best - 32768u);
                tmpHits[sLastSlot] = h;
            }
        }
        if (tid == 0) hitCount[rep] = (unsigned) base;
        __syncthreads();
    }
}

__global__ void compact_rep_hits_kernel(const pg_hit *__restrict__ tmpHits, const unsigned long long *__restrict__ start,
                                        const unsigned *__restrict__ hitCount, const unsigned long long *__restrict__ hitOffset,
                                        unsigned keyLo, unsigned keyHi, pg_hit *__restrict__ hits) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (unsigned rep = keyLo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rep < keyHi; rep += warpsTotal) {
        const unsigned c = hitCount[rep];
        if (c == 0) continue;
        const unsigned long long s0 = start[rep], o = hitOffset[rep];
        for (unsigned i = lane; i < c; i += 32) hits[o + i] = tmpHits[s0 + i];
    }
}

// exclusive scan of per-block counts (single block, sequential chunks; counts <= 2^24 blocks)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const unsigned *__restrict__ counts, unsigned long long nBlocks,
                                                           unsigned long long *__restrict__ offsets, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long sWarp[32];
    __shared__ unsigned long long sCarry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) sCarry = 0;
    __syncthreads();
    for (unsigned long long b = 0; b < nBlocks; b += 1024) {
        const unsigned long long i = b + tid;
        const unsigned long long c = (i < nBlocks) ? counts[i] : 0ULL;
        unsigned long long v = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long nb = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += nb; }
        if (lane == 31) sWarp[w] = v;
        __syncthreads();
        unsigned long long woff = 0;
        for (int ww = 0; ww < w; ww++) woff += sWarp[ww];
        const unsigned long long carry = sCarry;
        if (i < nBlocks) offsets[i] = carry + woff + v - c;
        __syncthreads();
        if (tid == 1023) sCarry = carry + woff + v;
        __syncthreads();
    }
    if (tid == 0) *total = sCarry;
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static int bits_for(unsigned maxValue) {
    int b = 1;
    while (b < 32 && (maxValue >> b) != 0) b++;
    return b;
}

int km_setup_constants(const pg_seqdb *db, const pg_km_params *p, KmConst &c, cudaStream_t stream) {
    const bool nt = db->dbtype == PG_DBTYPE_NUCLEOTIDES;
    PG_CHECK(p->kmer_size >= 2 && p->kmer_size <= (nt ? 31 : 14), "kmermatcher: unsupported k (aa: 2..14 in base-(alph-1) < 2^63, nt: 2..31)");
    PG_CHECK(nt || p->alph_size == 13 || p->alph_size == 21, "kmermatcher: --alph-size must be 13 or 21 for amino acids");
    const unsigned char *tab = nt ? PG_NT_AA2NUM : (p->alph_size == 21 ? PG_AA_AA2NUM : PG_RED_AA2NUM);
    PG_CUDA(cudaMemcpyToSymbolAsync(c_aa2num, tab, 256, 0, cudaMemcpyHostToDevice, stream));
    c.k = p->kmer_size; c.nt = nt ? 1 : 0;
    c.xCode = nt ? 4 : p->alph_size - 1;
    c.base = nt ? 4u : (unsigned) (p->alph_size - 1);
    c.seed = (unsigned long long) p->hash_shift;
    c.kmersPerSeq = p->kmers_per_seq; c.scale = p->kmers_per_seq_scale;
    c.ignoreMulti = p->ignore_multi_kmer;
    c.hashStart = p->hash_start; c.hashEnd = p->hash_end;
    c.includeOnlyExtendable = p->include_only_extendable;
    c.covMode = p->cov_mode; c.covThr = p->cov_thr;
    c.wide = 0; c.rankOf = nullptr; c.byRank = nullptr; c.keys = db->keys; c.lens = db->lens;
    return 0;
}

// kmermatcher.cpp:797-802: KmerPosition<short> iff DBReader::getMaxSeqLen() < SHRT_MAX, and getMaxSeqLen is the longest
// ENTRY (residues + "\n\0", DBReader.cpp:811), i.e. seqLen + 2.
static bool km_is_wide(const pg_seqdb *db) { return db->max_seq_len + 2 >= 32767u; }

// rank tables of the wide layout: sequences ordered by (seqLen desc, key asc)
static __global__ void wide_rank_keys_kernel(const unsigned *__restrict__ lens, const unsigned *__restrict__ keys, unsigned n, Rec *__restrict__ out) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Rec r;
        r.w0 = ((unsigned long long) (0xFFFFFFFFu - (lens[i] - 2u)) << 32) | keys[i];
        r.w1 = i;
        out[i] = r;
    }
}
static __global__ void wide_rank_tables_kernel(const Rec *__restrict__ sorted, unsigned n, unsigned *__restrict__ rankOf, unsigned *__restrict__ byRank) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned si = (unsigned) sorted[i].w1;
        byRank[i] = si;
        rankOf[si] = i;
    }
}

static int km_prepare_wide(Context *ctx, const pg_seqdb *db, KmConst &c) {
    if (!km_is_wide(db)) return 0;
    cudaStream_t s = ctx->stream;
    const unsigned n = (unsigned) db->n;
    PG_CHECK(ctx->seqLo == 0 && ctx->seqHi >= n && ctx->ownLo == 0 && ctx->ownHi == 0xFFFFFFFFu,
             "kmermatcher: the wide (T=int) record layout is single-GPU only for now");
    PG_TRY(ctx->wideTabs.reserve(sizeof(unsigned) * 2 * ((size_t) n + 1)));
    PG_TRY(ctx->recA.reserve(sizeof(Rec) * ((size_t) n + 1)));
    PG_TRY(ctx->recB.reserve(sizeof(Rec) * ((size_t) n + 1)));
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(n)));
    unsigned *rankOf = ctx->wideTabs.as<unsigned>(), *byRank = rankOf + n + 1;
    wide_rank_keys_kernel<<<NUM_SMS * 4, 256, 0, s>>>(db->lens, db->keys, n, ctx->recA.as<Rec>());
    RadixPlan plan; plan.npasses = 0;
    plan_add_bits(plan, 0, 0, 64);
    Rec *sorted = nullptr;
    PG_TRY(radix_sort(ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), n, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches));
    wide_rank_tables_kernel<<<NUM_SMS * 4, 256, 0, s>>>(sorted, n, rankOf, byRank);
    ctx->launches += 2;
    PG_CUDA(cudaStreamSynchronize(s));      // recA / recB may be re-allocated by the extraction that follows
    PG_CUDA(cudaGetLastError());
    c.wide = 1; c.rankOf = rankOf; c.byRank = byRank;
    return 0;
}

// computeKmerCount (kmermatcher.cpp:576-585): upper bound on emitted records
static __global__ void kmer_count_kernel(const unsigned *__restrict__ lens, unsigned lo, unsigned n, int k, int kps, float scale,
                                         unsigned long long *__restrict__ total) {
    unsigned long long mine = 0;
    for (unsigned i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int seqLen = (int) lens[i] - 2;
        const int adj = max(1, seqLen - k + 2);
        mine += (unsigned long long) min(adj, (int) ((float) kps + (scale * (float) seqLen)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

template <int NMAX, int KT, int NTM>
static int launch_extract_warp_t(const pg_seqdb &db, const unsigned *list, const unsigned *listCount, unsigned hostCount, const KmConst &c,
                                 Rec *out, unsigned long long *outCount, unsigned long long outCap, unsigned long long *partHist,
                                 cudaStream_t stream, uint64_t *launches) {
    const size_t smem = 4 * ((size_t) NMAX * sizeof(PCand) + (size_t) (NMAX + 1) * sizeof(Rec) + (NMAX + 40));
    PG_CUDA(cudaFuncSetAttribute(extract_warp_kernel<NMAX, KT, NTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    unsigned blocks = (hostCount + 3) / 4;
    const unsigned maxBlocks = NUM_SMS * 32;
    if (blocks > maxBlocks) blocks = maxBlocks;
    extract_warp_kernel<NMAX, KT, NTM><<<blocks, 128, smem, stream>>>(db, list, listCount, c, out, outCount, outCap, partHist);
    if (launches) (*launches)++;
    return 0;
}

// the workflow defaults (aa k = 14, nt k = 22) get fully unrolled instances, anything else the generic one
template <int NMAX>
static int launch_extract_warp(const pg_seqdb &db, const unsigned *list, const unsigned *listCount, unsigned hostCount, const KmConst &c,
                               Rec *out, unsigned long long *outCount, unsigned long long outCap, unsigned long long *partHist,
                               cudaStream_t stream, uint64_t *launches) {
    if (hostCount == 0) return 0;
    if (!c.nt && c.k == 14) return launch_extract_warp_t<NMAX, 14, 0>(db, list, listCount, hostCount, c, out, outCount, outCap, partHist, stream, launches);
    if (c.nt && c.k == 22) return launch_extract_warp_t<NMAX, 22, 1>(db, list, listCount, hostCount, c, out, outCount, outCap, partHist, stream, launches);
    return launch_extract_warp_t<NMAX, 0, -1>(db, list, listCount, hostCount, c, out, outCount, outCap, partHist, stream, launches);
}

// Stage 1: extraction.  Leaves the records in ws.recA, returns their count.
int km_extract(Context *ctx, const pg_seqdb *db, const pg_km_params *p, const KmConst &c, uint64_t *nRecords) {
    cudaStream_t s = ctx->stream;
    const unsigned n = (unsigned) db->n;
    PG_CHECK(c.ignoreMulti || db->max_seq_len - c.k + 1 <= 1024, "kmermatcher: --ignore-multi-kmer 0 with sequences of more than 1024 k-mers is not supported");
    PG_TRY(ctx->small.reserve(4096));
    unsigned long long *d_total = ctx->small.as<unsigned long long>();        // [0] capacity estimate
    unsigned long long *d_outCount = d_total + 1;                             // [1] emitted records
    unsigned *d_clsCount = (unsigned *) (d_total + 8);                        // [8..] 4 class counters
    PG_CUDA(cudaMemsetAsync(d_total, 0, 256, s));
    const unsigned sLo = std::min(ctx->seqLo, n), sHi = std::min(ctx->seqHi, n);
    const unsigned long long hint = (sLo == 0 && sHi == n) ? ctx->kmerTotalHint : 0;     // km_choose_splits counted already
    ctx->kmerTotalHint = 0;
    if (!hint) kmer_count_kernel<<<NUM_SMS * 4, 256, 0, s>>>(db->lens, sLo, sHi, c.k, c.kmersPerSeq, c.scale, d_total);
    PG_TRY(ctx->lists.reserve(sizeof(unsigned) * 4 * (size_t) n + 16));
    unsigned *lists = ctx->lists.as<unsigned>();
    classify_kernel<<<(sHi - sLo + 255) / 256 + 1, 256, 0, s>>>(db->lens, sLo, sHi, n, c.k, lists, d_clsCount);
    ctx->launches += 2;
    // small read-backs go through mapped pinned memory (pg::read_back), not the copy engine: a cudaMemcpyAsync would
    // queue behind the previous iteration's result transfers that are still running on the copy stream
    unsigned long long h_total = 0; unsigned h_cls[4];
    {
        unsigned long long hb[10];                     // [0] capacity estimate, [8..9] the four class counters
        PG_TRY(read_back(ctx, hb, d_total, sizeof(hb)));
        h_total = hint ? hint - 1 : hb[0];
        memcpy(h_cls, hb + 8, sizeof(h_cls));
    }
    // a hash-range split holds 1 / splitDiv of the records (XXH64 is uniform) + 12.5 % + slack; if a skewed input
    // overflows that, the caller doubles the number of splits (rc 2)
    const unsigned long long cap = ctx->splitDiv > 1 ? h_total / ctx->splitDiv + h_total / (8ull * ctx->splitDiv) + 65536ull : h_total + 1;
    PG_TRY(ctx->recA.reserve(sizeof(Rec) * cap));
    if (!ctx->extractOnly) PG_TRY(ctx->recB.reserve(sizeof(Rec) * cap));       // the sort's second buffer (not for the fused multi-GPU exchange)
    Rec *out = ctx->recA.as<Rec>();
    // histograms of the partition digits, counted while the records leave shared memory (consumed by km_group_bucketed)
    ctx->preHistValid = false;
    PG_TRY(ctx->preHist.reserve(sizeof(unsigned long long) * EXTRACT_HIST_BINS));
    unsigned long long *d_partHist = ctx->preHist.as<unsigned long long>();
    PG_CUDA(cudaMemsetAsync(d_partHist, 0, sizeof(unsigned long long) * EXTRACT_HIST_BINS, s));
    PG_TRY(launch_extract_warp<64>(*db, lists + 0 * (size_t) n, d_clsCount + 0, h_cls[0], c, out, d_outCount, cap, d_partHist, s, &ctx->launches));
    PG_TRY(launch_extract_warp<256>(*db, lists + 1 * (size_t) n, d_clsCount + 1, h_cls[1], c, out, d_outCount, cap, d_partHist, s, &ctx->launches));
    PG_TRY(launch_extract_warp<1024>(*db, lists + 2 * (size_t) n, d_clsCount + 2, h_cls[2], c, out, d_outCount, cap, d_partHist, s, &ctx->launches));
    if (h_cls[3]) {
        const unsigned maxL = db->max_seq_len;
        size_t n2 = 1;
        const size_t maxCand = (size_t) maxL + 1;   // worst case: every window shares one score (low-complexity sequence)
        while (n2 < maxCand) n2 <<= 1;
        const size_t perBlock = (((size_t) maxL + 32) & ~(size_t) 15) + n2 * sizeof(Cand) + (maxCand + 2) * sizeof(Rec);
        unsigned blocks = h_cls[3] < (unsigned) NUM_SMS * 2 ? h_cls[3] : NUM_SMS * 2;
        PG_TRY(ctx->scratch.reserve(perBlock * blocks));
        extract_block_kernel<<<blocks, 256, 0, s>>>(*db, lists + 3 * (size_t) n, d_clsCount + 3, c, out, d_outCount, cap,
                                                    ctx->scratch.as<unsigned char>(), perBlock, d_partHist);
        ctx->launches++;
    }
    unsigned long long h_out = 0;
    PG_TRY(read_back(ctx, &h_out, d_outCount, sizeof(h_out)));
    PG_CUDA(cudaGetLastError());
    if (h_out > cap && ctx->splitDiv > 1) { *nRecords = h_out; return 2; }
    PG_CHECK(h_out <= cap, "kmermatcher: k-mer array overflow");
    *nRecords = h_out;
    ctx->preHistValid = true; ctx->preHistRecords = h_out;
    (void) p;
    return 0;
}

// Stage 2: sort #1 + group.  Input records in recA (n), output pair records in recA (count returned).
// fast path of stage 2 (see hash_group_kernel); returns *ok = false if a bucket overflowed the shared-memory table
static int km_group_bucketed(Context *ctx, const KmConst &c, uint64_t nRecords, uint64_t *nPairs, bool *ok, bool preHist) {
    cudaStream_t s = ctx->stream;
    *ok = false;
    int B = 1;
    while (B < 24 && (nRecords >> B) > ctx->bucketTarget) B++;          // ~512 records per bucket on average
    if ((nRecords >> B) > ctx->bucketTarget + ctx->bucketTarget / 3) return 0;                  // more than 2^24 buckets would be needed: use the full sort
    const bool bigFirst = ctx->bucketTarget > 560;        // average bucket beyond the small instance's capacity (768 records)
    const unsigned nBuckets = 1u << B;
    const unsigned long long hashMask = c.nt ? ~(1ULL << 63) : ~0ULL;   // nt: bit 63 is the strand flag (kmermatcher.h:77-96)
    RadixPlan plan; plan.npasses = 0;
    if (ctx->digitBits > 8) plan_add_hash_bits_w(plan, hashMask, 0, B, ctx->digitBits);
    else plan_add_hash_bits(plan, hashMask, 0, B);
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(nRecords, ctx->digitBits)));
    PG_TRY(ctx->buckets.reserve(sizeof(unsigned long long) * 2 * (size_t) nBuckets + 64));
    unsigned long long *d_start = ctx->buckets.as<unsigned long long>();
    unsigned long long *d_end = d_start + nBuckets;
    unsigned long long *d_min = ctx->small.as<unsigned long long>() + 28;     // [28] min k-mer, [29] overflow flag
    unsigned *d_over = (unsigned *) (d_min + 1);
    Rec *sorted = nullptr;
    cudaEventRecord(ctx->ev[EV_SORT1_BEGIN], s);
    // the last partition pass knows every record's final position: it also leaves the bucket boundaries and the smallest
    // k-mer; otherwise (wide digits) a sweep over the partitioned records finds them
    const bool fusedBounds = radix_emits_bounds(plan);
    RadixBounds rb;
    rb.start = d_start; rb.end = d_end; rb.minKey = d_min; rb.hashMask = hashMask; rb.bucketMask = nBuckets - 1;
    PG_CUDA(cudaMemsetAsync(d_min, 0xFF, sizeof(unsigned long long), s));
    PG_CUDA(cudaMemsetAsync(d_over, 0, sizeof(unsigned), s));
    if (fusedBounds) {
        PG_CUDA(cudaMemsetAsync(d_start, 0xFF, sizeof(unsigned long long) * (size_t) nBuckets, s));
        PG_CUDA(cudaMemsetAsync(d_end, 0, sizeof(unsigned long long) * (size_t) nBuckets, s));
    }
    PG_TRY(radix_sort(ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), nRecords, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches,
                      ctx->ev[EV_SCATTER1_BEGIN], ctx->ev[EV_SCATTER1_END], fusedBounds ? &rb : nullptr,
                      preHist ? ctx->preHist.as<unsigned long long>() : nullptr));
    ctx->timings.sort1_passes = (uint32_t) plan.npasses;
    cudaEventRecord(ctx->ev[EV_SORT1_END], s);
    if (!fusedBounds) {
        PG_CUDA(cudaMemsetAsync(d_start, 0, sizeof(unsigned long long) * 2 * (size_t) nBuckets, s));
        bucket_bounds_kernel<<<NUM_SMS * 16, 256, 0, s>>>(sorted, nRecords, hashMask, nBuckets - 1, d_start, d_end, d_min);
    }
    Rec *outBuf = (sorted == ctx->recA.as<Rec>()) ? ctx->recB.as<Rec>() : ctx->recA.as<Rec>();
    unsigned long long *d_cnt = ctx->small.as<unsigned long long>() + 2;
    PG_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
    // buckets above the small instance's capacity are listed (after the d_end array's alignment slack: reuse blockCounts)
    PG_TRY(ctx->blockCounts.reserve(sizeof(unsigned) * (2 * (size_t) nBuckets + 8)));
    unsigned *d_bigList = ctx->blockCounts.as<unsigned>() + 4;
    unsigned *d_bigCnt = ctx->blockCounts.as<unsigned>();
    unsigned *d_hugeList = d_bigList + nBuckets + 4;
    unsigned *d_hugeCnt = d_over;                 // the former overflow flag: number of buckets on the spill list
    PG_CUDA(cudaMemsetAsync(d_bigCnt, 0, sizeof(unsigned), s));
    // multi-GPU: the smallest k-mer of the whole job (all-reduced by pg_shard_iteration), not of this rank's share
    unsigned long long *d_first = d_min;
    km_min_kmer_slot(ctx, &d_first);
    if (bigFirst) {
        hash_group_kernel<HG_TABLE_BIG, HG_ITEMS_BIG, 0><<<std::min<unsigned>(nBuckets, NUM_SMS * 32), HG_THREADS, 0, s>>>(
            sorted, d_start, d_end, nBuckets, hashMask, d_first, c, outBuf, d_cnt, d_bigList, d_bigCnt, d_hugeList, d_hugeCnt);
    } else {
        hash_group_kernel<HG_TABLE_SMALL, HG_ITEMS_SMALL, 0><<<std::min<unsigned>(nBuckets, NUM_SMS * 64), HG_THREADS, 0, s>>>(
            sorted, d_start, d_end, nBuckets, hashMask, d_first, c, outBuf, d_cnt, d_bigList, d_bigCnt, d_hugeList, d_hugeCnt);
        hash_group_kernel<HG_TABLE_BIG, HG_ITEMS_BIG, 1><<<NUM_SMS * 4, HG_THREADS, 0, s>>>(
            sorted, d_start, d_end, nBuckets, hashMask, d_first, c, outBuf, d_cnt, d_bigList, d_bigCnt, d_hugeList, d_hugeCnt);
    }
    ctx->launches += 3;
    unsigned long long h = 0; unsigned nHuge = 0;
    PG_TRY(read_back(ctx, &nHuge, d_hugeCnt, sizeof(nHuge)));
    if (nHuge) {
        // spill list: copy the listed buckets out, sort them by the full k-mer, group them with the tile kernel; their pairs
        // are appended to the same output.  If the spill is most of the input the whole input takes that path instead.
        PG_TRY(ctx->scratch.reserve(sizeof(unsigned long long) * ((size_t) nHuge + 2)));
        unsigned long long *d_off = ctx->scratch.as<unsigned long long>();
        huge_offsets_kernel<<<1, 32, 0, s>>>(d_hugeList, nHuge, d_start, d_end, d_off);
        unsigned long long nSpill = 0;
        PG_TRY(read_back(ctx, &nSpill, d_off + nHuge, sizeof(nSpill)));
        if (nSpill > nRecords / 2) {
            // the records are only permuted (still all in `sorted`); hand them back in recA for the full sort
            if (sorted != ctx->recA.as<Rec>()) PG_CUDA(cudaMemcpyAsync(ctx->recA.p, sorted, sizeof(Rec) * nRecords, cudaMemcpyDeviceToDevice, s));
            return 0;
        }
        PG_TRY(ctx->spill.reserve(sizeof(Rec) * 2 * (size_t) (nSpill + 1)));
        Rec *spA = ctx->spill.as<Rec>(), *spB = spA + (nSpill + 1);
        huge_gather_kernel<<<std::min<unsigned>(nHuge, NUM_SMS * 8), 256, 0, s>>>(sorted, d_hugeList, nHuge, d_start, d_end, d_off, spA);
        RadixPlan full; full.npasses = 0;
        plan_add_bits(full, 0, 0, c.nt ? 63 : 64);
        Rec *spSorted = nullptr;
        PG_TRY(radix_sort(spA, spB, nSpill, full, ctx->radixWs.p, ctx->radixWs.cap, s, &spSorted, &ctx->launches));
        group_kernel<false><<<(unsigned) ((nSpill + GROUP_TILE - 1) / GROUP_TILE), GROUP_THREADS, 0, s>>>(spSorted, nSpill, c, outBuf, d_cnt, d_first);
        ctx->launches += 3;
        ctx->timings.spilled_records = nSpill;
    }
    cudaEventRecord(ctx->ev[EV_GROUP_END], s);
    PG_TRY(read_back(ctx, &h, d_cnt, sizeof(h)));
    PG_CUDA(cudaGetLastError());
    *nPairs = h;
    ctx->pairsInA = (outBuf == ctx->recA.as<Rec>());
    *ok = true;
    return 0;
}

// Stage 2: group the k-mer records and emit the pair records.  Input records in recA (n), output pair records in
// recA or recB (ctx->pairsInA).
int km_group(Context *ctx, const pg_seqdb *db, const KmConst &c, uint64_t nRecords, uint64_t *nPairs) {
    cudaStream_t s = ctx->stream;
    *nPairs = 0;
    // the extraction's digit histograms describe exactly the records it has just left in recA: valid for the call that follows
    // it directly, never for records that arrived from other ranks
    const bool preHist = ctx->preHistValid && ctx->preHistRecords == nRecords && !ctx->noPreHist;
    ctx->preHistValid = false;
    if (nRecords == 0) return 0;
    if (!ctx->forceFullSort && !c.wide) {
        bool ok = false;
        PG_TRY(km_group_bucketed(ctx, c, nRecords, nPairs, &ok, preHist));
        if (ok) return 0;
    }
    RadixPlan plan; plan.npasses = 0;
    plan_add_bits(plan, 0, 0, c.nt ? 63 : 64);   // nt: bit 63 is the strand flag, not part of the key (kmermatcher.h:77-96)
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(nRecords)));
    Rec *sorted = nullptr;
    cudaEventRecord(ctx->ev[EV_SORT1_BEGIN], s);
    PG_TRY(radix_sort(ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), nRecords, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches,
                      ctx->ev[EV_SCATTER1_BEGIN], ctx->ev[EV_SCATTER1_END]));
    ctx->timings.sort1_passes = (uint32_t) plan.npasses;
    cudaEventRecord(ctx->ev[EV_SORT1_END], s);
    Rec *outBuf = (sorted == ctx->recA.as<Rec>()) ? ctx->recB.as<Rec>() : ctx->recA.as<Rec>();
    unsigned long long *d_cnt = ctx->small.as<unsigned long long>() + 2;
    PG_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
    const unsigned blocks = (unsigned) ((nRecords + GROUP_TILE - 1) / GROUP_TILE);
    if (c.wide) group_kernel<true><<<blocks, GROUP_THREADS, 0, s>>>(sorted, nRecords, c, outBuf, d_cnt);
    else group_kernel<false><<<blocks, GROUP_THREADS, 0, s>>>(sorted, nRecords, c, outBuf, d_cnt);
    ctx->launches++;
    cudaEventRecord(ctx->ev[EV_GROUP_END], s);
    unsigned long long h = 0;
    PG_TRY(read_back(ctx, &h, d_cnt, sizeof(h)));
    PG_CUDA(cudaGetLastError());
    *nPairs = h;
    ctx->pairsInA = (outBuf == ctx->recA.as<Rec>());
    (void) db;
    return 0;
}

// fast path of stage 3 (see reduce_rep_warp_kernel); *ok = false if a representative exceeded the CTA capacity
static int km_reduce_segmented(Context *ctx, const pg_seqdb *db, Rec **pairsIO, Rec **tmpIO, uint64_t nPairs, pg_hit **d_hits, uint64_t *nHits, bool *ok) {
    cudaStream_t s = ctx->stream;
    *ok = false;
    Rec *pairs = *pairsIO, *tmp = *tmpIO;
    const int keyBits = bits_for(db->max_key);
    const unsigned nKeys = db->max_key + 1;
    const unsigned keyLo = std::min(ctx->ownLo, nKeys), keyHi = std::min(ctx->ownHi, nKeys);   // representatives of this rank
    RadixPlan plan; plan.npasses = 0;
    if (ctx->digitBits > 8) plan_add_bits_w(plan, 0, 32, 32 + keyBits, ctx->digitBits);
    else if (keyLo > 0 || keyHi < nKeys) plan_add_rebased_high_bits(plan, keyLo, bits_for(keyHi > keyLo ? keyHi - 1 - keyLo : 0));   // multi-GPU: only the owned key range
    else plan_add_bits(plan, 0, 32, 32 + keyBits);
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(nPairs, ctx->digitBits)));
    Rec *sorted = pairs;
    // per-representative tables: only the owned key range [keyLo, keyHi) (multi-GPU: 1 / world of the keys), addressed by the
    // key itself through pointers shifted by keyLo
    const size_t nT = keyHi > keyLo ? (size_t) (keyHi - keyLo) : 0;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t) 15; return r; };
    const size_t oStart = take(sizeof(unsigned long long) * (nT + 1)), oEnd = take(sizeof(unsigned long long) * (nT + 1));
    const size_t oCnt = take(sizeof(unsigned) * (nT + 1));
    const size_t oOff = take(sizeof(unsigned long long) * (nT + 2)), oBig = take(sizeof(unsigned) * (nT + 1));
    const size_t oMid = take(sizeof(unsigned) * (nT + 1));
    const size_t oHuge = take(sizeof(unsigned) * (nT + 1));
    const size_t oScan = take(scan_workspace_bytes(nT));
    const size_t oMinT = take(sizeof(unsigned) * (nT + 1));
    PG_TRY(ctx->buckets2.reserve(o));
    unsigned char *bb = ctx->buckets2.as<unsigned char>();
    unsigned long long *d_start0 = (unsigned long long *) (bb + oStart), *d_end0 = (unsigned long long *) (bb + oEnd);
    unsigned *d_hcnt0 = (unsigned *) (bb + oCnt), *d_big = (unsigned *) (bb + oBig), *d_minT0 = (unsigned *) (bb + oMinT), *d_mid = (unsigned *) (bb + oMid);
    unsigned long long *d_hoff0 = (unsigned long long *) (bb + oOff);
    unsigned long long *d_start = d_start0 - keyLo, *d_end = d_end0 - keyLo, *d_hoff = d_hoff0 - keyLo;
    unsigned *d_hcnt = d_hcnt0 - keyLo, *d_minT = d_minT0 - keyLo;
    unsigned *d_over = (unsigned *) (ctx->small.as<unsigned long long>() + 32);     // [32] overflow flag, big count, mid count; [34] total hits
    unsigned *d_bigCnt = d_over + 1, *d_midCnt = d_over + 2, *d_hugeCnt = d_over + 3;
    unsigned *d_huge = (unsigned *) (bb + oHuge);
    unsigned long long *d_total = ctx->small.as<unsigned long long>() + 34;
    PG_CUDA(cudaMemsetAsync(d_start0, 0, oOff, s));   // start, end, hit counts
    PG_CUDA(cudaMemsetAsync(d_over, 0, 4 * sizeof(unsigned), s));
    PG_CUDA(cudaMemsetAsync(d_minT0, 0xFF, sizeof(unsigned) * (nT + 1), s));
    // the last pass of the sort knows every pair's final position: it also leaves the per-representative segment bounds and
    // smallest targets (otherwise seg_bounds_kernel sweeps the sorted pairs once more)
    const bool fusedSegments = radix_emits_segments(plan);
    RadixBounds rb;
    rb.kind = 1; rb.start = d_start; rb.end = d_end; rb.minLow = d_minT;
    if (fusedSegments) PG_CUDA(cudaMemsetAsync(d_start0, 0xFF, sizeof(unsigned long long) * (nT + 1), s));
    PG_TRY(radix_sort(pairs, tmp, nPairs, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches, nullptr, nullptr, fusedSegments ? &rb : nullptr));
    cudaEventRecord(ctx->ev[EV_SORT2_END], s);
    Rec *other = (sorted == pairs) ? tmp : pairs;
    if (!fusedSegments) seg_bounds_kernel<<<NUM_SMS * 16, 256, 0, s>>>(sorted, nPairs, d_start, d_end, d_minT);
    pg_hit *tmpHits = reinterpret_cast<pg_hit *>(other);   // a hit is 16 bytes like a record, at most one per pair
    reduce_rep_warp_kernel<0><<<NUM_SMS * 32, 256, 0, s>>>(sorted, nPairs, d_start, d_end, d_minT, keyLo, keyHi, tmpHits, d_hcnt, d_big, d_bigCnt, d_mid, d_midCnt, d_over);
    // 129 .. 512 pairs: warp per representative, > 512: CTA per representative, both aggregating by (target, diagonal)
    // before they sort; what does not fit their tables is sorted in full by reduce_rep_block_kernel
    reduce_rep_hashwarp_kernel<<<NUM_SMS * 8, RH_WARPS * 32, 0, s>>>(sorted, nPairs, d_start, d_end, d_minT, d_mid, d_midCnt, tmpHits, d_hcnt, d_over);
    reduce_rep_hashcta_kernel<<<NUM_SMS * 4, 256, 0, s>>>(sorted, nPairs, d_start, d_end, d_minT, d_big, d_bigCnt, d_huge, d_hugeCnt, tmpHits, d_hcnt, d_over);
    PG_CUDA(cudaFuncSetAttribute(reduce_rep_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_BLOCK_MAX * (int) sizeof(unsigned long long)));
    reduce_rep_block_kernel<<<NUM_SMS, 256, SEG_BLOCK_MAX * sizeof(unsigned long long), s>>>(sorted, nPairs, d_start, d_end, d_minT, d_huge, d_hugeCnt, tmpHits, d_hcnt, d_over);
    ctx->launches += 5;
    PG_TRY(exclusive_scan_u32(d_hcnt0, d_hoff0, nT, d_total, bb + oScan, scan_workspace_bytes(nT), s, &ctx->launches));
    unsigned long long h = 0; unsigned over = 0;
    {
        unsigned long long hb[3];                      // small[32]: overflow flag (low word) ... small[34]: total hits
        PG_TRY(read_back(ctx, hb, d_over, sizeof(hb)));
        over = (unsigned) hb[0]; h = hb[2];
    }
    PG_CUDA(cudaGetLastError());
    if (over) { *pairsIO = sorted; *tmpIO = other; return 0; }
    PG_TRY(ctx->hits.reserve(sizeof(pg_hit) * (h + 1)));
    PG_CUDA(cudaStreamWaitEvent(s, ctx->evHitsCopied, 0));   // an asynchronous copy of the previous call's hits may still read the buffer
    compact_rep_hits_kernel<<<NUM_SMS * 16, 256, 0, s>>>(tmpHits, d_start, d_hcnt, d_hoff, keyLo, keyHi, ctx->hits.as<pg_hit>());
    ctx->launches++;
    cudaEventRecord(ctx->ev[EV_REDUCE_END], s);
    PG_CUDA(cudaGetLastError());
    *d_hits = ctx->hits.as<pg_hit>();
    *nHits = h;
    *ok = true;
    return 0;
}

// Stage 3: sort #2 + best-diagonal reduction.  Pairs are in `pairs` (device), scratch in `tmp`.
int km_reduce(Context *ctx, const pg_seqdb *db, Rec *pairs, Rec *tmp, uint64_t nPairs, pg_hit **d_hits, uint64_t *nHits) {
    cudaStream_t s = ctx->stream;
    *nHits = 0; *d_hits = nullptr;
    if (nPairs == 0) return 0;
    const bool wide = km_is_wide(db);
    // the segmented path keeps about ten arrays indexed by representative KEY: only for (nearly) dense key spaces; a DB
    // with sparse keys (a subset that kept its ids) takes the full sort, which needs no per-key tables
    const bool sparseKeys = (unsigned long long) db->max_key + 1ull > 4ull * db->n + 1024ull;
    if (!ctx->forceFullSort && !wide && !sparseKeys) {
        bool ok = false;
        PG_TRY(km_reduce_segmented(ctx, db, &pairs, &tmp, nPairs, d_hits, nHits, &ok));
        if (ok) return 0;
    }
    const int keyBits = bits_for(db->max_key);
    RadixPlan plan; plan.npasses = 0;
    plan_add_bits(plan, 1, 0, wide ? 32 : 16);     // diagonal (biased so that the signed order is kept)
    plan_add_bits(plan, 0, 0, keyBits);            // target id
    plan_add_bits(plan, 0, 32, 32 + keyBits);      // representative
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(nPairs)));
    Rec *sorted = nullptr;
    PG_TRY(radix_sort(pairs, tmp, nPairs, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches));
    cudaEventRecord(ctx->ev[EV_SORT2_END], s);
    const unsigned long long blocks = (nPairs + 255) / 256;
    PG_TRY(ctx->blockCounts.reserve(sizeof(unsigned) * blocks + sizeof(unsigned long long) * (blocks + 2)));
    unsigned *d_counts = ctx->blockCounts.as<unsigned>();
    unsigned long long *d_offsets = (unsigned long long *) (ctx->blockCounts.as<unsigned char>() + ((sizeof(unsigned) * blocks + 15) & ~(size_t) 15));
    unsigned long long *d_total = ctx->small.as<unsigned long long>() + 3;
    if (wide) reduce_count_kernel<true><<<(unsigned) blocks, 256, 0, s>>>(sorted, nPairs, d_counts);
    else reduce_count_kernel<false><<<(unsigned) blocks, 256, 0, s>>>(sorted, nPairs, d_counts);
    scan_counts_kernel<<<1, 1024, 0, s>>>(d_counts, blocks, d_offsets, d_total);
    unsigned long long h = 0;
    PG_TRY(read_back(ctx, &h, d_total, sizeof(h)));
    PG_TRY(ctx->hits.reserve(sizeof(pg_hit) * (h + 1)));
    PG_CUDA(cudaStreamWaitEvent(s, ctx->evHitsCopied, 0));
    if (wide) reduce_emit_kernel<true><<<(unsigned) blocks, 256, 0, s>>>(sorted, nPairs, d_offsets, ctx->hits.as<pg_hit>());
    else reduce_emit_kernel<false><<<(unsigned) blocks, 256, 0, s>>>(sorted, nPairs, d_offsets, ctx->hits.as<pg_hit>());
    ctx->launches += 3;
    cudaEventRecord(ctx->ev[EV_REDUCE_END], s);
    PG_CUDA(cudaGetLastError());
    *d_hits = ctx->hits.as<pg_hit>();
    *nHits = h;
    return 0;
}

static void record_empty_group_events(Context *ctx) {
    cudaStream_t s = ctx->stream;
    cudaEventRecord(ctx->ev[EV_SORT1_BEGIN], s); cudaEventRecord(ctx->ev[EV_SCATTER1_BEGIN], s); cudaEventRecord(ctx->ev[EV_SCATTER1_END], s);
    cudaEventRecord(ctx->ev[EV_SORT1_END], s); cudaEventRecord(ctx->ev[EV_GROUP_END], s);
}

// kmermatcherInner's split decision (kmermatcher.cpp:608-624): how many equal hash ranges are needed so that the two
// record buffers of one split fit the memory limit.  The estimate of the record count is computeKmerCount (:576-585).
static int km_choose_splits(Context *ctx, const pg_seqdb *db, const pg_km_params *p, unsigned *splits) {
    *splits = 1;
    if (ctx->forceSplits) { *splits = ctx->forceSplits; return 0; }
    cudaStream_t s = ctx->stream;
    PG_TRY(ctx->small.reserve(4096));
    unsigned long long *d_total = ctx->small.as<unsigned long long>() + 50;
    PG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s));
    kmer_count_kernel<<<NUM_SMS * 4, 256, 0, s>>>(db->lens, 0, (unsigned) db->n, p->kmer_size, p->kmers_per_seq, p->kmers_per_seq_scale, d_total);
    ctx->launches++;
    unsigned long long total = 0;
    PG_TRY(read_back(ctx, &total, d_total, sizeof(total)));
    ctx->kmerTotalHint = total + 1;             // km_extract does not count again
    const unsigned long long need = 2ull * sizeof(Rec) * (total + 1) + radix_workspace_bytes(total);
    unsigned long long limit = ctx->memLimit;
    // cudaMemGetInfo waits for the device (it would serialise this call with the previous step's result transfers on the copy
    // stream): only asked when the stage needs a sizeable part of the device memory
    if (limit == 0 && (need <= ctx->deviceMemBytes / 4 || need <= ctx->recA.cap + ctx->recB.cap + ctx->radixWs.cap)) return 0;   // small, or the buffers exist already
    if (limit == 0) {
        size_t freeB = 0, totalB = 0;
        PG_CUDA(cudaMemGetInfo(&freeB, &totalB));
        // what the stage's own buffers already hold is reusable
        limit = (unsigned long long) (0.9 * (double) (freeB + ctx->recA.cap + ctx->recB.cap + ctx->radixWs.cap));
    }
    if (need <= limit) return 0;
    unsigned sp = 2;
    while (sp < 65536u && 2ull * sizeof(Rec) * (total / sp + total / (8ull * sp) + 65536ull) + radix_workspace_bytes(total / sp + 65536ull) > limit) sp <<= 1;
    PG_CHECK(2ull * sizeof(Rec) * (total / sp + total / (8ull * sp) + 65536ull) <= limit || sp < 65536u,
             "kmermatcher: --split-memory-limit is too small for even one of 65536 hash-range splits");
    *splits = sp;
    return 0;
}

// kmermatcher in hash-range splits (setupKmerSplits, kmermatcher.cpp:736-778): every split extracts, partitions and
// groups only the k-mers whose 16-bit hash lies in its range; the pair records of all splits are collected in
// ctx->pairAcc.  Equal k-mers have equal hashes, so the union of the splits' groups is the unsplit run's set of groups.
static int km_run_splits(Context *ctx, const pg_seqdb *db, const pg_km_params *p, KmConst &c, unsigned splits, uint64_t *nRecTotal, uint64_t *nPairsTotal) {
    cudaStream_t s = ctx->stream;
    for (;;) {
        ctx->splitDiv = splits;
        uint64_t recTotal = 0, pairTotal = 0;
        bool overflow = false;
        for (unsigned sp = 0; sp < splits && !overflow; sp++) {
            // ranges over the reference's own 16-bit hash (hashStart / hashEnd inclusive), clipped to the caller's range
            const unsigned lo = (unsigned) ((65536ull * sp) / splits), hi = (unsigned) ((65536ull * (sp + 1)) / splits) - 1u;
            c.hashStart = std::max(lo, p->hash_start); c.hashEnd = std::min(hi, p->hash_end);
            if (c.hashStart > c.hashEnd) continue;
            uint64_t nRec = 0, nPairs = 0;
            const int rc = km_extract(ctx, db, p, c, &nRec);
            if (rc == 2) { overflow = true; break; }
            if (rc) { ctx->splitDiv = 1; return rc; }
            if (sp == 0) cudaEventRecord(ctx->ev[EV_EXTRACT_END], s);
            const int rg = km_group(ctx, db, c, nRec, &nPairs);
            if (rg) { ctx->splitDiv = 1; return rg; }
            if (nRec == 0) record_empty_group_events(ctx);
            if (nPairs) {
                // grow-and-copy accumulation of the pair records
                const size_t needBytes = sizeof(Rec) * (pairTotal + nPairs + 1);
                if (needBytes > ctx->pairAcc.cap) {
                    DevBuf bigger;
                    if (bigger.reserve(needBytes + needBytes / 2)) { ctx->splitDiv = 1; return 1; }
                    if (pairTotal) cudaMemcpyAsync(bigger.p, ctx->pairAcc.p, sizeof(Rec) * pairTotal, cudaMemcpyDeviceToDevice, s);
                    cudaStreamSynchronize(s);
                    ctx->pairAcc.release();
                    ctx->pairAcc = bigger;
                }
                const Rec *src = ctx->pairsInA ? ctx->recA.as<Rec>() : ctx->recB.as<Rec>();
                cudaMemcpyAsync(ctx->pairAcc.as<Rec>() + pairTotal, src, sizeof(Rec) * nPairs, cudaMemcpyDeviceToDevice, s);
            }
            recTotal += nRec; pairTotal += nPairs;
        }
        if (overflow) {
            PG_CHECK(splits < 65536u, "kmermatcher: a single 16-bit hash value holds more k-mer records than the memory limit allows");
            splits <<= 1;
            continue;
        }
        ctx->splitDiv = 1;
        *nRecTotal = recTotal; *nPairsTotal = pairTotal;
        ctx->timings.splits = splits;
        return 0;
    }
}

// Whole kmermatcher on one GPU; hits stay on the device (ctx->hits).
int km_run(Context *ctx, const pg_seqdb *db, const pg_km_params *p, pg_hit **d_hits, uint64_t *nHits) {
    KmConst c;
    cudaStream_t s = ctx->stream;
    PG_TRY(km_setup_constants(db, p, c, s));
    cudaEventRecord(ctx->ev[EV_KM_BEGIN], s);
    uint64_t nRec = 0, nPairs = 0;
    PG_TRY(km_prepare_wide(ctx, db, c));
    unsigned splits = 1;
    PG_TRY(km_choose_splits(ctx, db, p, &splits));
    ctx->timings.splits = 1;
    if (splits > 1) {
        PG_TRY(km_run_splits(ctx, db, p, c, splits, &nRec, &nPairs));
        // reduce all splits' pairs together: pairAcc is the input, recA the scratch of sort #2
        ctx->recB.release();
        PG_TRY(ctx->recA.reserve(sizeof(Rec) * (nPairs + 1)));
        if (nPairs == 0) { cudaEventRecord(ctx->ev[EV_SORT2_END], s); cudaEventRecord(ctx->ev[EV_REDUCE_END], s); }
        PG_TRY(km_reduce(ctx, db, ctx->pairAcc.as<Rec>(), ctx->recA.as<Rec>(), nPairs, d_hits, nHits));
        PG_CUDA(cudaStreamSynchronize(s));
        ctx->pairAcc.release();
        ctx->timings.n_kmer_records = nRec;
        ctx->timings.n_pair_records = nPairs;
        ctx->timings.n_hits = *nHits;
        ctx->timings.sort1_bytes = (uint64_t) nRec * sizeof(Rec) * 2;
        ctx->tExtract = ctx->tGroup = ctx->tReduce = true;
        return 0;
    }
    PG_TRY(km_extract(ctx, db, p, c, &nRec));
    cudaEventRecord(ctx->ev[EV_EXTRACT_END], s);
    PG_TRY(km_group(ctx, db, c, nRec, &nPairs));
    Rec *pairs = ctx->pairsInA ? ctx->recA.as<Rec>() : ctx->recB.as<Rec>();
    Rec *tmp = ctx->pairsInA ? ctx->recB.as<Rec>() : ctx->recA.as<Rec>();
    if (nRec == 0) {
        cudaEventRecord(ctx->ev[EV_SORT1_BEGIN], s); cudaEventRecord(ctx->ev[EV_SCATTER1_BEGIN], s); cudaEventRecord(ctx->ev[EV_SCATTER1_END], s); cudaEventRecord(ctx->ev[EV_SORT1_END], s); cudaEventRecord(ctx->ev[EV_GROUP_END], s);
    }
    if (nPairs == 0) { cudaEventRecord(ctx->ev[EV_SORT2_END], s); cudaEventRecord(ctx->ev[EV_REDUCE_END], s); }
    PG_TRY(km_reduce(ctx, db, pairs, tmp, nPairs, d_hits, nHits));
    ctx->timings.n_kmer_records = nRec;
    ctx->timings.n_pair_records = nPairs;
    ctx->timings.n_hits = *nHits;
    ctx->timings.sort1_bytes = (uint64_t) nRec * sizeof(Rec) * 2;
    ctx->tExtract = ctx->tGroup = ctx->tReduce = true;
    return 0;
}

// ---- multi-GPU: the k-mer hash space is sharded over the ranks (the reference's split mechanism,
// kmermatcher.cpp:736-778: each split extracts only the k-mers whose 16-bit hash falls in its range), the
// (rep, target, diagonal) pairs are then routed to the rank that owns the representative (contiguous key
// ranges) with one all-to-all; sort #2 and everything downstream is local to the owner. -------------------
constexpr int SHARD_HIST_BINS = 4096;

// owner of a representative = the interval of bounds[0..world] (ascending keys, bounds[0] = 0) that contains it
__global__ void tag_owner_kernel(Rec *__restrict__ pairs, unsigned long long n, const unsigned *__restrict__ bounds, unsigned world) {
    __shared__ unsigned sB[257];
    for (unsigned i = threadIdx.x; i <= world; i += blockDim.x) sB[i] = bounds[i];
    __syncthreads();
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned rep = (unsigned) (pairs[i].w0 >> 32);
        unsigned lo = 0, hi = world;              // invariant: sB[lo] <= rep, (hi == world or rep < sB[hi])
        while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (rep >= sB[mid]) lo = mid; else hi = mid; }
        pairs[i].w1 = (pairs[i].w1 & 0xFFFFFFULL) | ((unsigned long long) lo << 24);
    }
}

// pair records per slice of the representative key space: bin = rep * SHARD_HIST_BINS / (max_key + 1)
__global__ void rep_hist_kernel(const Rec *__restrict__ pairs, unsigned long long n, unsigned long long keySpan, unsigned long long *__restrict__ hist) {
    __shared__ unsigned sH[SHARD_HIST_BINS];
    for (int i = threadIdx.x; i < SHARD_HIST_BINS; i += blockDim.x) sH[i] = 0;
    __syncthreads();
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned long long rep = pairs[i].w0 >> 32;
        const unsigned bin = (unsigned) ((rep * SHARD_HIST_BINS) / keySpan);
        // neighbouring records often share the representative: aggregate per warp before touching shared memory
        const unsigned peers = __match_any_sync(__activemask(), bin);
        if ((int) lane_id() == __ffs(peers) - 1) atomicAdd(&sH[bin], __popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SHARD_HIST_BINS; i += blockDim.x) if (sH[i]) atomicAdd(&hist[i], (unsigned long long) sH[i]);
}

// pair records (recA/recB per ctx->pairsInA) -> tagged with the rank that owns the representative and partitioned by it
int km_shard_route(Context *ctx, int world, const unsigned *bounds, uint64_t *counts) {
    PG_CHECK(world >= 1 && world <= 256, "pg_shard_route: world size must be in [1, 256]");
    PG_CHECK(bounds[0] == 0, "pg_shard_route: bounds[0] must be 0");
    for (int r = 0; r < world; r++) PG_CHECK(bounds[r] <= bounds[r + 1], "pg_shard_route: bounds must be ascending");
    cudaStream_t s = ctx->stream;
    const uint64_t nPairs = ctx->shardPairCount;
    Rec *pairs = ctx->pairsInA ? ctx->recA.as<Rec>() : ctx->recB.as<Rec>();
    Rec *tmp = ctx->pairsInA ? ctx->recB.as<Rec>() : ctx->recA.as<Rec>();
    for (int r = 0; r < world; r++) counts[r] = 0;
    ctx->shardPairs = pairs;
    if (nPairs == 0) return 0;
    PG_TRY(ctx->small.reserve(4096));
    unsigned *d_bounds = (unsigned *) (ctx->small.as<unsigned long long>() + 64);     // [64..] 257 x u32
    PG_CUDA(cudaMemcpyAsync(d_bounds, bounds, sizeof(unsigned) * (world + 1), cudaMemcpyHostToDevice, s));
    // one partition pass whose digit is the owner: the interval of bounds[] that holds the representative
    RadixPlan plan; plan.npasses = 0;
    plan_add_interval(plan, d_bounds, (unsigned) world);
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(nPairs)));
    Rec *sorted = nullptr;
    PG_TRY(radix_sort(pairs, tmp, nPairs, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches));
    ctx->launches++;
    unsigned long long h[256];
    PG_CUDA(cudaMemcpyAsync(h, ctx->radixWs.p, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost, s));   // digit histogram of the pass
    PG_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < world; r++) counts[r] = h[r];
    ctx->shardPairs = sorted;
    ctx->pairsInA = (sorted == ctx->recA.as<Rec>());
    return 0;
}

// after km_group: remember the pairs and histogram them over the representative key space
static int shard_pairs_ready(Context *ctx, const pg_seqdb *db, uint64_t nPairs, uint64_t *hist) {
    cudaStream_t s = ctx->stream;
    ctx->shardPairs = ctx->pairsInA ? ctx->recA.as<Rec>() : ctx->recB.as<Rec>();
    ctx->shardPairCount = nPairs;
    if (!hist) return 0;
    for (int i = 0; i < SHARD_HIST_BINS; i++) hist[i] = 0;
    if (nPairs == 0) return 0;
    PG_TRY(ctx->buckets2.reserve(sizeof(unsigned long long) * SHARD_HIST_BINS));
    unsigned long long *d_hist = ctx->buckets2.as<unsigned long long>();
    PG_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * SHARD_HIST_BINS, s));
    rep_hist_kernel<<<NUM_SMS * 4, 512, 0, s>>>(ctx->shardPairs, nPairs, (unsigned long long) db->max_key + 1ull, d_hist);
    ctx->launches++;
    PG_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof(unsigned long long) * SHARD_HIST_BINS, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

void km_equal_key_bounds(unsigned max_key, int world, unsigned *bounds) {
    const unsigned long long per = ((unsigned long long) max_key + (unsigned long long) world) / (unsigned long long) world;
    for (int r = 0; r <= world; r++) bounds[r] = (r == world) ? 0xFFFFFFFFu : (unsigned) std::min<unsigned long long>(per * (unsigned long long) r, 0xFFFFFFFFull);
}

int km_shard_pairs(Context *ctx, const pg_seqdb *db, const pg_km_params *p, int world, uint64_t *counts) {
    PG_CHECK(world >= 1 && world <= 256, "pg_shard_pairs: world size must be in [1, 256]");
    PG_CHECK(!km_is_wide(db), "multi-GPU kmermatcher: sequences >= 32765 residues (wide T=int records) are single-GPU only for now");
    KmConst c;
    cudaStream_t s = ctx->stream;
    PG_TRY(km_setup_constants(db, p, c, s));
    cudaEventRecord(ctx->ev[EV_KM_BEGIN], s);
    uint64_t nRec = 0, nPairs = 0;
    PG_TRY(km_extract(ctx, db, p, c, &nRec));
    cudaEventRecord(ctx->ev[EV_EXTRACT_END], s);
    PG_TRY(km_group(ctx, db, c, nRec, &nPairs));
    if (nRec == 0) record_empty_group_events(ctx);
    ctx->timings.n_kmer_records = nRec; ctx->timings.n_pair_records = nPairs;
    ctx->timings.sort1_bytes = (uint64_t) nRec * sizeof(Rec) * 2;
    ctx->tExtract = ctx->tGroup = true;
    PG_TRY(shard_pairs_ready(ctx, db, nPairs, nullptr));
    unsigned bounds[257];
    km_equal_key_bounds(db->max_key, world, bounds);
    return km_shard_route(ctx, world, bounds, counts);
}

// Two-exchange decomposition, phase 0: k-mer extraction of this rank's slice of the SEQUENCES (all hash values), the
// records partitioned by the rank that owns the k-mer: owner = (top 8 bits of mix64(k-mer)) * world / 256.  Any
// partition under which equal k-mers meet is equivalent; this one is independent of the bucket bits of the hash join.
int km_shard_extract(Context *ctx, const pg_seqdb *db, const pg_km_params *p, int rank, int world, uint64_t *counts) {
    PG_CHECK(world >= 1 && world <= 256 && rank >= 0 && rank < world, "pg_shard_extract: bad rank / world size");
    PG_CHECK(!km_is_wide(db), "multi-GPU kmermatcher: sequences >= 32765 residues (wide T=int records) are single-GPU only for now");
    KmConst c;
    cudaStream_t s = ctx->stream;
    PG_TRY(km_setup_constants(db, p, c, s));
    cudaEventRecord(ctx->ev[EV_KM_BEGIN], s);
    uint64_t nRec = 0;
    ctx->seqLo = (unsigned) ((unsigned long long) db->n * (unsigned) rank / (unsigned) world);
    ctx->seqHi = (unsigned) ((unsigned long long) db->n * (unsigned) (rank + 1) / (unsigned) world);
    const int rc = km_extract(ctx, db, p, c, &nRec);
    ctx->seqLo = 0; ctx->seqHi = 0xFFFFFFFFu;
    ctx->preHistValid = false;                    // these records are exchanged before they are grouped
    if (rc) return rc;
    for (int r = 0; r < world; r++) counts[r] = 0;
    ctx->shardPairs = ctx->recA.as<Rec>(); ctx->shardPairCount = nRec;
    ctx->timings.n_kmer_records = nRec;
    ctx->tExtract = true;
    if (nRec == 0 || world == 1) { counts[0] = nRec; cudaEventRecord(ctx->ev[EV_EXTRACT_END], s); return 0; }
    RadixPlan plan; plan.npasses = 0;
    plan_add_hash_bits(plan, c.nt ? ~(1ULL << 63) : ~0ULL, 56, 64);
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(nRec)));
    Rec *sorted = nullptr;
    PG_TRY(radix_sort(ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), nRec, plan, ctx->radixWs.p, ctx->radixWs.cap, s, &sorted, &ctx->launches));
    cudaEventRecord(ctx->ev[EV_EXTRACT_END], s);
    unsigned long long h[256];
    PG_CUDA(cudaMemcpyAsync(h, ctx->radixWs.p, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaStreamSynchronize(s));
    for (int b = 0; b < 256; b++) counts[(unsigned) b * (unsigned) world / 256u] += h[b];
    ctx->shardPairs = sorted;
    return 0;
}

// phase 0 without the partition pass: the records of this rank's slice of the sequences, unordered, in recA (the fused
// partition + exchange of pg_shard.cu takes them from there)
int km_shard_extract_only(Context *ctx, const pg_seqdb *db, const pg_km_params *p, int rank, int world, uint64_t *nRecords) {
    PG_CHECK(world >= 1 && world <= 256 && rank >= 0 && rank < world, "pg_shard_extract: bad rank / world size");
    PG_CHECK(!km_is_wide(db), "multi-GPU kmermatcher: sequences >= 32765 residues (wide T=int records) are single-GPU only for now");
    KmConst c;
    cudaStream_t s = ctx->stream;
    PG_TRY(km_setup_constants(db, p, c, s));
    cudaEventRecord(ctx->ev[EV_KM_BEGIN], s);
    uint64_t nRec = 0;
    ctx->seqLo = (unsigned) ((unsigned long long) db->n * (unsigned) rank / (unsigned) world);
    ctx->seqHi = (unsigned) ((unsigned long long) db->n * (unsigned) (rank + 1) / (unsigned) world);
    ctx->extractOnly = true;
    const int rc = km_extract(ctx, db, p, c, &nRec);
    ctx->extractOnly = false;
    ctx->seqLo = 0; ctx->seqHi = 0xFFFFFFFFu;
    ctx->preHistValid = false;                    // these records are exchanged before they are grouped
    if (rc) return rc;
    cudaEventRecord(ctx->ev[EV_EXTRACT_END], s);
    ctx->shardPairs = ctx->recA.as<Rec>(); ctx->shardPairCount = nRec;
    ctx->timings.n_kmer_records = nRec;
    ctx->tExtract = true;
    *nRecords = nRec;
    return 0;
}

// after km_shard_group: where the pair records are
void km_shard_pairs_location(Context *ctx, Rec **pairs, uint64_t *n) { *pairs = ctx->shardPairs; *n = ctx->shardPairCount; }

// phase 1: the k-mer records this rank received (every record of the k-mers it owns) -> sort #1 + group -> pair
// records (left on the device for pg_shard_route) and their histogram over the representative key space
int km_shard_group(Context *ctx, const pg_seqdb *db, const pg_km_params *p, const void *d_records, uint64_t nRec, uint64_t *hist) {
    PG_CHECK(!km_is_wide(db), "multi-GPU kmermatcher: sequences >= 32765 residues (wide T=int records) are single-GPU only for now");
    KmConst c;
    cudaStream_t s = ctx->stream;
    PG_TRY(km_setup_constants(db, p, c, s));
    PG_TRY(ctx->recA.reserve(sizeof(Rec) * (nRec + 1)));
    PG_TRY(ctx->recB.reserve(sizeof(Rec) * (nRec + 1)));
    if (nRec && d_records != ctx->recA.p) PG_CUDA(cudaMemcpyAsync(ctx->recA.p, d_records, sizeof(Rec) * nRec, cudaMemcpyDeviceToDevice, s));
    uint64_t nPairs = 0;
    ctx->preHistValid = false;                    // records received from the other ranks
    PG_TRY(km_group(ctx, db, c, nRec, &nPairs));
    if (nRec == 0) record_empty_group_events(ctx);
    ctx->timings.n_kmer_records = nRec; ctx->timings.n_pair_records = nPairs;
    ctx->timings.sort1_bytes = (uint64_t) nRec * sizeof(Rec) * 2;
    ctx->tGroup = true;
    return shard_pairs_ready(ctx, db, nPairs, hist);
}

int km_shard_reduce(Context *ctx, const pg_seqdb *db, const void *d_pairs, uint64_t nPairs, pg_hit **d_hits, uint64_t *nHits) {
    cudaStream_t s = ctx->stream;
    PG_TRY(ctx->recA.reserve(sizeof(Rec) * (nPairs + 1)));
    PG_TRY(ctx->recB.reserve(sizeof(Rec) * (nPairs + 1)));
    if (nPairs && d_pairs != ctx->recA.p) PG_CUDA(cudaMemcpyAsync(ctx->recA.p, d_pairs, sizeof(Rec) * nPairs, cudaMemcpyDeviceToDevice, s));
    cudaEventRecord(ctx->ev[EV_GROUP_END], s);
    if (nPairs == 0) { cudaEventRecord(ctx->ev[EV_SORT2_END], s); cudaEventRecord(ctx->ev[EV_REDUCE_END], s); }
    PG_TRY(km_reduce(ctx, db, ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), nPairs, d_hits, nHits));
    ctx->timings.n_hits = *nHits;
    ctx->tReduce = true;
    return 0;
}

}  // namespace pg

// diagnostic used by the tests: the k-mer records of stage 1 (order unspecified), n x {w0, w1}
extern "C" int pg_debug_extract(pg_context *ctx, const pg_seqdb *db, const pg_km_params *p, uint64_t **recs, uint64_t *n) {
    using namespace pg;
    PG_CHECK(ctx && db && p && recs && n, "pg_debug_extract: null argument");
    cudaSetDevice(ctx->device);
    KmConst c;
    PG_TRY(km_setup_constants(db, p, c, ctx->stream));
    uint64_t nRec = 0;
    PG_TRY(km_extract(ctx, db, p, c, &nRec));
    ctx->preHistValid = false;
    uint64_t *h = nullptr;
    PG_TRY(alloc_pinned(sizeof(Rec) * (nRec + 1), (void **) &h));
    PG_CUDA(cudaMemcpyAsync(h, ctx->recA.p, sizeof(Rec) * nRec, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    *recs = h; *n = nRec;
    return 0;
}
