// pg_extend.cu -- GPU greedy contig extension (assembleresults / nuclassembleresults).
//
// Replaces doassembly (reference src/assembler/assembleresult.cpp:110-356) and doNuclAssembly
// (src/assembler/nuclassembleresult.cpp:144-398):
//   pass A  extend_kernel       one thread per query: the priority queue of the reference is replayed
//                               exactly (libstdc++ push_heap/pop_heap sequence, so that the nucleotide
//                               comparator -- not a strict weak ordering -- pops in the same order),
//                               the growing contig is kept as a rope of (source sequence, offset, length,
//                               reversed) segments, parked alignments are re-scored against the rope;
//   scan                        output offsets from the contig lengths;
//   pass B  materialize_kernel  one warp per output sequence copies the rope segments / the original.
#include "pg_internal.cuh"
#include "pg_scan.cuh"
#include "pg_tables.h"

#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace pg {

struct ExConst {
    // CompareNuclResultByScore evaluates lgamma()/log() only at small integers.  For perfect-identity
    // overlaps its p equals beta1/(beta1+beta2) exactly, so the 0.45 / 0.55 cut-offs are hit as exact
    // rational ties (e.g. overlap lengths 98 vs 120) and the reference's answer then hangs on the last
    // ulps of the HOST libm.  The host therefore tabulates lgamma(n) and log(n) with its own libm (the
    // one the reference binary would call on this machine) and the kernel reads those values.
    const double *lgammaTab;
    const double *logTab;
    unsigned tabN;
    int nt;
    int alph;
    float seqIdThr;
    int maxSeqLen;
    int keepTarget;
    double lambda, logK;   // computeRawScoreFromBitScore (EvalueComputation.h:22-24)
};

__constant__ unsigned char c_ex_a2n[256];
__constant__ signed char c_ex_mat[21 * 21];
__constant__ unsigned char c_ex_revN[256];   // nt letter -> reverse-complement letter with X -> N (getRevFragment)

struct __align__(4) ExRes {      // Matcher::result_t subset + the per-target useReverse flag
    unsigned dbKey;
    int score;
    float seqId;
    unsigned alnLength;
    int qStartPos, qEndPos;
    unsigned qLen;
    int dbStartPos, dbEndPos;
    unsigned dbLen;
    unsigned rev;
};

struct ExSeg {                   // one rope segment: bytes [start, start+len) of sequence `src`, optionally
    unsigned src;                // reverse-complemented (the reversed fragment of that range)
    unsigned start;
    unsigned len;
    unsigned rev;
};

// CompareResultByScore (assembleresult.cpp:19-36)
__device__ __forceinline__ bool cmp_aa(const ExRes &r1, const ExRes &r2) {
    if (r1.score < r2.score) return true;
    if (r2.score < r1.score) return false;
    if (r1.alnLength < r2.alnLength) return true;
    if (r2.alnLength < r1.alnLength) return false;
    if (r1.dbKey > r2.dbKey) return true;
    if (r2.dbKey > r1.dbKey) return false;
    return false;
}
// CompareNuclResultByScore (nuclassembleresult.cpp:36-70)
__device__ __forceinline__ double tab_lgamma(const ExConst &c, unsigned long long n) { return n < c.tabN ? __ldg(c.lgammaTab + n) : lgamma((double) n); }
__device__ __forceinline__ double tab_log(const ExConst &c, unsigned long long n) { return n < c.tabN ? __ldg(c.logTab + n) : log((double) n); }
__device__ bool cmp_nt(const ExRes &r1, const ExRes &r2, const ExConst &c) {
    const unsigned mm1 = (unsigned) ((double) ((1.0f - r1.seqId) * (float) r1.alnLength) + 0.5);
    const unsigned mm2 = (unsigned) ((double) ((1.0f - r2.seqId) * (float) r2.alnLength) + 0.5);
    const unsigned alpha1 = mm1 + 1, alpha2 = mm2 + 1;
    const unsigned beta1 = r1.alnLength - mm1 + 1, beta2 = r2.alnLength - mm2 + 1;
    const double log_c = (tab_lgamma(c, beta1 + beta2) + tab_lgamma(c, alpha1 + beta1)) -
                         (tab_lgamma(c, alpha1 + beta1 + beta2) + tab_lgamma(c, beta1));
    double log_r = 0.0, p = 0.0;
    for (unsigned long long idx = 0; idx < alpha2; idx++) {
        p += exp(log_r + log_c);
        log_r = tab_log(c, alpha1 + idx) + tab_log(c, beta2 + idx) - (tab_log(c, idx + 1) + tab_log(c, idx + alpha1 + beta1 + beta2)) + log_r;
    }
    if (p < 0.45) return true;
    if (p > 0.55) return false;
    if (r1.dbLen - r1.alnLength < r2.dbLen - r2.alnLength) return true;
    if (r1.dbLen - r1.alnLength > r2.dbLen - r2.alnLength) return false;
    return true;
}
__device__ __forceinline__ bool cmp_res(const ExRes &a, const ExRes &b, const ExConst &c) { return c.nt ? cmp_nt(a, b, c) : cmp_aa(a, b); }

// libstdc++ std::__push_heap / std::__adjust_heap (bits/stl_heap.h), which std::priority_queue uses.
__device__ void heap_push_hole(ExRes *first, long hole, long top, const ExRes &value, const ExConst &nt) {
    long parent = (hole - 1) / 2;
    while (hole > top && cmp_res(first[parent], value, nt)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
__device__ void heap_push(ExRes *first, long &size, const ExRes &value, const ExConst &nt) {   // push_back + push_heap
    first[size] = value;
    size++;
    heap_push_hole(first, size - 1, 0, value, nt);
}
__device__ void heap_pop(ExRes *first, long &size, const ExConst &nt) {                        // pop_heap + pop_back
    if (size > 1) {
        const long len = size - 1;
        const ExRes value = first[len];
        first[len] = first[0];
        long hole = 0, second = 0;
        while (second < (len - 1) / 2) {
            second = 2 * (second + 1);
            if (cmp_res(first[second], first[second - 1], nt)) second--;
            first[hole] = first[second];
            hole = second;
        }
        if ((len & 1) == 0 && second == (len - 2) / 2) {
            second = 2 * (second + 1);
            first[hole] = first[second - 1];
            hole = second - 1;
        }
        heap_push_hole(first, hole, 0, value, nt);
    }
    size--;
}

struct Rope {
    ExSeg *segs;
    int n;
    unsigned len;
    const char *data;                       // the sequence DB (held by value: no pointer to the kernel parameter block)
    const unsigned long long *offsets;
    __device__ unsigned char at(unsigned i) const {
        int s = 0;
        while (i >= segs[s].len) { i -= segs[s].len; s++; }
        const ExSeg g = segs[s];
        const char *base = data + offsets[g.src];
        if (!g.rev) return (unsigned char) base[g.start + i];
        return c_ex_revN[(unsigned char) base[g.start + g.len - 1 - i]];
    }
};

// a target as it is compared during re-scoring: forward, or its full reverse complement (X -> N)
__device__ __forceinline__ unsigned char target_at(const char *t, unsigned tLen, unsigned rev, unsigned i) {
    return rev ? c_ex_revN[(unsigned char) t[tLen - 1 - i]] : (unsigned char) t[i];
}

// (float) strtod(text of Util::fastSeqIdToBuffer(seqId)): the value the assembler parses back from aln_N
__device__ __forceinline__ float seqid_text_roundtrip(float seqId) {
    if (seqId == 1.0f) return 1.0f;
    const int n = (int) __fmul_rn(seqId, 1000.0f);
    return (float) ((double) n / 1000.0);
}

constexpr unsigned EX_WARP_MAX_ALNS = 32;   // queries with more alignments take the heap path

// per-query state carried between the rounds of the wavefront
struct ExState {
    long long hsize;
    unsigned leftOff, rightOff;     // of the round that just ended (needed to re-score its parked hits)
    int nPark;
    int ropeN;
    unsigned ropeLen;
    unsigned querySeqLen;
    unsigned couldExtend;
};

__global__ void aln_ranges_kernel(const pg_seqdb db, const pg_aln *__restrict__ alns, unsigned long long nAlns,
                                  unsigned long long *__restrict__ alnStart, unsigned *__restrict__ alnCount,
                                  unsigned *__restrict__ activeList, unsigned *__restrict__ activeCount, unsigned *__restrict__ bigList) {
    const unsigned long long j = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nAlns) return;
    const unsigned qk = alns[j].query;
    if (j > 0 && alns[j - 1].query == qk) return;
    unsigned long long k = j;
    while (k < nAlns && alns[k].query == qk) k++;
    const unsigned qi = find_id(db.keys, (unsigned) db.n, qk);
    alnStart[qi] = j;
    alnCount[qi] = (unsigned) (k - j);
    if (k - j >= 2) activeList[atomicAdd(activeCount, 1u)] = qi;   // only the self alignment: nothing can be popped for extension
    if (k - j > EX_WARP_MAX_ALNS) bigList[atomicAdd(activeCount + 2, 1u)] = qi;   // queries that need the heap path even for amino acids
}

// the same from the per-sequence alignment counts rescorediagonal left behind (fused iteration): only the work lists
__global__ void active_from_counts_kernel(const unsigned *__restrict__ alnCount, unsigned long long n,
                                          unsigned *__restrict__ activeList, unsigned *__restrict__ activeCount, unsigned *__restrict__ bigList) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned c = i < n ? alnCount[i] : 0u;
    const unsigned lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, c >= 2), big = __ballot_sync(0xFFFFFFFFu, c > EX_WARP_MAX_ALNS);
    unsigned base = 0;
    if (lane == 0 && m) base = atomicAdd(activeCount, (unsigned) __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (c >= 2) activeList[base + __popc(m & ((1u << lane) - 1u))] = (unsigned) i;
    unsigned base2 = 0;
    if (lane == 0 && big) base2 = atomicAdd(activeCount + 2, (unsigned) __popc(big));
    base2 = __shfl_sync(0xFFFFFFFFu, base2, 0);
    if (c > EX_WARP_MAX_ALNS) bigList[base2 + __popc(big & ((1u << lane) - 1u))] = (unsigned) i;
}

__global__ void init_out_kernel(const pg_seqdb db, unsigned *__restrict__ outLen) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < db.n) outLen[i] = db.lens[i];
}

// One round of the reference's outer `while (!alnQueue.empty())` loop for every query of `list`:
// (first round) fill the queue, or (later rounds) push the re-scored parked hits that still pass --min-seq-id;
// then pop / extend / park until the queue is empty (assembleresult.cpp:193-289).  Queries whose parked hits
// must be re-scored append them to `work` and themselves to `nextList`; the others are finished.
__global__ void __launch_bounds__(128) extend_round_kernel(const pg_seqdb db, const pg_aln *__restrict__ alns,
                                                           const unsigned long long *__restrict__ alnStart, const unsigned *__restrict__ alnCount,
                                                           const ExConst c, int firstRound,
                                                           const unsigned *__restrict__ list, const unsigned *__restrict__ listCount,
                                                           unsigned *__restrict__ nextList, unsigned *__restrict__ nextCount,
                                                           uint2 *__restrict__ work, unsigned long long *__restrict__ workCount,
                                                           ExState *__restrict__ states, ExRes *__restrict__ heapBuf, ExRes *__restrict__ parkBuf,
                                                           ExSeg *__restrict__ segBuf, unsigned *__restrict__ segCount,
                                                           unsigned *__restrict__ outLen, unsigned char *__restrict__ extended,
                                                           unsigned char *__restrict__ used) {
    const unsigned li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= *listCount) return;
    const unsigned qi = list[li];
    const unsigned nAl = alnCount[qi];
    if (!c.nt && nAl <= EX_WARP_MAX_ALNS) return;           // amino acids: handled by extend_round_warp_kernel
    const unsigned long long a0 = alnStart[qi];
    ExRes *heap = heapBuf + a0;
    ExRes *park = parkBuf + a0;
    ExSeg *segs = segBuf + a0 + qi;      // capacity nAl + 1
    const unsigned queryKey = db.keys[qi];
    ExState st;
    Rope rope; rope.segs = segs; rope.data = db.data; rope.offsets = db.offsets;
    long hsize = 0;
    if (firstRound) {
        st.querySeqLen = db.lens[qi] - 2;
        segs[0].src = qi; segs[0].start = 0; segs[0].len = st.querySeqLen; segs[0].rev = 0;
        rope.n = 1; rope.len = st.querySeqLen;
        st.couldExtend = 0;
        // fill the queue (assembleresult.cpp:159-188 / nuclassembleresult.cpp:197-226)
        for (unsigned i = 0; i < nAl; i++) {
            const pg_aln a = alns[a0 + i];
            ExRes r;
            r.dbKey = a.target;
            r.seqId = seqid_text_roundtrip(a.seq_id);
            r.qStartPos = a.q_start; r.qEndPos = a.q_end; r.qLen = (unsigned) a.q_len;
            r.dbStartPos = a.db_start; r.dbEndPos = a.db_end; r.dbLen = (unsigned) a.db_len;
            const int adjQ = (r.qStartPos == -1) ? 0 : r.qStartPos;
            const int adjD = (r.dbStartPos == -1) ? 0 : r.dbStartPos;
            r.alnLength = (unsigned) (max(abs(r.qEndPos - adjQ), abs(r.dbEndPos - adjD)) + 1);   // Matcher.cpp:201-203
            const int rawScore = (int) (((c.logK + (double) a.bits * log(2.0)) / c.lambda) + 0.5);
            const float scorePerCol = __fdiv_rn((float) rawScore, (float) ((double) r.alnLength + 0.5));
            if (!c.nt) {
                const float alnLen = (float) r.alnLength;
                const float ids = __fmul_rn(r.seqId, alnLen);
                r.seqId = (float) ((double) ids / ((double) alnLen + 0.5));
            }
            r.score = (int) __fmul_rn(scorePerCol, 100.0f);
            r.rev = 0;
            if (c.nt) {
                if (r.qStartPos > r.qEndPos) {
                    r.rev = 1;
                    const int t = r.qStartPos; r.qStartPos = r.qEndPos; r.qEndPos = t;
                    const unsigned dbStartPos = (unsigned) r.dbStartPos;
                    r.dbStartPos = (int) (r.dbLen - (unsigned) r.dbEndPos - 1u);
                    r.dbEndPos = (int) (r.dbLen - dbStartPos - 1u);
                }
            }
            heap_push(heap, hsize, r, c);
        }
    } else {
        st = states[qi];
        rope.n = st.ropeN; rope.len = st.ropeLen;
        st.querySeqLen = rope.len;                               // querySeqLen = query.length() (assembleresult.cpp:291)
        // refill the queue with the re-scored parked hits (assembleresult.cpp:309-311), in parked order
        for (int ai = 0; ai < st.nPark; ai++) {
            const ExRes r = park[ai];
            if (r.seqId >= c.seqIdThr) heap_push(heap, hsize, r, c);
        }
    }
    const unsigned querySeqLen = st.querySeqLen;
    bool finished = true;
    if (hsize > 0) {
        unsigned leftOff = 0, rightOff = 0;
        int nPark = 0;
        while (true) {
            // selectFragmentToExtend (assembleresult.cpp:40-57)
            bool got = false;
            ExRes best;
            while (hsize > 0) {
                const ExRes res = heap[0];
                heap_pop(heap, hsize, c);
                const bool notRightStartAndLeftStart = !(res.dbStartPos == 0 && res.qStartPos == 0);
                const bool rightStart = res.dbStartPos == 0 && (res.dbEndPos != (int) res.dbLen - 1);
                const bool leftStart = res.qStartPos == 0 && (res.qEndPos != (int) res.qLen - 1);
                const bool isNotIdentity = (res.dbKey != queryKey);
                if ((rightStart || leftStart) && notRightStartAndLeftStart && isNotIdentity) { best = res; got = true; break; }
            }
            if (!got) break;
            const unsigned targetId = find_id_db(db, best.dbKey);
            const unsigned targetSeqLen = db.lens[targetId] - 2;
            if (best.dbStartPos == 0) {
                if ((targetSeqLen - (unsigned) (best.dbEndPos + 1)) <= rightOff) continue;
            } else if (best.qStartPos == 0) {
                if (best.dbStartPos <= (int) leftOff) continue;
            }
            const unsigned dbStartPos = (unsigned) best.dbStartPos, dbEndPos = (unsigned) best.dbEndPos;
            const unsigned qStartPos = (unsigned) best.qStartPos, qEndPos = (unsigned) best.qEndPos;
            if (dbStartPos == 0 && qEndPos == (querySeqLen - 1)) {            // right extension
                if (rightOff > 0) { park[nPark++] = best; continue; }
                const unsigned fragLen = targetSeqLen - (dbEndPos + 1);
                if (c.nt && (unsigned long long) rope.len + fragLen >= (unsigned long long) c.maxSeqLen) break;   // nucl only (:271-275)
                ExSeg g; g.src = targetId; g.len = fragLen; g.rev = best.rev;
                g.start = best.rev ? 0u : dbEndPos + 1;                       // reversed: rev(target[0, fragLen))
                segs[rope.n++] = g;
                rope.len += fragLen;
                rightOff += fragLen;
                used[targetId] = 1;
            } else if (qStartPos == 0 && dbEndPos == (targetSeqLen - 1)) {    // left extension
                if (leftOff > 0) { park[nPark++] = best; continue; }
                const unsigned fragLen = dbStartPos;
                if ((unsigned long long) rope.len + fragLen >= (unsigned long long) c.maxSeqLen) break;
                ExSeg g; g.src = targetId; g.len = fragLen; g.rev = best.rev;
                g.start = best.rev ? (targetSeqLen - dbStartPos) : 0u;        // reversed: rev(target[tLen-dbStart, tLen))
                for (int s = rope.n; s > 0; s--) segs[s] = segs[s - 1];
                segs[0] = g;
                rope.n++;
                rope.len += fragLen;
                leftOff += fragLen;
                used[targetId] = 1;
            }
        }
        if (leftOff > 0 || rightOff > 0) st.couldExtend = 1;
        // `if (!alnQueue.empty()) break;` ends the query; otherwise the parked hits are re-scored on the new contig
        if (hsize == 0 && nPark > 0) {
            finished = false;
            st.hsize = 0; st.leftOff = leftOff; st.rightOff = rightOff; st.nPark = nPark;
            st.ropeN = rope.n; st.ropeLen = rope.len;
            states[qi] = st;
            const unsigned long long w0 = atomicAdd(workCount, (unsigned long long) nPark);
            for (int i = 0; i < nPark; i++) work[w0 + i] = make_uint2(qi, (unsigned) i);
            nextList[atomicAdd(nextCount, 1u)] = qi;
        }
    }
    if (finished && st.couldExtend) {
        extended[qi] = 1;
        outLen[qi] = rope.len + 2;
        segCount[qi] = (unsigned) rope.n;
    }
}

// Amino-acid fast path of one round: one WARP per query, one lane per alignment (queries with at most 32
// alignments).  CompareResultByScore is a strict total order (score, alnLength, dbKey are never all equal for two
// hits of one query), so popping the priority queue == repeatedly taking the maximum of the remaining elements: a
// 5-step shuffle arg-max over registers instead of a chain of dependent heap loads.  Elements that fail the
// selectFragmentToExtend predicate would be popped and discarded by the reference, so their lanes simply start dead.
__global__ void __launch_bounds__(256) extend_round_warp_kernel(const pg_seqdb db, const pg_aln *__restrict__ alns,
                                                                const unsigned long long *__restrict__ alnStart, const unsigned *__restrict__ alnCount,
                                                                const ExConst c, int firstRound,
                                                                const unsigned *__restrict__ list, const unsigned *__restrict__ listCount,
                                                                unsigned *__restrict__ nextList, unsigned *__restrict__ nextCount,
                                                                uint2 *__restrict__ work, unsigned long long *__restrict__ workCount,
                                                                ExState *__restrict__ states, ExRes *__restrict__ parkBuf,
                                                                ExSeg *__restrict__ segBuf, unsigned *__restrict__ segCount,
                                                                unsigned *__restrict__ outLen, unsigned char *__restrict__ extended,
                                                                unsigned char *__restrict__ used) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned nList = *listCount;
    const unsigned warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (unsigned li = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); li < nList; li += warpsTotal) {
        const unsigned qi = list[li];
        const unsigned nAl = alnCount[qi];
        if (nAl > EX_WARP_MAX_ALNS) continue;                 // handled by extend_round_kernel
        const unsigned long long a0 = alnStart[qi];
        ExRes *park = parkBuf + a0;
        ExSeg *segs = segBuf + a0 + qi;
        const unsigned queryKey = db.keys[qi];
        ExState st;
        int ropeN; unsigned ropeLen;
        ExRes r; r.dbKey = 0; r.score = 0; r.seqId = 0; r.alnLength = 0; r.qStartPos = r.qEndPos = 0; r.qLen = 0; r.dbStartPos = r.dbEndPos = 0; r.dbLen = 0; r.rev = 0;
        bool alive = false;
        if (firstRound) {
            st.querySeqLen = db.lens[qi] - 2;
            if (lane == 0) { segs[0].src = qi; segs[0].start = 0; segs[0].len = st.querySeqLen; segs[0].rev = 0; }
            ropeN = 1; ropeLen = st.querySeqLen;
            st.couldExtend = 0;
            if (lane < nAl) {
                const pg_aln a = alns[a0 + lane];
                r.dbKey = a.target;
                r.seqId = seqid_text_roundtrip(a.seq_id);
                r.qStartPos = a.q_start; r.qEndPos = a.q_end; r.qLen = (unsigned) a.q_len;
                r.dbStartPos = a.db_start; r.dbEndPos = a.db_end; r.dbLen = (unsigned) a.db_len;
                const int adjQ = (r.qStartPos == -1) ? 0 : r.qStartPos;
                const int adjD = (r.dbStartPos == -1) ? 0 : r.dbStartPos;
                r.alnLength = (unsigned) (max(abs(r.qEndPos - adjQ), abs(r.dbEndPos - adjD)) + 1);
                const int rawScore = (int) (((c.logK + (double) a.bits * log(2.0)) / c.lambda) + 0.5);
                const float scorePerCol = __fdiv_rn((float) rawScore, (float) ((double) r.alnLength + 0.5));
                const float alnLen = (float) r.alnLength;
                const float ids = __fmul_rn(r.seqId, alnLen);
                r.seqId = (float) ((double) ids / ((double) alnLen + 0.5));
                r.score = (int) __fmul_rn(scorePerCol, 100.0f);
                alive = true;
            }
        } else {
            st = states[qi];
            ropeN = st.ropeN; ropeLen = st.ropeLen;
            st.querySeqLen = ropeLen;
            if ((int) lane < st.nPark) { r = park[lane]; alive = r.seqId >= c.seqIdThr; }
        }
        const unsigned querySeqLen = st.querySeqLen;
        const bool entered = alive;                           // this element is in the queue of this round
        // every lane resolves ITS target once (index + length): the pops below then need no dependent global loads
        unsigned myTargetId = 0, myTargetLen = 0;
        if (alive) { myTargetId = find_id_db(db, r.dbKey); myTargetLen = db.lens[myTargetId] - 2; }
        // selectFragmentToExtend's predicate (assembleresult.cpp:40-57): failing elements are popped and dropped
        if (alive) {
            const bool notRightStartAndLeftStart = !(r.dbStartPos == 0 && r.qStartPos == 0);
            const bool rightStart = r.dbStartPos == 0 && (r.dbEndPos != (int) r.dbLen - 1);
            const bool leftStart = r.qStartPos == 0 && (r.qEndPos != (int) r.qLen - 1);
            alive = (rightStart || leftStart) && notRightStartAndLeftStart && (r.dbKey != queryKey);
        }
        unsigned leftOff = 0, rightOff = 0;
        int nPark = 0;
        bool queueNotEmpty = false;   // set by the length-limit `break`: the queue keeps its lower-priority elements
        while (true) {
            const unsigned aliveMask = __ballot_sync(0xFFFFFFFFu, alive);
            if (aliveMask == 0) break;
            // arg-max by (score, alnLength, smaller dbKey)
            int bs = alive ? r.score : INT_MIN; unsigned bl = r.alnLength, bk = r.dbKey; int bLane = alive ? (int) lane : -1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int os = __shfl_xor_sync(0xFFFFFFFFu, bs, o);
                const unsigned ol = __shfl_xor_sync(0xFFFFFFFFu, bl, o), ok = __shfl_xor_sync(0xFFFFFFFFu, bk, o);
                const int oLane = __shfl_xor_sync(0xFFFFFFFFu, bLane, o);
                const bool take = (oLane >= 0) && (bLane < 0 || os > bs || (os == bs && (ol > bl || (ol == bl && ok < bk))));
                if (take) { bs = os; bl = ol; bk = ok; bLane = oLane; }
            }
            const int wl = bLane;                                  // the same in every lane
            if ((int) lane == wl) alive = false;                   // popped
            const int bDbStart = __shfl_sync(0xFFFFFFFFu, r.dbStartPos, wl), bDbEnd = __shfl_sync(0xFFFFFFFFu, r.dbEndPos, wl);
            const int bQStart = __shfl_sync(0xFFFFFFFFu, r.qStartPos, wl), bQEnd = __shfl_sync(0xFFFFFFFFu, r.qEndPos, wl);
            const unsigned targetId = __shfl_sync(0xFFFFFFFFu, myTargetId, wl);
            const unsigned targetSeqLen = __shfl_sync(0xFFFFFFFFu, myTargetLen, wl);
            if (bDbStart == 0) {
                if ((targetSeqLen - (unsigned) (bDbEnd + 1)) <= rightOff) continue;
            } else if (bQStart == 0) {
                if (bDbStart <= (int) leftOff) continue;
            }
            const unsigned dbStartPos = (unsigned) bDbStart, dbEndPos = (unsigned) bDbEnd;
            const unsigned qStartPos = (unsigned) bQStart, qEndPos = (unsigned) bQEnd;
            if (dbStartPos == 0 && qEndPos == (querySeqLen - 1)) {            // right extension
                if (rightOff > 0) { if ((int) lane == wl) park[nPark] = r; nPark++; continue; }
                const unsigned fragLen = targetSeqLen - (dbEndPos + 1);
                if (lane == 0) {
                    ExSeg g; g.src = targetId; g.len = fragLen; g.rev = 0; g.start = dbEndPos + 1;
                    segs[ropeN] = g;
                    used[targetId] = 1;
                }
                ropeN++; ropeLen += fragLen; rightOff += fragLen;
            } else if (qStartPos == 0 && dbEndPos == (targetSeqLen - 1)) {    // left extension
                if (leftOff > 0) { if ((int) lane == wl) park[nPark] = r; nPark++; continue; }
                const unsigned fragLen = dbStartPos;
                if ((unsigned long long) ropeLen + fragLen >= (unsigned long long) c.maxSeqLen) {
                    // `break` (assembleresult.cpp:258-262): everything with a lower priority than this hit -- selectable
                    // or not -- is still in the reference's queue, and a non-empty queue ends the query (:287-288)
                    const bool lower = entered && (int) lane != wl &&
                                       (r.score < bs || (r.score == bs && (r.alnLength < bl || (r.alnLength == bl && r.dbKey > bk))));
                    queueNotEmpty = __ballot_sync(0xFFFFFFFFu, lower) != 0;
                    break;
                }
                if (lane == 0) {
                    ExSeg g; g.src = targetId; g.len = fragLen; g.rev = 0; g.start = 0;
                    for (int sI = ropeN; sI > 0; sI--) segs[sI] = segs[sI - 1];
                    segs[0] = g;
                    used[targetId] = 1;
                }
                ropeN++; ropeLen += fragLen; leftOff += fragLen;
            }
        }
        if (leftOff > 0 || rightOff > 0) st.couldExtend = 1;
        bool finished = true;
        __syncwarp();
        if (!queueNotEmpty && nPark > 0) {
            finished = false;
            if (lane == 0) {
                st.hsize = 0; st.leftOff = leftOff; st.rightOff = rightOff; st.nPark = nPark; st.ropeN = ropeN; st.ropeLen = ropeLen;
                states[qi] = st;
                const unsigned long long w0 = atomicAdd(workCount, (unsigned long long) nPark);
                for (int i = 0; i < nPark; i++) work[w0 + i] = make_uint2(qi, (unsigned) i);
                nextList[atomicAdd(nextCount, 1u)] = qi;
            }
        }
        if (finished && st.couldExtend && lane == 0) {
            extended[qi] = 1;
            outLen[qi] = ropeLen + 2;
            segCount[qi] = (unsigned) ropeN;
        }
        __syncwarp();
    }
}

// Amino acids, queries with at most 32 alignments: the WHOLE query in one warp -- every round of the reference's outer
// loop (assembleresult.cpp:193-313) without leaving the kernel; one lane per alignment.
//  * CompareResultByScore is a strict total order (score, alnLength, smaller dbKey; never all equal for two hits of one
//    query), so popping the priority queue == repeatedly taking the maximum of the remaining elements.  The priority
//    is packed into two 32-bit words (biased score | alnLength << 5 | rank of the dbKey among the query's hits) and
//    a pop is two warp REDUX.MAX + one ballot instead of a heap operation.  Elements that fail the
//    selectFragmentToExtend predicate would be popped and discarded by the reference: their lanes start dead.
//  * The growing contig (a rope of segments) lives in shared memory while the query is processed and is written to
//    the segment buffer once at the end.
//  * A parked hit stays in the registers of its lane; when the queue has drained the same warp re-scores the
//    parked hits on the new contig (lane-parallel operand fetch, then the warp sums the diagonals one after the
//    other) and starts the next round with them.  No per-round state, work lists or host synchronisation.
struct WarpRope {                 // shared-memory rope of one warp (amino acids: never reversed)
    const char *base[EX_WARP_MAX_ALNS + 1];   // first byte of the segment
    unsigned len[EX_WARP_MAX_ALNS + 1];
    __device__ __forceinline__ unsigned char at(unsigned i) const {
        int s = 0;
        while (i >= len[s]) { i -= len[s]; s++; }
        return (unsigned char) base[s][i];
    }
};

__global__ void __launch_bounds__(256) extend_query_warp_kernel(const pg_seqdb db, const pg_aln *__restrict__ alns,
                                                                const unsigned long long *__restrict__ alnStart, const unsigned *__restrict__ alnCount,
                                                                const ExConst c, const unsigned *__restrict__ list, const unsigned *__restrict__ listCount,
                                                                ExSeg *__restrict__ segBuf, unsigned *__restrict__ segCount,
                                                                unsigned *__restrict__ outLen, unsigned char *__restrict__ extended,
                                                                unsigned char *__restrict__ used) {
    __shared__ unsigned char sA2n[256];
    __shared__ signed char sMat[21 * 21];
    __shared__ WarpRope sRope[8];
    __shared__ ExSeg sSegs[8][EX_WARP_MAX_ALNS + 1];          // the same rope as (source, start, length) for the output
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sA2n[i] = c_ex_a2n[i];
    for (int i = threadIdx.x; i < 21 * 21; i += blockDim.x) sMat[i] = c_ex_mat[i];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpRope &rope = sRope[w];
    ExSeg *segs = sSegs[w];
    const unsigned nList = *listCount;
    const unsigned warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (unsigned li = blockIdx.x * (blockDim.x >> 5) + w; li < nList; li += warpsTotal) {
        const unsigned qi = list[li];
        const unsigned nAl = alnCount[qi];
        if (nAl > EX_WARP_MAX_ALNS) continue;                 // handled by the round-based heap path
        const unsigned long long a0 = alnStart[qi];
        const unsigned queryKey = db.keys[qi];
        unsigned querySeqLen = db.lens[qi] - 2;
        if (lane == 0) {
            rope.base[0] = db.data + db.offsets[qi]; rope.len[0] = querySeqLen;
            segs[0].src = qi; segs[0].start = 0; segs[0].len = querySeqLen; segs[0].rev = 0;
        }
        int ropeN = 1; unsigned ropeLen = querySeqLen;
        bool couldExtend = false;
        ExRes r; r.dbKey = 0; r.score = 0; r.seqId = 0; r.alnLength = 0; r.qStartPos = r.qEndPos = 0; r.qLen = 0; r.dbStartPos = r.dbEndPos = 0; r.dbLen = 0; r.rev = 0;
        bool alive = false;
        if (lane < nAl) {
            const pg_aln a = alns[a0 + lane];
            r.dbKey = a.target;
            r.seqId = seqid_text_roundtrip(a.seq_id);
            r.qStartPos = a.q_start; r.qEndPos = a.q_end; r.qLen = (unsigned) a.q_len;
            r.dbStartPos = a.db_start; r.dbEndPos = a.db_end; r.dbLen = (unsigned) a.db_len;
            const int adjQ = (r.qStartPos == -1) ? 0 : r.qStartPos;
            const int adjD = (r.dbStartPos == -1) ? 0 : r.dbStartPos;
            r.alnLength = (unsigned) (max(abs(r.qEndPos - adjQ), abs(r.dbEndPos - adjD)) + 1);
            const int rawScore = (int) (((c.logK + (double) a.bits * log(2.0)) / c.lambda) + 0.5);
            const float scorePerCol = __fdiv_rn((float) rawScore, (float) ((double) r.alnLength + 0.5));
            const float alnLen = (float) r.alnLength;
            const float ids = __fmul_rn(r.seqId, alnLen);
            r.seqId = (float) ((double) ids / ((double) alnLen + 0.5));
            r.score = (int) __fmul_rn(scorePerCol, 100.0f);
            alive = true;
        }
        // tie-break of the comparator: the smaller dbKey wins => rank = number of hits of this query with a larger key
        unsigned keyRank = 0;
        for (unsigned j = 0; j < nAl; j++) keyRank += (__shfl_sync(0xFFFFFFFFu, r.dbKey, j) > r.dbKey) ? 1u : 0u;
        // the target of a hit is only touched if the hit can be selected at all: resolved lazily, once
        // (its length is the alignment record's dbLen: rescorediagonal wrote db_len = sequence length)
        const unsigned myTargetLen = r.dbLen;
        unsigned myTargetId = 0xFFFFFFFFu;
        const char *myTargetSeq = nullptr;
        while (true) {                                        // one iteration = one round of the reference's outer loop
            const bool entered = alive;                       // this element is in the queue of this round
            // selectFragmentToExtend's predicate (assembleresult.cpp:40-57): failing elements are popped and dropped
            if (alive) {
                const bool notRightStartAndLeftStart = !(r.dbStartPos == 0 && r.qStartPos == 0);
                const bool rightStart = r.dbStartPos == 0 && (r.dbEndPos != (int) r.dbLen - 1);
                const bool leftStart = r.qStartPos == 0 && (r.qEndPos != (int) r.qLen - 1);
                alive = (rightStart || leftStart) && notRightStartAndLeftStart && (r.dbKey != queryKey);
            }
            if (alive && myTargetId == 0xFFFFFFFFu) {
                myTargetId = find_id_db(db, r.dbKey);
                myTargetSeq = db.data + db.offsets[myTargetId];
            }
            // packed priority: (score, alnLength, smaller dbKey)
            const unsigned prHi = (unsigned) r.score ^ 0x80000000u;
            const unsigned prLo = (r.alnLength << 5) | keyRank;
            unsigned leftOff = 0, rightOff = 0;
            bool parked = false;
            int nPark = 0;
            bool queueNotEmpty = false;   // set by the length-limit `break`: the queue keeps its lower-priority elements
            while (true) {
                const unsigned aliveMask = __ballot_sync(0xFFFFFFFFu, alive);
                if (aliveMask == 0) break;
                const unsigned maxHi = __reduce_max_sync(0xFFFFFFFFu, alive ? prHi : 0u);
                const bool top = alive && prHi == maxHi;
                const unsigned maxLo = __reduce_max_sync(0xFFFFFFFFu, top ? prLo : 0u);
                const int wl = __ffs(__ballot_sync(0xFFFFFFFFu, top && prLo == maxLo)) - 1;   // the same in every lane
                if ((int) lane == wl) alive = false;                   // popped
                const int bDbStart = __shfl_sync(0xFFFFFFFFu, r.dbStartPos, wl), bDbEnd = __shfl_sync(0xFFFFFFFFu, r.dbEndPos, wl);
                const int bQStart = __shfl_sync(0xFFFFFFFFu, r.qStartPos, wl), bQEnd = __shfl_sync(0xFFFFFFFFu, r.qEndPos, wl);
                const unsigned targetSeqLen = __shfl_sync(0xFFFFFFFFu, myTargetLen, wl);
                if (bDbStart == 0) {
                    if ((targetSeqLen - (unsigned) (bDbEnd + 1)) <= rightOff) continue;
                } else if (bQStart == 0) {
                    if (bDbStart <= (int) leftOff) continue;
                }
                const unsigned dbStartPos = (unsigned) bDbStart, dbEndPos = (unsigned) bDbEnd;
                const unsigned qStartPos = (unsigned) bQStart, qEndPos = (unsigned) bQEnd;
                if (dbStartPos == 0 && qEndPos == (querySeqLen - 1)) {            // right extension
                    if (rightOff > 0) { if ((int) lane == wl) parked = true; nPark++; continue; }
                    const unsigned fragLen = targetSeqLen - (dbEndPos + 1);
                    if ((int) lane == wl) {
                        rope.base[ropeN] = myTargetSeq + dbEndPos + 1; rope.len[ropeN] = fragLen;
                        ExSeg g; g.src = myTargetId; g.len = fragLen; g.rev = 0; g.start = dbEndPos + 1;
                        segs[ropeN] = g;
                        used[myTargetId] = 1;
                    }
                    ropeN++; ropeLen += fragLen; rightOff += fragLen;
                } else if (qStartPos == 0 && dbEndPos == (targetSeqLen - 1)) {    // left extension
                    if (leftOff > 0) { if ((int) lane == wl) parked = true; nPark++; continue; }
                    const unsigned fragLen = dbStartPos;
                    if ((unsigned long long) ropeLen + fragLen >= (unsigned long long) c.maxSeqLen) {
                        // `break` (assembleresult.cpp:258-262): everything with a lower priority than this hit -- selectable
                        // or not -- is still in the reference's queue, and a non-empty queue ends the query (:287-288)
                        const bool lower = entered && (int) lane != wl && (prHi < maxHi || (prHi == maxHi && prLo < maxLo));
                        queueNotEmpty = __ballot_sync(0xFFFFFFFFu, lower) != 0;
                        break;
                    }
                    if ((int) lane == wl) {
                        for (int sI = ropeN; sI > 0; sI--) { rope.base[sI] = rope.base[sI - 1]; rope.len[sI] = rope.len[sI - 1]; segs[sI] = segs[sI - 1]; }
                        rope.base[0] = myTargetSeq; rope.len[0] = fragLen;
                        ExSeg g; g.src = myTargetId; g.len = fragLen; g.rev = 0; g.start = 0;
                        segs[0] = g;
                        used[myTargetId] = 1;
                    }
                    ropeN++; ropeLen += fragLen; leftOff += fragLen;
                }
                __syncwarp();                                 // the rope is edited by the popped lane, one edit at a time
            }
            if (leftOff > 0 || rightOff > 0) couldExtend = true;
            __syncwarp();                                     // the rope in shared memory is visible to the warp
            if (queueNotEmpty || nPark == 0) break;
            // ---- re-score the parked hits on the new contig (assembleresult.cpp:293-307, updateAlignment :70-108)
            int diag = 0;
            unsigned qOff = 0, tOff = 0, len = 0, first = 0, last = 0;
            bool valid = false;
            if (parked) {
                diag = (int) ((unsigned) r.qStartPos + leftOff) - r.dbStartPos;
                const unsigned dist = (unsigned) abs(diag);
                if (diag >= 0 && dist < ropeLen) { len = min(myTargetLen, ropeLen - dist); qOff = dist; valid = len > 0; }
                else if (diag < 0 && dist < myTargetLen) { len = min(myTargetLen - dist, ropeLen); tOff = dist; valid = len > 0; }
                if (valid) {
                    // the four end residues are fetched together (a short-circuited `||` would expose their latencies one by one)
                    const unsigned char q0 = rope.at(qOff), t0 = (unsigned char) myTargetSeq[tOff];
                    const unsigned char qE = rope.at(qOff + len - 1), tE = (unsigned char) myTargetSeq[tOff + len - 1];
                    first = (q0 == '*' || t0 == '*') ? 1u : 0u;
                    last = len - 1;
                    if (last > 0 && (qE == '*' || tE == '*')) last--;
                }
            }
            long long mySum = 0; int myIds = 0;
            unsigned todo = __ballot_sync(0xFFFFFFFFu, valid);
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const char *t = (const char *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) myTargetSeq, j);
                const unsigned qo = __shfl_sync(0xFFFFFFFFu, qOff, j), to = __shfl_sync(0xFFFFFFFFu, tOff, j);
                const unsigned f = __shfl_sync(0xFFFFFFFFu, first, j), l = __shfl_sync(0xFFFFFFFFu, last, j);
                int sum = 0, ids = 0;
                // score over [first, last]; identities over [qS, qE) = columns [first, last)  (exclusive end of updateAlignment)
                for (unsigned pos = f + lane; pos <= l; pos += 32) {
                    const unsigned char a = rope.at(qo + pos), b = (unsigned char) t[to + pos];
                    sum += sMat[sA2n[a] * c.alph + sA2n[b]];
                    if (pos < l) ids += (a == b) ? 1 : 0;
                }
                sum = __reduce_add_sync(0xFFFFFFFFu, sum); ids = __reduce_add_sync(0xFFFFFFFFu, ids);
                if ((int) lane == j) { mySum = sum; myIds = ids; }
            }
            alive = false;
            if (parked) {
                int start = -1, end = -1; unsigned score = 0, diagLen = 0; int idCnt = 0;
                const unsigned dist = (unsigned) abs(diag);
                if ((diag >= 0 && dist < ropeLen) || (diag < 0 && dist < myTargetLen)) diagLen = len;
                if (valid) {
                    if (mySum < 0) mySum = 0;
                    start = (int) first; end = (int) last; score = (unsigned) mySum; idCnt = myIds;
                }
                const int d2 = max(abs(diag), 0);
                int qS, qE, dS, dE;
                if (diag >= 0) { qS = start + d2; qE = end + d2; dS = start; dE = end; }
                else { qS = start; qE = end; dS = start + d2; dE = end + d2; }
                r.seqId = __fdiv_rn((float) idCnt, __fsub_rn((float) qE, (float) qS));
                r.qLen = ropeLen; r.dbLen = myTargetLen;
                r.alnLength = diagLen;
                const float scorePerCol = __fdiv_rn((float) score, (float) ((double) r.alnLength + 0.5));
                r.score = (int) __fmul_rn(scorePerCol, 100.0f);
                r.qStartPos = qS; r.qEndPos = qE; r.dbStartPos = dS; r.dbEndPos = dE;
                alive = r.seqId >= c.seqIdThr;                // re-queued only if it still passes --min-seq-id (:309-311)
            }
            querySeqLen = ropeLen;                            // querySeqLen = query.length() (assembleresult.cpp:291)
        }
        if (couldExtend) {
            ExSeg *gsegs = segBuf + a0 + qi;                  // capacity nAl + 1
            if ((int) lane < ropeN) gsegs[lane] = segs[lane];
            if (lane == 0) {
                if (ropeN > 32) gsegs[32] = segs[32];
                extended[qi] = 1;
                outLen[qi] = ropeLen + 2;
                segCount[qi] = (unsigned) ropeN;
            }
        }
        __syncwarp();
    }
}

// Re-scoring of the parked alignments on the extended contigs (assembleresult.cpp:293-307).  A warp takes 32
// parked hits: every lane fetches the operands of ITS hit (state, rope, target, overlap geometry -- a chain of
// dependent loads that now overlaps across the lanes), the warp then sums the 32 diagonals one after the other with
// all lanes striding the columns, and every lane finishes its own hit (updateAlignment, :70-108).
__global__ void __launch_bounds__(256) extend_rescore_kernel(const pg_seqdb db, const unsigned long long *__restrict__ alnStart,
                                                             const ExConst c, const uint2 *__restrict__ work,
                                                             const unsigned long long *__restrict__ workCount,
                                                             const ExState *__restrict__ states, ExRes *__restrict__ parkBuf,
                                                             ExSeg *__restrict__ segBuf) {
    __shared__ unsigned char sA2n[256];
    __shared__ signed char sMat[21 * 21];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sA2n[i] = c_ex_a2n[i];
    for (int i = threadIdx.x; i < 21 * 21; i += blockDim.x) sMat[i] = c_ex_mat[i];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long nWork = *workCount;
    const unsigned long long nBatches = (nWork + 31) / 32;
    const unsigned long long warpsTotal = (unsigned long long) gridDim.x * (blockDim.x >> 5);
    for (unsigned long long batch = (unsigned long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); batch < nBatches; batch += warpsTotal) {
        const unsigned long long wi = batch * 32 + lane;
        const bool live = wi < nWork;
        // ---- phase 1: my hit
        ExRes r; r.dbKey = 0; r.score = 0; r.seqId = 0; r.alnLength = 0; r.qStartPos = r.qEndPos = 0; r.qLen = 0; r.dbStartPos = r.dbEndPos = 0; r.dbLen = 0; r.rev = 0;
        Rope rope; rope.segs = nullptr; rope.data = db.data; rope.offsets = db.offsets; rope.n = 0; rope.len = 0;
        const char *tSeq = nullptr; unsigned tLen = 0;
        unsigned long long slot = 0;
        int diag = 0;
        unsigned qOff = 0, tOff = 0, len = 0, first = 0, last = 0;
        bool valid = false;
        if (live) {
            const uint2 it = work[wi];
            const unsigned qi = it.x;
            const unsigned long long a0 = alnStart[qi];
            const ExState st = states[qi];
            rope.segs = segBuf + a0 + qi; rope.n = st.ropeN; rope.len = st.ropeLen;
            slot = a0 + it.y;
            r = parkBuf[slot];
            const unsigned tId = find_id_db(db, r.dbKey);
            unsigned tEntry;
            tSeq = seq_entry(db, tId, &tEntry);
            tLen = tEntry - 2;
            diag = (int) ((unsigned) r.qStartPos + st.leftOff) - r.dbStartPos;
            const unsigned dist = (unsigned) abs(diag);
            if (diag >= 0 && dist < rope.len) { len = min(tLen, rope.len - dist); qOff = dist; valid = len > 0; }
            else if (diag < 0 && dist < tLen) { len = min(tLen - dist, rope.len); tOff = dist; valid = len > 0; }
            if (valid) {
                const unsigned char q0 = rope.at(qOff), t0 = target_at(tSeq, tLen, r.rev, tOff);
                const unsigned char qE = rope.at(qOff + len - 1), tE = target_at(tSeq, tLen, r.rev, tOff + len - 1);
                first = (q0 == '*' || t0 == '*') ? 1u : 0u;
                last = len - 1;
                if (last > 0 && (qE == '*' || tE == '*')) last--;
            }
        }
        // ---- phase 2: the warp sums the diagonals one by one
        long long mySum = 0; int myIds = 0;
        unsigned todo = __ballot_sync(0xFFFFFFFFu, valid);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            Rope q; q.data = db.data; q.offsets = db.offsets;
            q.segs = (ExSeg *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) rope.segs, j);
            q.n = __shfl_sync(0xFFFFFFFFu, rope.n, j);
            q.len = __shfl_sync(0xFFFFFFFFu, rope.len, j);
            const char *t = (const char *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) tSeq, j);
            const unsigned tl = __shfl_sync(0xFFFFFFFFu, tLen, j);
            const unsigned tRev = __shfl_sync(0xFFFFFFFFu, r.rev, j);
            const unsigned qo = __shfl_sync(0xFFFFFFFFu, qOff, j), to = __shfl_sync(0xFFFFFFFFu, tOff, j);
            const unsigned f = __shfl_sync(0xFFFFFFFFu, first, j), l = __shfl_sync(0xFFFFFFFFu, last, j);
            long long sum = 0; int ids = 0;
            // score over [first, last]; identities over [qS, qE) = columns [first, last)  (exclusive end of updateAlignment)
            for (unsigned pos = f + lane; pos <= l; pos += 32) {
                const unsigned char a = q.at(qo + pos), b = target_at(t, tl, tRev, to + pos);
                sum += sMat[sA2n[a] * c.alph + sA2n[b]];
                if (pos < l) ids += (a == b) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o); ids += __shfl_xor_sync(0xFFFFFFFFu, ids, o); }
            if ((int) lane == j) { mySum = sum; myIds = ids; }
        }
        // ---- phase 3: updateAlignment of my hit
        if (live) {
            int start = -1, end = -1; unsigned score = 0, diagLen = 0; int idCnt = 0;
            const unsigned dist = (unsigned) abs(diag);
            if ((diag >= 0 && dist < rope.len) || (diag < 0 && dist < tLen)) diagLen = len;
            if (valid) {
                if (mySum < 0) mySum = 0;
                start = (int) first; end = (int) last; score = (unsigned) mySum; idCnt = myIds;
            }
            const int d2 = max(abs(diag), 0);
            int qS, qE, dS, dE;
            if (diag >= 0) { qS = start + d2; qE = end + d2; dS = start; dE = end; }
            else { qS = start; qE = end; dS = start + d2; dE = end + d2; }
            r.seqId = __fdiv_rn((float) idCnt, __fsub_rn((float) qE, (float) qS));
            r.qLen = rope.len; r.dbLen = tLen;
            r.alnLength = diagLen;
            const float scorePerCol = __fdiv_rn((float) score, (float) ((double) r.alnLength + 0.5));
            r.score = (int) __fmul_rn(scorePerCol, 100.0f);
            r.qStartPos = qS; r.qEndPos = qE; r.dbStartPos = dS; r.dbEndPos = dE;
            parkBuf[slot] = r;
        }
    }
}

// keep[i] = 1 if sequence i is written to the output DB (assembleresult.cpp:316-342)
__global__ void keep_kernel(unsigned long long n, int keepTarget, const unsigned char *__restrict__ extended,
                            const unsigned char *__restrict__ used, unsigned *__restrict__ keep, unsigned *__restrict__ outLen,
                            const unsigned *__restrict__ keys, unsigned ownLo, unsigned ownHi) {
    const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned key = keys[i];
    const bool k = (key >= ownLo && key < ownHi) && (extended[i] || keepTarget || !used[i]);
    keep[i] = k ? 1u : 0u;
    if (!k) outLen[i] = 0;
}

// Output DB: a warp takes 32 consecutive sequences.  Every lane fetches the metadata of ITS sequence (coalesced loads,
// the dependent chain keep -> offset -> length overlaps across the lanes) and writes its index entry; the warp then
// copies the kept sequences one after the other: the original bytes, or the rope segments of a new contig.
__global__ void __launch_bounds__(256) materialize_kernel(const pg_seqdb db, const unsigned long long *__restrict__ alnStart,
                                                          const ExSeg *__restrict__ segBuf, const unsigned *__restrict__ segCount,
                                                          const unsigned *__restrict__ outLen, const unsigned long long *__restrict__ outOff,
                                                          const unsigned *__restrict__ keep, const unsigned long long *__restrict__ keepIdx,
                                                          const unsigned char *__restrict__ extended,
                                                          char *__restrict__ outData, unsigned long long *__restrict__ outOffsets,
                                                          unsigned *__restrict__ outLens, unsigned *__restrict__ outKeys,
                                                          unsigned char *__restrict__ outExtended) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long warpsTotal = (unsigned long long) gridDim.x * (blockDim.x >> 5);
    const unsigned long long nBatches = (db.n + 31) / 32;
    for (unsigned long long batch = (unsigned long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); batch < nBatches; batch += warpsTotal) {
        const unsigned long long qi = batch * 32 + lane;
        const bool live = qi < db.n && keep[qi] != 0;
        unsigned long long o = 0; unsigned len = 0, nseg = 0;
        const char *src = nullptr; const ExSeg *segs = nullptr;
        if (live) {
            o = outOff[qi]; len = outLen[qi]; nseg = segCount[qi];
            const unsigned long long slot = keepIdx[qi];
            outOffsets[slot] = o; outLens[slot] = len; outKeys[slot] = db.keys[qi]; outExtended[slot] = extended[qi];
            if (nseg == 0) src = db.data + db.offsets[qi];
            else segs = segBuf + alnStart[qi] + qi;
        }
        // unchanged sequences (the majority): four at a time, the loads of all four in flight before the first store -- one
        // sequence per step is one exposed memory latency per ~50 bytes
        unsigned plain = __ballot_sync(0xFFFFFFFFu, live && nseg == 0);
        while (plain) {
            const char *sp[4]; char *dp[4]; unsigned ll[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                sp[k] = nullptr; dp[k] = nullptr; ll[k] = 0;
                if (plain) {
                    const int j = __ffs(plain) - 1;
                    plain &= plain - 1;
                    sp[k] = (const char *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) src, j);
                    dp[k] = outData + __shfl_sync(0xFFFFFFFFu, o, j);
                    ll[k] = __shfl_sync(0xFFFFFFFFu, len, j);
                }
            }
            char a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                a[k] = 0; b[k] = 0;
                if (lane < ll[k]) a[k] = sp[k][lane];
                if (lane + 32 < ll[k]) b[k] = sp[k][lane + 32];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (lane < ll[k]) dp[k][lane] = a[k];
                if (lane + 32 < ll[k]) dp[k][lane + 32] = b[k];
                for (unsigned i = lane + 64; i < ll[k]; i += 32) dp[k][i] = sp[k][i];
            }
        }
        // new contigs: the rope segments one after the other
        unsigned todo = __ballot_sync(0xFFFFFFFFu, live && nseg != 0);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            char *dst = outData + __shfl_sync(0xFFFFFFFFu, o, j);
            const unsigned ns = __shfl_sync(0xFFFFFFFFu, nseg, j);
            const ExSeg *sg = (const ExSeg *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) segs, j);
            unsigned w = 0;
            for (unsigned s = 0; s < ns; s++) {
                const ExSeg g = sg[s];
                const char *sp = db.data + db.offsets[g.src];
                for (unsigned i = lane; i < g.len; i += 32)
                    dst[w + i] = g.rev ? (char) c_ex_revN[(unsigned char) sp[g.start + g.len - 1 - i]] : sp[g.start + i];
                w += g.len;
            }
            if (lane == 0) { dst[w] = '\n'; dst[w + 1] = '\0'; }
        }
    }
}

int ex_run(Context *ctx, const pg_seqdb *db, const pg_aln *d_alns, uint64_t nAlns, const pg_ex_params *p, pg_seqdb **outDb, unsigned char **d_extended) {
    cudaStream_t s = ctx->stream;
    PG_CHECK(p->rescore_mode == 3, "assembleresults: only --rescore-mode 3 (END_TO_END) is implemented on the GPU path");
    const bool nt = db->dbtype == PG_DBTYPE_NUCLEOTIDES;
    ExConst c;
    c.nt = nt; c.alph = nt ? 5 : 21; c.seqIdThr = p->seq_id_thr; c.maxSeqLen = p->max_seq_len; c.keepTarget = p->keep_target;
    const double *a = nt ? PG_NT_ALP : PG_AA_ALP;
    c.lambda = a[0]; c.logK = log(a[1]);
    unsigned char revN[256];
    for (int i = 0; i < 256; i++) { unsigned char r = PG_NT_NUM2AA[PG_NT_REVERSE[PG_NT_AA2NUM[i]]]; revN[i] = (r == 'X') ? 'N' : r; }
    signed char mat[21 * 21] = {0};
    if (nt) { for (int i = 0; i < 25; i++) mat[i] = PG_NT_SUBMAT[i]; } else { for (int i = 0; i < 441; i++) mat[i] = PG_AA_SUBMAT[i]; }
    PG_CUDA(cudaMemcpyToSymbolAsync(c_ex_a2n, nt ? PG_NT_AA2NUM : PG_AA_AA2NUM, 256, 0, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_ex_revN, revN, 256, 0, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_ex_mat, mat, sizeof(mat), 0, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaStreamSynchronize(s));

    c.lgammaTab = nullptr; c.logTab = nullptr; c.tabN = 0;
    if (nt) {
        const unsigned want = (unsigned) std::min<long long>(2LL * std::max(p->max_seq_len, (int) db->max_seq_len) + 64, 1 << 22);
        if (ctx->ntTabN < want) {
            std::vector<double> tab(2 * (size_t) want);
            tab[0] = 0.0; tab[want] = 0.0;
            for (unsigned i = 1; i < want; i++) { tab[i] = lgamma((double) i); tab[want + i] = log((double) i); }
            PG_TRY(ctx->ntTab.reserve(sizeof(double) * tab.size()));
            PG_CUDA(cudaMemcpyAsync(ctx->ntTab.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, s));
            PG_CUDA(cudaStreamSynchronize(s));
            ctx->ntTabN = want;
        }
        c.tabN = ctx->ntTabN;
        c.lgammaTab = ctx->ntTab.as<double>();
        c.logTab = ctx->ntTab.as<double>() + ctx->ntTabN;
    }
    const bool trace = getenv("PG_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[pg_trace] ex_run %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    cudaEventRecord(ctx->ev[EV_EX_BEGIN], s);
    const uint64_t n = db->n;
    // meta arrays (per sequence)
    const size_t a16 = 15;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + a16) & ~a16; return r; };
    const size_t oAlnStart = take(sizeof(unsigned long long) * (n + 1));
    const size_t oAlnCount = take(sizeof(unsigned) * (n + 1));
    const size_t oSegCount = take(sizeof(unsigned) * (n + 1));
    const size_t oOutLen = take(sizeof(unsigned) * (n + 1));
    const size_t oKeep = take(sizeof(unsigned) * (n + 1));
    const size_t oOutOff = take(sizeof(unsigned long long) * (n + 2));
    const size_t oKeepIdx = take(sizeof(unsigned long long) * (n + 2));
    const size_t oExt = take(n + 1);
    const size_t oUsed = take(n + 1);
    const size_t oScan = take(scan_workspace_bytes(n));
    PG_TRY(ctx->exMeta.reserve(o));
    unsigned char *m = ctx->exMeta.as<unsigned char>();
    unsigned long long *alnStart = (unsigned long long *) (m + oAlnStart);
    unsigned *alnCount = (unsigned *) (m + oAlnCount);
    unsigned *segCount = (unsigned *) (m + oSegCount);
    unsigned *outLen = (unsigned *) (m + oOutLen);
    unsigned *keep = (unsigned *) (m + oKeep);
    unsigned long long *outOff = (unsigned long long *) (m + oOutOff);
    unsigned long long *keepIdx = (unsigned long long *) (m + oKeepIdx);
    unsigned char *ext = m + oExt, *used = m + oUsed;
    void *scanWs = m + oScan;
    PG_CUDA(cudaMemsetAsync(m, 0, o, s));
    // the extension's work arrays (heaps, rope segments, work lists: ~120 B per alignment) live in the kmermatcher's two
    // record buffers when those are large enough: their records are dead by now, and everything runs on this stream or on
    // the auxiliary stream forked from it.  30 GB less at 50 M reads.
    const size_t workBytes = sizeof(ExRes) * 2 * (nAlns + 1);
    const size_t segBytes = (sizeof(ExSeg) * (nAlns + n + 1) + 255) & ~(size_t) 255;
    const size_t listBytes = sizeof(unsigned) * 3 * (n + 1) + sizeof(uint2) * (nAlns + 1) + sizeof(ExState) * (n + 1) + 256;
    const bool workInRecB = !ctx->noScratchAlias && ctx->recB.cap >= workBytes;
    const bool segsInRecA = !ctx->noScratchAlias && ctx->recA.cap >= segBytes + listBytes;
    if (!workInRecB) PG_TRY(ctx->exWork.reserve(workBytes));
    if (!segsInRecA) { PG_TRY(ctx->exSegs.reserve(segBytes)); PG_TRY(ctx->exLists.reserve(listBytes)); }
    ExRes *heapBuf = workInRecB ? ctx->recB.as<ExRes>() : ctx->exWork.as<ExRes>();
    ExRes *parkBuf = heapBuf + (nAlns + 1);
    ExSeg *segBuf = segsInRecA ? ctx->recA.as<ExSeg>() : ctx->exSegs.as<ExSeg>();
    // work lists of the wavefront
    unsigned char *lb = segsInRecA ? ctx->recA.as<unsigned char>() + segBytes : ctx->exLists.as<unsigned char>();
    ExState *states = (ExState *) lb;
    uint2 *work = (uint2 *) (lb + ((sizeof(ExState) * (n + 1) + 15) & ~(size_t) 15));
    unsigned *listA = (unsigned *) ((unsigned char *) work + ((sizeof(uint2) * (nAlns + 1) + 15) & ~(size_t) 15));
    unsigned *listB = listA + (n + 1);
    unsigned *listC = listB + (n + 1);
    unsigned long long *d_cnt = ctx->small.as<unsigned long long>() + 24;   // [24] work count, [25] list counts (2 x u32)
    unsigned *d_listCnt = (unsigned *) (d_cnt + 1);                          // [0],[1] list counters, [2] number of large queries
    PG_CUDA(cudaMemsetAsync(d_cnt, 0, 24, s));
    lap("setup/reserve/memset");
    init_out_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(*db, outLen);
    if (nAlns && d_alns == ctx->rsOut && ctx->rsCnt) {
        alnStart = const_cast<unsigned long long *>(ctx->rsOff);      // read-only from here on
        alnCount = const_cast<unsigned *>(ctx->rsCnt);
        active_from_counts_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(alnCount, n, listA, d_listCnt, listC);
    } else if (nAlns) aln_ranges_kernel<<<(unsigned) ((nAlns + 255) / 256), 256, 0, s>>>(*db, d_alns, nAlns, alnStart, alnCount, listA, d_listCnt, listC);
    ctx->launches += 2;
    unsigned hCnt[3] = {0, 0, 0};
    PG_TRY(read_back(ctx, hCnt, d_listCnt, 3 * sizeof(unsigned)));
    const bool needHeap = nt || hCnt[2] > 0;
    lap("init_out+aln_ranges");
    unsigned active = hCnt[0];
    // listA (count d_listCnt[0]) is the list of all active queries.  The one-kernel warp path reads it for its whole run, while
    // the rounds of the large queries proceed next to it on the auxiliary stream: the rounds therefore never write listA or
    // its count again -- they read it in round 0 and ping-pong between listB (count [1]) and listC (count [3]) afterwards.
    // (Round 1 used to recycle listA and zero its count: CTAs of the warp kernel that started after that point found an empty
    // list and silently skipped their queries -- only visible once the warp kernel runs in several waves AND a large query
    // needs a second round, i.e. from a few million reads on.)
    unsigned *cur = listA, *nxt = listB;
    int curIdx = 0, nxtIdx = 1;
    // amino acids: the queries with <= 32 alignments (all of them, usually) run to completion inside one kernel; only
    // the larger ones go through the rounds below
    const bool fusedWarp = !nt && getenv("PG_EX_WAVEFRONT") == nullptr;
    cudaStream_t rs = s;                                      // the stream the rounds run on
    if (fusedWarp && active > 0) {
        if (needHeap) {
            // the rounds of the large queries run on the auxiliary stream next to the warp kernel: disjoint queries
            // (`used` is only ever set to 1 by both)
            rs = ctx->auxStream;
            PG_CUDA(cudaEventRecord(ctx->evAuxFork, s));
            PG_CUDA(cudaStreamWaitEvent(rs, ctx->evAuxFork, 0));
        }
        extend_query_warp_kernel<<<std::min<unsigned>((active + 7) / 8, NUM_SMS * 32), 256, 0, s>>>(
            *db, d_alns, alnStart, alnCount, c, cur, d_listCnt + curIdx, segBuf, segCount, outLen, ext, used);
        ctx->launches++;
        if (!needHeap) active = 0;
        // the rounds start from the list of the large queries (listC, count [2]) that the setup kernel left, not from a sweep
        // over all active queries
        else { cur = listC; curIdx = 2; active = hCnt[2]; }
        lap("extend_query_warp");
    }
    for (int round = 0; active > 0; round++) {
        PG_CHECK(round < 100000, "assembleresults: extension did not converge");
        PG_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), rs));
        PG_CUDA(cudaMemsetAsync(d_listCnt + nxtIdx, 0, sizeof(unsigned), rs));
        if (!nt && !fusedWarp) {
            // amino acids, wavefront variant (PG_EX_WAVEFRONT=1): warp per query for the queries with <= 32 alignments,
            // heap replay for the rest (both kernels read the same list and skip the queries of the other class)
            extend_round_warp_kernel<<<std::min<unsigned>((active + 7) / 8, NUM_SMS * 32), 256, 0, rs>>>(
                *db, d_alns, alnStart, alnCount, c, round == 0, cur, d_listCnt + curIdx, nxt, d_listCnt + nxtIdx, work, d_cnt, states,
                parkBuf, segBuf, segCount, outLen, ext, used);
            ctx->launches++;
        }
        if (needHeap) {
            extend_round_kernel<<<(active + 127) / 128, 128, 0, rs>>>(*db, d_alns, alnStart, alnCount, c, round == 0, cur, d_listCnt + curIdx,
                                                                     nxt, d_listCnt + nxtIdx, work, d_cnt, states, heapBuf, parkBuf,
                                                                     segBuf, segCount, outLen, ext, used);
            ctx->launches++;
        }
        extend_rescore_kernel<<<NUM_SMS * 8, 256, 0, rs>>>(*db, alnStart, c, work, d_cnt, states, parkBuf, segBuf);
        ctx->launches += 1;
        // The list only shrinks from round to round and every kernel reads its length on the device, so the host needs the
        // length only to stop: it is read back after round 0 (the list drops from all queries to the large ones) and then every
        // 8th round -- up to 7 empty rounds at the end instead of one host round trip per round (the chain of rounds is as long
        // as the largest query's alignment list, whatever the number of GPUs).
        if (round == 0 || (round & 7) == 0) {
            PG_TRY(read_back_on(ctx, rs, hCnt, d_listCnt + nxtIdx, sizeof(unsigned)));
            active = hCnt[0];
        }
        cur = nxt; curIdx = nxtIdx;
        if (cur == listB) { nxt = listC; nxtIdx = 3; } else { nxt = listB; nxtIdx = 1; }
        if (trace && round < 3) lap("  round");
    }
    if (rs != s) {
        PG_CUDA(cudaEventRecord(ctx->evAuxJoin, rs));
        PG_CUDA(cudaStreamWaitEvent(s, ctx->evAuxJoin, 0));
    }
    lap("rounds (rest)");
    keep_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(n, c.keepTarget, ext, used, keep, outLen, db->keys, ctx->ownLo, ctx->ownHi);
    ctx->launches += 1;
    unsigned long long *d_tot = ctx->small.as<unsigned long long>() + 5;   // [5] bytes, [6] kept
    PG_TRY(exclusive_scan_u32(outLen, outOff, n, d_tot, scanWs, scan_workspace_bytes(n), s, &ctx->launches));
    PG_TRY(exclusive_scan_u32(keep, keepIdx, n, d_tot + 1, scanWs, scan_workspace_bytes(n), s, &ctx->launches));
    unsigned long long h[2] = {0, 0};
    PG_TRY(read_back(ctx, h, d_tot, sizeof(h)));
    PG_CUDA(cudaGetLastError());
    lap("keep + scans");
    pg_seqdb *out = new pg_seqdb();
    out->n = h[1]; out->data_bytes = h[0]; out->dbtype = db->dbtype;
    PG_CUDA(cudaMallocAsync(&out->data, h[0] + 16, s));
    PG_CUDA(cudaMallocAsync(&out->offsets, sizeof(unsigned long long) * (h[1] + 1), s));
    PG_CUDA(cudaMallocAsync(&out->lens, sizeof(unsigned) * (h[1] + 1), s));
    PG_CUDA(cudaMallocAsync(&out->keys, sizeof(unsigned) * (h[1] + 1), s));
    unsigned char *outExt = nullptr;
    PG_CUDA(cudaMallocAsync(&outExt, h[1] + 1, s));
    lap("cudaMallocAsync x5");
    materialize_kernel<<<NUM_SMS * 16, 256, 0, s>>>(*db, alnStart, segBuf, segCount, outLen, outOff, keep, keepIdx, ext,
                                                    out->data, out->offsets, out->lens, out->keys, outExt);
    ctx->launches++;
    cudaEventRecord(ctx->ev[EV_EX_END], s);
    PG_CUDA(cudaGetLastError());
    PG_TRY(seqdb_finalize(ctx, out));
    lap("materialize + finalize");
    *outDb = out;
    *d_extended = outExt;
    ctx->exRan = true;
    return 0;
}

}  // namespace pg
