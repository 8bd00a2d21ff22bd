// pg_rescore.cu -- GPU rescorediagonal, --rescore-mode 3 (END_TO_END).
//
// Replaces doRescorediagonal (reference lib/mmseqs/src/alignment/rescorediagonal.cpp:45-379) with
//   DistanceCalculator::computeUngappedAlignment / ungappedAlignmentByDiagonal /
//   computeGlobalSubstitutionStartEndDistance   (alignment/DistanceCalculator.h:94-113,115-175,204-220)
//   EvalueComputation::computeEvalue/computeBitScore (alignment/EvalueComputation.h:18-40) + ALP area
//   (lib/alp/sls_pvalues.cpp:366-525, sls_basic.hpp:195-198)
// One warp scores one prefilter line: the substitution table (21x21 or 5x5 int8) and the ASCII->code
// table are staged in shared memory, lanes stride over the diagonal, warp shuffles reduce.
// Compiled with -fmad=false so the double-precision E-value follows the CPU's operation order.
#include "pg_internal.cuh"
#include "pg_scan.cuh"
#include "pg_tables.h"

namespace pg {

struct RsConst {
    int nt;
    int alph;                 // 21 or 5
    float seqIdThr;
    double evalThr;
    int covMode;
    float covThr;
    int alnLenThr;
    int seqIdMode;
    double dbRes;             // getAminoAcidDBSize of the target DB
    unsigned ownLo, ownHi;    // multi-GPU: self lines only for the queries this rank owns (key range)
    unsigned selfLo, nSelf;   // ... = the index range [selfLo, selfLo + nSelf) of the key-sorted DB
    // ALP (Gumbel + finite size correction) parameters
    double lambda, K, a_I, b_I, a_J, b_J, alpha_I, beta_I, alpha_J, beta_J, sigma, tau, vi_thr, vj_thr, c_thr, logK;
};

__constant__ unsigned char c_rs_a2n[256];
__constant__ signed char c_rs_mat[21 * 21];
__constant__ unsigned char c_rs_rev[256];     // nt: letter -> reverse-complement letter (num2aa[reverse[aa2num]])

__device__ __forceinline__ double normal_probability(double x) { return 0.5 * erfc(-sqrt(0.5) * x); }

// pvalues::get_appr_tail_prob_with_cov_without_errors, area only (sls_pvalues.cpp:366-525);
// AlignmentEvaluer::area(score, seqlen1 = query, seqlen2 = db) passes (m_, n_) = (seqlen2, seqlen1).
__device__ double alp_area(const RsConst &p, double y, double seqlen1, double seqlen2) {
    const double const_val = 1.0 / sqrt(2.0 * 3.1415926535897932384626433832795);
    const double m_ = seqlen2, n_ = seqlen1;
    const double m_li_y = m_ - (p.a_I * y + p.b_I);
    const double vi_y = fmax(p.vi_thr, p.alpha_I * y + p.beta_I);
    const double sqrt_vi_y = sqrt(vi_y);
    const double m_F = (sqrt_vi_y == 0.0) ? 1e100 : m_li_y / sqrt_vi_y;
    const double P_m_F = normal_probability(m_F);
    const double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    const double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    const double n_lj_y = n_ - (p.a_J * y + p.b_J);
    const double vj_y = fmax(p.vj_thr, p.alpha_J * y + p.beta_J);
    const double sqrt_vj_y = sqrt(vj_y);
    const double n_F = (sqrt_vj_y == 0.0) ? 1e100 : n_lj_y / sqrt_vj_y;
    const double P_n_F = normal_probability(n_F);
    const double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    const double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
    const double c_y = fmax(p.c_thr, p.sigma * y + p.tau);
    return p1 * p2 + c_y * (P_m_F * P_n_F);
}

// SmithWaterman::computeCov (StripedSmithWaterman.cpp:1055-1057), unsigned arithmetic
__device__ __forceinline__ float compute_cov(unsigned s, unsigned e, unsigned len) {
    return (float) (min(len, max(s, e)) - min(s, e) + 1u) / (float) len;
}

__device__ __forceinline__ bool has_coverage(float covThr, int covMode, float qc, float tc) {   // Util.cpp:553-567
    switch (covMode) {
        case 0: return qc >= covThr && tc >= covThr;
        case 1: return qc >= covThr;
        case 2: return tc >= covThr;
        default: return true;
    }
}
__device__ __forceinline__ bool rs_can_be_covered(float covThr, int covMode, float q, float t) {   // Util.cpp:533-551
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

struct QView {             // the query as it is aligned: forward, or the full reverse complement
    const char *s;
    int len;
    bool rev;
    __device__ __forceinline__ unsigned char at(int i, const unsigned char *sRev) const {
        return rev ? sRev[(unsigned char) s[len - 1 - i]] : (unsigned char) s[i];
    }
};

struct DiagAln { int start, end; unsigned score, diagLen, dist; int diagonal; int idCnt; };

// ungappedAlignmentByDiagonal + computeGlobalSubstitutionStartEndDistance for one candidate diagonal, plus the
// identity count over the same columns (rescorediagonal.cpp:277-282, case-folded); executed by a full warp,
// result valid in all lanes.
__device__ DiagAln align_by_diagonal(const QView &q, const char *t, unsigned tLen, int diagonal, int alph,
                                     const unsigned char *sA2n, const signed char *sMat, const unsigned char *sRev) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned qLen = (unsigned) q.len;
    const unsigned dist = (unsigned) abs(diagonal);
    DiagAln r; r.start = -1; r.end = -1; r.score = 0; r.diagLen = 0; r.dist = dist; r.diagonal = diagonal; r.idCnt = 0;
    unsigned qOff, tOff, len;
    if (diagonal >= 0 && dist < qLen) { len = min(tLen, qLen - dist); qOff = dist; tOff = 0; }
    else if (diagonal < 0 && dist < tLen) { len = min(tLen - dist, qLen); qOff = 0; tOff = dist; }
    else return r;
    r.diagLen = len;
    if (len == 0) return r;
    const unsigned char q0 = q.at(qOff, sRev), t0 = (unsigned char) t[tOff];
    const unsigned char qE = q.at(qOff + len - 1, sRev), tE = (unsigned char) t[tOff + len - 1];
    const unsigned first = (q0 == '*' || t0 == '*') ? 1u : 0u;
    unsigned last = len - 1;
    if (last > 0 && (qE == '*' || tE == '*')) last--;
    int sum = 0, ids = 0;
    for (unsigned pos = first + lane; pos <= last; pos += 32) {
        const unsigned char qc = q.at(qOff + pos, sRev), tc = (unsigned char) t[tOff + pos];
        sum += sMat[sA2n[qc] * alph + sA2n[tc]];
        ids += ((qc & (unsigned char) ~0x20) == (tc & (unsigned char) ~0x20)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o); ids += __shfl_xor_sync(0xFFFFFFFFu, ids, o); }
    if (sum < 0) sum = 0;
    r.start = (int) first; r.end = (int) last; r.score = (unsigned) sum; r.idCnt = ids;
    return r;
}

// The same for ONE THREAD and a forward query: used for batches of short sequences (amino-acid fragments), where a
// warp striding ~50 columns is mostly idle.  Both sequences are streamed as aligned 32-bit words (funnel-shifted
// to the byte offset), four columns per pair of loads.  Integer sums: the result equals the warp version's.
__device__ DiagAln align_by_diagonal_thread(const char *qs, unsigned qLen, const char *t, unsigned tLen, int diagonal, int alph,
                                            const unsigned char *sA2n, const signed char *sMat, const signed char *sPair) {
    const unsigned dist = (unsigned) abs(diagonal);
    DiagAln r; r.start = -1; r.end = -1; r.score = 0; r.diagLen = 0; r.dist = dist; r.diagonal = diagonal; r.idCnt = 0;
    unsigned qOff, tOff, len;
    if (diagonal >= 0 && dist < qLen) { len = min(tLen, qLen - dist); qOff = dist; tOff = 0; }
    else if (diagonal < 0 && dist < tLen) { len = min(tLen - dist, qLen); qOff = 0; tOff = dist; }
    else return r;
    r.diagLen = len;
    if (len == 0) return r;
    const unsigned char q0 = (unsigned char) qs[qOff], t0 = (unsigned char) t[tOff];
    const unsigned char qE = (unsigned char) qs[qOff + len - 1], tE = (unsigned char) t[tOff + len - 1];
    const unsigned first = (q0 == '*' || t0 == '*') ? 1u : 0u;
    unsigned last = len - 1;
    if (last > 0 && (qE == '*' || tE == '*')) last--;
    int sum = 0, ids = 0;
    if (last >= first) {
        const unsigned n = last - first + 1;
        const unsigned long long qa0 = (unsigned long long) (qs + qOff + first), ta0 = (unsigned long long) (t + tOff + first);
        const unsigned *qw = reinterpret_cast<const unsigned *>(qa0 & ~3ULL), *tw = reinterpret_cast<const unsigned *>(ta0 & ~3ULL);
        const unsigned qsh = (unsigned) (qa0 & 3ULL) * 8u, tsh = (unsigned) (ta0 & 3ULL) * 8u;
        unsigned qPrev = __ldg(qw), tPrev = __ldg(tw);
        for (unsigned i = 0; i < n; i += 4) {
            const unsigned qNext = __ldg(qw + (i >> 2) + 1), tNext = __ldg(tw + (i >> 2) + 1);   // at most 7 bytes past the column range: inside the DB's tail slack
            unsigned q4 = __funnelshift_r(qPrev, qNext, qsh), t4 = __funnelshift_r(tPrev, tNext, tsh);
            qPrev = qNext; tPrev = tNext;
            const unsigned keep = (n - i < 4) ? 0xFFFFFFFFu >> (8 * (4 - (n - i))) : 0xFFFFFFFFu;     // the columns of this word inside the range
            // bytes >= 0x7E (never in a sequence DB; 0x7F / 0x7E is the padding pair below) take the checked path
            const bool plainAscii = (((q4 | t4 | (q4 + 0x02020202u) | (t4 + 0x02020202u)) & 0x80808080u) & keep) == 0;
            // last word: the columns past the range become the padding pair, which scores 0 and is not an identity
            q4 = (q4 & keep) | (0x7F7F7F7Fu & ~keep); t4 = (t4 & keep) | (0x7E7E7E7Eu & ~keep);
            // identities of four columns at once, case-folded (rescorediagonal.cpp:277-282)
            ids += __popc(__vcmpeq4(q4 & 0xDFDFDFDFu, t4 & 0xDFDFDFDFu)) >> 3;
            if (plainAscii) {                                 // one look-up per column in the pair table
#pragma unroll
                for (int b = 0; b < 4; b++) sum += sPair[((q4 >> (8 * b)) & 0x7Fu) * 128u + ((t4 >> (8 * b)) & 0x7Fu)];
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const unsigned qc = (q4 >> (8 * b)) & 0xFFu, tc = (t4 >> (8 * b)) & 0xFFu;
                    if (i + b < n) sum += sMat[sA2n[qc] * alph + sA2n[tc]];
                }
            }
        }
    }
    if (sum < 0) sum = 0;
    r.start = (int) first; r.end = (int) last; r.score = (unsigned) sum; r.idCnt = ids;
    return r;
}

constexpr int RS_THREAD_MAX_LEN = 512;   // batches whose sequences are all at most this long are scored one item per thread

// items [0, nHits): prefilter hit j;  items [nHits, nHits + n): the "key\t0\t0" self line of query (item - nHits).
// A warp takes 32 items: every lane first fetches the operands of ITS item (so the dependent loads
// hit -> key -> offset overlap across the lanes), the warp then scores the 32 diagonals one after the other with
// all lanes striding the columns, and finally every lane does the double-precision E-value / acceptance
// arithmetic of its own item -- the fp64 exp/erfc sequence is evaluated once per item instead of once per lane.
__global__ void __launch_bounds__(256) rescore_kernel(const pg_seqdb db, const pg_hit *__restrict__ hits, unsigned long long nHits,
                                                      const RsConst c, pg_aln *__restrict__ res, unsigned char *__restrict__ acc) {
    __shared__ unsigned char sA2n[256];
    __shared__ unsigned char sRev[256];
    __shared__ signed char sMat[21 * 21];
    __shared__ signed char sPair[128 * 128];        // score of an ASCII pair in one look-up (the reference's createAsciiSubMat, SubstitutionMatrix.h:56-73)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { sA2n[i] = c_rs_a2n[i]; sRev[i] = c_rs_rev[i]; }
    for (int i = threadIdx.x; i < 21 * 21; i += blockDim.x) sMat[i] = c_rs_mat[i];
    __syncthreads();
    // (0x7F, 0x7E) is the padding pair of a partial last word: score 0
    for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x)
        sPair[i] = (i == 0x7F * 128 + 0x7E) ? (signed char) 0 : sMat[sA2n[i >> 7] * c.alph + sA2n[i & 127]];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long nItems = nHits + c.nSelf;
    const unsigned long long nBatches = (nItems + 31) / 32;
    const unsigned long long warpsTotal = (unsigned long long) gridDim.x * (blockDim.x >> 5);
    for (unsigned long long batch = (unsigned long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); batch < nBatches; batch += warpsTotal) {
        const unsigned long long item = batch * 32 + lane;
        // ---- phase 1: operands of my item
        unsigned qKey = 0, tKey = 0, qi = 0, ti = 0; int prefScore = 0; unsigned short diag16 = 0;
        bool live = item < nItems;
        if (live) {
            if (item < nHits) {
                const pg_hit h = hits[item];
                qKey = h.rep; tKey = h.target; prefScore = h.score; diag16 = (unsigned short) (short) h.diag;
                qi = find_id_db(db, qKey);
                ti = find_id_db(db, tKey);
            } else {
                qi = ti = c.selfLo + (unsigned) (item - nHits);
                qKey = tKey = db.keys[qi];
            }
        }
        const char *qPtr = nullptr, *tPtr = nullptr; int qLen = 0, dbLen = 0;
        bool scoreIt = false;
        if (live) {
            unsigned ql, tl;
            qPtr = seq_entry(db, qi, &ql); qLen = (int) ql - 2;
            tPtr = seq_entry(db, ti, &tl); dbLen = (int) tl - 2;
            scoreIt = rs_can_be_covered(c.covThr, c.covMode, (float) qLen, (float) dbLen);
            if (!scoreIt) acc[item] = 0;                       // `continue` at rescorediagonal.cpp:214-216
        }
        // ---- phase 2: the warp scores the items one by one
        DiagAln mine; mine.start = -1; mine.end = -1; mine.score = 0; mine.diagLen = 0; mine.dist = 0; mine.diagonal = 0; mine.idCnt = 0;
        unsigned todo = __ballot_sync(0xFFFFFFFFu, scoreIt);
        // short forward sequences (amino-acid fragments): every lane scores its own item
        const bool threadPath = !c.nt && __all_sync(0xFFFFFFFFu, !scoreIt || (qLen <= RS_THREAD_MAX_LEN && dbLen <= RS_THREAD_MAX_LEN));
        if (threadPath) {
            todo = 0;
            if (scoreIt) {
                // computeUngappedAlignment: every diagonal congruent to diag16 modulo 65536, the strictly best one wins
                const unsigned tl = (unsigned) dbLen;
                for (unsigned d = 1; d <= 1 + tl / 32768; d++) {
                    const int real = (int) (0u - d * 65536u + diag16);
                    if (!(real < 0 && (unsigned) (-real) < tl) && !(real >= 0 && (unsigned) real < (unsigned) qLen)) continue;
                    const DiagAln tmp = align_by_diagonal_thread(qPtr, (unsigned) qLen, tPtr, tl, real, c.alph, sA2n, sMat, sPair);
                    if (tmp.score > mine.score) mine = tmp;
                }
                for (unsigned d = 0; d <= (unsigned) qLen / 65536; d++) {
                    const int real = (int) (d * 65536u + diag16);
                    if (!(real < 0 && (unsigned) (-real) < tl) && !(real >= 0 && (unsigned) real < (unsigned) qLen)) continue;
                    const DiagAln tmp = align_by_diagonal_thread(qPtr, (unsigned) qLen, tPtr, tl, real, c.alph, sA2n, sMat, sPair);
                    if (tmp.score > mine.score) mine = tmp;
                }
            }
        }
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            QView q;
            q.s = (const char *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) qPtr, j);
            const char *t = (const char *) __shfl_sync(0xFFFFFFFFu, (unsigned long long) tPtr, j);
            q.len = __shfl_sync(0xFFFFFFFFu, qLen, j);
            const unsigned tl = (unsigned) __shfl_sync(0xFFFFFFFFu, dbLen, j);
            const unsigned d16 = (unsigned) __shfl_sync(0xFFFFFFFFu, (int) diag16, j);
            q.rev = c.nt && (__shfl_sync(0xFFFFFFFFu, prefScore, j) < 0);
            // computeUngappedAlignment: try every diagonal congruent to diag16 modulo 65536, keep the strictly best
            DiagAln best; best.start = -1; best.end = -1; best.score = 0; best.diagLen = 0; best.dist = 0; best.diagonal = 0; best.idCnt = 0;
            for (unsigned d = 1; d <= 1 + tl / 32768; d++) {
                const int real = (int) (0u - d * 65536u + d16);
                if (!(real < 0 && (unsigned) (-real) < tl) && !(real >= 0 && (unsigned) real < (unsigned) q.len)) continue;   // score 0, never the best
                const DiagAln tmp = align_by_diagonal(q, t, tl, real, c.alph, sA2n, sMat, sRev);
                if (tmp.score > best.score) best = tmp;
            }
            for (unsigned d = 0; d <= (unsigned) q.len / 65536; d++) {
                const int real = (int) (d * 65536u + d16);
                if (!(real < 0 && (unsigned) (-real) < tl) && !(real >= 0 && (unsigned) real < (unsigned) q.len)) continue;
                const DiagAln tmp = align_by_diagonal(q, t, tl, real, c.alph, sA2n, sMat, sRev);
                if (tmp.score > best.score) best = tmp;
            }
            if ((int) lane == j) mine = best;
        }
        // ---- phase 3: E-value, identity, coverage, acceptance of my item
        if (scoreIt) {
            const bool isIdentity = (qi == ti);                 // same DB on both sides (rescorediagonal.cpp:205)
            const bool rev = c.nt && prefScore < 0;
            const int distance = (int) mine.score;
            const double epa = c.K * exp(-c.lambda * (double) distance);
            const double evalue = epa * alp_area(c, (double) distance, (double) qLen, c.dbRes);
            const int bitScore = (int) (((c.lambda * (double) distance - c.logK) / log(2.0)) + 0.5);
            const int alnLen = (mine.end - mine.start) + 1;
            int qS, qE, dS, dE;
            if (mine.diagonal >= 0) { qS = mine.start + (int) mine.dist; qE = mine.end + (int) mine.dist; dS = mine.start; dE = mine.end; }
            else { qS = mine.start; qE = mine.end; dS = mine.start + (int) mine.dist; dE = mine.end + (int) mine.dist; }
            float seqIdF = 0.0f;
            if (evalue <= c.evalThr || isIdentity) {
                // zero-score hit (start = end = -1): the reference compares the byte before both sequences; for the
                // identity hit both are the same byte => one identity
                const int idCnt = (qS < 0) ? 1 : mine.idCnt;
                if (c.seqIdMode == 1) seqIdF = __fdiv_rn((float) idCnt, (float) min(qLen, dbLen));
                else if (c.seqIdMode == 2) seqIdF = __fdiv_rn((float) idCnt, (float) max(qLen, dbLen));
                else seqIdF = __fdiv_rn((float) idCnt, (float) alnLen);
            }
            const double seqId = (double) seqIdF;
            const float queryCov = compute_cov((unsigned) qS, (unsigned) qE, (unsigned) qLen);
            const float targetCov = compute_cov((unsigned) dS, (unsigned) dE, (unsigned) dbLen);
            if (rev) { qS = qLen - qS - 1; qE = qLen - qE - 1; }
            const bool hasCov = has_coverage(c.covThr, c.covMode, queryCov, targetCov);
            const bool hasSeqId = seqId >= (double) (c.seqIdThr - 1.1920928955078125e-07f);   // FLT_EPSILON, float subtraction
            const bool hasEvalue = evalue <= c.evalThr;
            const bool hasAlnLen = alnLen >= c.alnLenThr;
            const bool accepted = isIdentity || (hasAlnLen && hasCov && hasSeqId && hasEvalue);
            if (accepted) {
                pg_aln out;
                out.query = qKey; out.target = tKey;
                out.bits = bitScore; out.seq_id = seqIdF; out.evalue = evalue;
                out.q_start = qS; out.q_end = qE; out.q_len = qLen; out.db_start = dS; out.db_end = dE; out.db_len = dbLen;
                res[item] = out;
            }
            acc[item] = accepted ? 1 : 0;
        }
    }
}

// number of alignment lines of every query: 1 (self) + accepted hits of its block
__global__ void count_per_query_kernel(const pg_seqdb db, const pg_hit *__restrict__ hits, unsigned long long nHits,
                                       const unsigned char *__restrict__ acc, unsigned *__restrict__ cnt) {
    const unsigned long long j = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nHits) return;
    const unsigned rep = hits[j].rep;
    if (j > 0 && hits[j - 1].rep == rep) return;
    unsigned c = 1;
    for (unsigned long long k = j; k < nHits && hits[k].rep == rep; k++) c += acc[k];
    cnt[find_id(db.keys, (unsigned) db.n, rep)] = c;
}

__global__ void fill_owned_kernel(const unsigned *__restrict__ keys, unsigned *p, unsigned long long n, unsigned ownLo, unsigned ownHi) {
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned k = keys[i];
        p[i] = (k >= ownLo && k < ownHi) ? 1u : 0u;
    }
}

// first index with key >= lo / key >= hi in the ascending key array
__global__ void key_lower_bound_kernel(const unsigned *__restrict__ keys, unsigned n, unsigned lo, unsigned hi, unsigned *__restrict__ out) {
    if (threadIdx.x < 2) {
        const unsigned want = threadIdx.x == 0 ? lo : hi;
        unsigned a = 0, b = n;
        while (a < b) { const unsigned m = (a + b) >> 1; if (keys[m] < want) a = m + 1; else b = m; }
        out[threadIdx.x] = a;
    }
}

__global__ void gather_self_kernel(const pg_aln *__restrict__ res, unsigned long long nHits, unsigned long long n,
                                   const unsigned long long *__restrict__ off, const unsigned char *__restrict__ acc, unsigned selfLo, pg_aln *__restrict__ out) {
    const unsigned long long k = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;     // k-th owned query
    if (k < n && acc[nHits + k]) out[off[selfLo + k]] = res[nHits + k];
}

__global__ void gather_hits_kernel(const pg_seqdb db, const pg_hit *__restrict__ hits, unsigned long long nHits,
                                   const unsigned char *__restrict__ acc, const pg_aln *__restrict__ res,
                                   const unsigned long long *__restrict__ off, pg_aln *__restrict__ out) {
    const unsigned long long j = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nHits) return;
    const unsigned rep = hits[j].rep;
    if (j > 0 && hits[j - 1].rep == rep) return;
    unsigned long long w = off[find_id(db.keys, (unsigned) db.n, rep)] + 1;
    for (unsigned long long k = j; k < nHits && hits[k].rep == rep; k++)
        if (acc[k]) out[w++] = res[k];
}

int rs_run(Context *ctx, const pg_seqdb *db, const pg_hit *d_hits, uint64_t nHits, const pg_rs_params *p, pg_aln **d_alns, uint64_t *nAlns) {
    cudaStream_t s = ctx->stream;
    PG_CHECK(p->rescore_mode == 3, "rescorediagonal: only --rescore-mode 3 (END_TO_END) is implemented on the GPU path");
    const bool nt = db->dbtype == PG_DBTYPE_NUCLEOTIDES;
    RsConst c;
    c.nt = nt; c.alph = nt ? 5 : 21;
    c.seqIdThr = p->seq_id_thr; c.evalThr = p->eval_thr; c.covMode = p->cov_mode; c.covThr = p->cov_thr;
    c.alnLenThr = p->aln_len_thr; c.seqIdMode = p->seq_id_mode; c.dbRes = db->residues;
    c.ownLo = ctx->ownLo; c.ownHi = ctx->ownHi;
    const double *a = nt ? PG_NT_ALP : PG_AA_ALP;
    c.lambda = a[0]; c.K = a[1]; c.a_I = a[2]; c.b_I = a[3]; c.a_J = a[4]; c.b_J = a[5];
    c.alpha_I = a[6]; c.beta_I = a[7]; c.alpha_J = a[8]; c.beta_J = a[9]; c.sigma = a[10]; c.tau = a[11];
    // pvalues::compute_tmp_values (sls_pvalues.cpp:342-364)
    c.vi_thr = std::max(2.0 * c.alpha_I / c.lambda, 0.0);
    c.vj_thr = std::max(2.0 * c.alpha_J / c.lambda, 0.0);
    c.c_thr = std::max(2.0 * c.sigma / c.lambda, 0.0);
    c.logK = log(c.K);
    unsigned char rev[256];
    for (int i = 0; i < 256; i++) rev[i] = PG_NT_NUM2AA[PG_NT_REVERSE[PG_NT_AA2NUM[i]]];
    signed char mat[21 * 21] = {0};
    if (nt) { for (int i = 0; i < 25; i++) mat[i] = PG_NT_SUBMAT[i]; } else { for (int i = 0; i < 441; i++) mat[i] = PG_AA_SUBMAT[i]; }
    PG_CUDA(cudaMemcpyToSymbolAsync(c_rs_a2n, nt ? PG_NT_AA2NUM : PG_AA_AA2NUM, 256, 0, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_rs_rev, rev, 256, 0, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_rs_mat, mat, sizeof(mat), 0, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaStreamSynchronize(s));   // host staging arrays above are stack memory

    // index range of the owned keys (keys are ascending): everything on one GPU
    unsigned selfLo = 0, selfHi = (unsigned) db->n;
    if (ctx->ownLo != 0 || ctx->ownHi != 0xFFFFFFFFu) {
        unsigned *d_lb = (unsigned *) (ctx->small.as<unsigned long long>() + 36);
        key_lower_bound_kernel<<<1, 32, 0, s>>>(db->keys, (unsigned) db->n, ctx->ownLo, ctx->ownHi, d_lb);
        unsigned h_lb[2] = {0, 0};
        PG_TRY(read_back(ctx, h_lb, d_lb, sizeof(h_lb)));
        selfLo = h_lb[0]; selfHi = h_lb[1];
    }
    c.selfLo = selfLo; c.nSelf = selfHi - selfLo;
    cudaEventRecord(ctx->ev[EV_RS_BEGIN], s);
    const uint64_t n = db->n, nItems = nHits + c.nSelf;
    // the staging array of all scored lines lives in the kmermatcher's first record buffer when that is large enough (its
    // records are dead once the hits exist; same stream, so the reuse is ordered): 12 GB less at 50 M reads
    const bool stageInRecA = !ctx->noScratchAlias && ctx->recA.cap >= sizeof(pg_aln) * (nItems + 1);
    if (!stageInRecA) PG_TRY(ctx->alnAll.reserve(sizeof(pg_aln) * (nItems + 1)));
    PG_TRY(ctx->flags.reserve(nItems + 16 + sizeof(unsigned) * (n + 1) + sizeof(unsigned long long) * (n + 2) + scan_workspace_bytes(n) + 64));
    pg_aln *res = stageInRecA ? ctx->recA.as<pg_aln>() : ctx->alnAll.as<pg_aln>();
    unsigned char *acc = ctx->flags.as<unsigned char>();
    size_t o = (nItems + 15) & ~(size_t) 15;
    unsigned *cnt = (unsigned *) (acc + o); o += (sizeof(unsigned) * (n + 1) + 15) & ~(size_t) 15;
    unsigned long long *off = (unsigned long long *) (acc + o); o += sizeof(unsigned long long) * (n + 2);
    void *scanWs = acc + o;
    unsigned long long *d_total = ctx->small.as<unsigned long long>() + 4;

    unsigned long long warps = (nItems + 31) / 32;
    unsigned blocks = (unsigned) std::min<unsigned long long>((warps + 7) / 8, (unsigned long long) NUM_SMS * 64);
    if (blocks == 0) blocks = 1;
    rescore_kernel<<<blocks, 256, 0, s>>>(*db, d_hits, nHits, c, res, acc);
    fill_owned_kernel<<<NUM_SMS * 4, 256, 0, s>>>(db->keys, cnt, n, c.ownLo, c.ownHi);
    if (nHits) count_per_query_kernel<<<(unsigned) ((nHits + 255) / 256), 256, 0, s>>>(*db, d_hits, nHits, acc, cnt);
    ctx->launches += 3;
    PG_TRY(exclusive_scan_u32(cnt, off, n, d_total, scanWs, scan_workspace_bytes(n), s, &ctx->launches));
    unsigned long long h = 0;
    PG_TRY(read_back(ctx, &h, d_total, sizeof(h)));
    PG_TRY(ctx->alns.reserve(sizeof(pg_aln) * (h + 1)));
    pg_aln *out = ctx->alns.as<pg_aln>();
    PG_CUDA(cudaStreamWaitEvent(s, ctx->evAlnsCopied, 0));   // an asynchronous copy of the previous call's alignments may still read the buffer
    if (c.nSelf) gather_self_kernel<<<(unsigned) ((c.nSelf + 255) / 256), 256, 0, s>>>(res, nHits, c.nSelf, off, acc, selfLo, out);
    if (nHits) gather_hits_kernel<<<(unsigned) ((nHits + 255) / 256), 256, 0, s>>>(*db, d_hits, nHits, acc, res, off, out);
    ctx->launches += 2;
    cudaEventRecord(ctx->ev[EV_RS_END], s);
    PG_CUDA(cudaGetLastError());
    *d_alns = out; *nAlns = h;
    ctx->rsOut = out; ctx->rsCnt = cnt; ctx->rsOff = off;
    ctx->timings.n_alns = h;
    ctx->rsRan = true;
    return 0;
}

}  // namespace pg
