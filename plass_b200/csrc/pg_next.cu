// pg_next.cu -- the two per-iteration helpers next to the hot path.
//
//   fs_run  replaces findassemblystart   src/assembler/findassemblystart.cpp:35-176   (plass STEP 0, data/assemble.sh:108-117)
//   cc_run  replaces cyclecheck          src/assembler/cyclecheck.cpp:71-274          (every penguin iteration, data/nuclassemble.sh:19-60)
//
// Both are byte / integer work over the sequence DB that is already resident in HBM: one warp (or one CTA for long
// sequences) per sequence, shared-memory staging, no tensor cores.
#include "pg_internal.cuh"
#include "pg_scan.cuh"
#include "pg_tables.h"

namespace pg {

// ================================================================================================================
// findassemblystart
// ================================================================================================================
// alignment ranges per QUERY sequence index: the alignments are ordered by query key
__global__ void fs_ranges_kernel(const pg_aln *__restrict__ alns, unsigned long long nAlns, const unsigned *__restrict__ keys, unsigned n,
                                 unsigned long long *__restrict__ start, unsigned long long *__restrict__ end) {
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < nAlns; i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned q = alns[i].query;
        const bool first = (i == 0) || alns[i - 1].query != q;
        const bool last = (i + 1 == nAlns) || alns[i + 1].query != q;
        if (first || last) {
            const unsigned idx = find_id(keys, n, q);
            if (idx != 0xFFFFFFFFu) {
                if (first) start[idx] = i;
                if (last) end[idx] = i + 1;
            }
        }
    }
}

// One warp per query: position of the first 'M' of the query (findPosOfM :12-23), the projected position in every
// aligned target (:100-112), the vote (:114-121) and, if at least 20 % of the members have "*M", the atomic max of the
// projected positions into addStop (:122-133).
__global__ void __launch_bounds__(256) fs_vote_kernel(const pg_seqdb db, const pg_aln *__restrict__ alns,
                                                      const unsigned long long *__restrict__ start, const unsigned long long *__restrict__ end,
                                                      int *__restrict__ addStop) {
    const int lane = threadIdx.x & 31;
    const unsigned n = (unsigned) db.n;
    const unsigned warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (unsigned qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); qi < n; qi += warpsTotal) {
        const unsigned long long s = start[qi], e = end[qi];
        if (e <= s) continue;                                    // no result entry for this key
        const char *q = db.data + db.offsets[qi];
        const int qBytes = (int) db.lens[qi] - 1;                // bytes before the entry's '\0'
        int posM = -1;
        for (int p0 = 0; p0 < qBytes && posM < 0; p0 += 32) {
            const int p = p0 + lane;
            const bool isM = p < qBytes && q[p] == 'M';
            const unsigned m = __ballot_sync(0xFFFFFFFFu, isM);
            if (m) posM = p0 + __ffs(m) - 1;
        }
        if (posM < 0) continue;
        const bool qStopM = posM > 0 && q[posM - 1] == '*';
        unsigned members = 1, stopM = qStopM ? 1u : 0u;          // the query itself (:91)
        for (unsigned long long a0 = s; a0 < e; a0 += 32) {
            const unsigned long long a = a0 + lane;
            bool member = false, hasStopM = false;
            if (a < e) {
                const pg_aln r = alns[a];
                const unsigned ti = find_id(db.keys, n, r.target);
                if (ti != 0xFFFFFFFFu && ti != qi) {
                    member = true;
                    if (r.q_start >= posM && posM <= r.q_end) {  // :103 (sic)
                        const int dbMPos = r.db_start + (posM - r.q_start);
                        const char *t = db.data + db.offsets[ti];
                        const bool hasM = dbMPos >= 0 && t[dbMPos] == 'M';
                        if (dbMPos > 0 && hasM) hasStopM = t[dbMPos - 1] == '*';
                    }
                }
            }
            members += __popc(__ballot_sync(0xFFFFFFFFu, member));
            stopM += __popc(__ballot_sync(0xFFFFFFFFu, hasStopM));
        }
        if (members <= 1) continue;
        const float frequency = (float) (int) stopM / (float) members;
        if (!(frequency >= 0.2f)) continue;
        if (lane == 0) atomicMax(&addStop[qi], posM);
        for (unsigned long long a0 = s; a0 < e; a0 += 32) {
            const unsigned long long a = a0 + lane;
            if (a < e) {
                const pg_aln r = alns[a];
                const unsigned ti = find_id(db.keys, n, r.target);
                if (ti != 0xFFFFFFFFu && ti != qi && r.q_start >= posM && posM <= r.q_end)
                    atomicMax(&addStop[ti], r.db_start + (posM - r.q_start));     // mPos = -1 entries never raise the maximum
            }
        }
    }
}

__global__ void fs_len_kernel(const unsigned *__restrict__ lens, const int *__restrict__ addStop, unsigned n, unsigned *__restrict__ outLen) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int m = addStop[i];
        outLen[i] = m < 0 ? lens[i] : lens[i] - (unsigned) m + 1u;      // "*" + residues[mPos..] + "\n\0" (:150-160)
    }
}

__global__ void __launch_bounds__(256) fs_write_kernel(const pg_seqdb db, const int *__restrict__ addStop, const unsigned *__restrict__ outLen,
                                                       const unsigned long long *__restrict__ outOff, char *__restrict__ oData,
                                                       unsigned long long *__restrict__ oOffsets, unsigned *__restrict__ oLens, unsigned *__restrict__ oKeys) {
    const int lane = threadIdx.x & 31;
    const unsigned n = (unsigned) db.n;
    const unsigned warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (unsigned i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warpsTotal) {
        const char *src = db.data + db.offsets[i];
        char *dst = oData + outOff[i];
        const int m = addStop[i];
        const unsigned len = outLen[i];
        if (m < 0) {
            for (unsigned b = lane; b < len; b += 32) dst[b] = src[b];
        } else {
            if (lane == 0) dst[0] = '*';
            for (unsigned b = lane; b + 1 < len; b += 32) dst[1 + b] = src[m + b];
        }
        if (lane == 0) { oOffsets[i] = outOff[i]; oLens[i] = len; oKeys[i] = db.keys[i]; }
    }
}

int fs_run(Context *ctx, const pg_seqdb *db, const pg_aln *d_alns, uint64_t nAlns, pg_seqdb **outDb, int **d_addStop) {
    cudaStream_t s = ctx->stream;
    PG_CHECK(db->dbtype == PG_DBTYPE_AMINO_ACIDS, "findassemblystart: amino-acid sequence DB expected");
    const unsigned n = (unsigned) db->n;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t) 15; return r; };
    const size_t oStart = take(sizeof(unsigned long long) * ((size_t) n + 1)), oEnd = take(sizeof(unsigned long long) * ((size_t) n + 1));
    const size_t oStop = take(sizeof(int) * ((size_t) n + 1)), oLen = take(sizeof(unsigned) * ((size_t) n + 1));
    const size_t oOff = take(sizeof(unsigned long long) * ((size_t) n + 2)), oScan = take(scan_workspace_bytes(n));
    PG_TRY(ctx->nextWork.reserve(o));
    PG_TRY(ctx->small.reserve(4096));
    unsigned char *bb = ctx->nextWork.as<unsigned char>();
    unsigned long long *start = (unsigned long long *) (bb + oStart), *end = (unsigned long long *) (bb + oEnd);
    int *addStop = (int *) (bb + oStop);
    unsigned *outLen = (unsigned *) (bb + oLen);
    unsigned long long *outOff = (unsigned long long *) (bb + oOff);
    PG_CUDA(cudaMemsetAsync(start, 0, oStop, s));                          // start, end
    PG_CUDA(cudaMemsetAsync(addStop, 0xFF, sizeof(int) * ((size_t) n + 1), s));   // -1
    if (nAlns && n) {
        fs_ranges_kernel<<<NUM_SMS * 8, 256, 0, s>>>(d_alns, nAlns, db->keys, n, start, end);
        fs_vote_kernel<<<NUM_SMS * 8, 256, 0, s>>>(*db, d_alns, start, end, addStop);
        ctx->launches += 2;
    }
    unsigned long long *d_tot = ctx->small.as<unsigned long long>() + 44;
    unsigned long long total = 0;
    if (n) {
        fs_len_kernel<<<NUM_SMS * 4, 256, 0, s>>>(db->lens, addStop, n, outLen);
        ctx->launches++;
        PG_TRY(exclusive_scan_u32(outLen, outOff, n, d_tot, bb + oScan, scan_workspace_bytes(n), s, &ctx->launches));
        PG_TRY(read_back(ctx, &total, d_tot, sizeof(total)));
    }
    pg_seqdb *out = new pg_seqdb();
    out->n = n; out->data_bytes = total; out->dbtype = db->dbtype;
    PG_CUDA(cudaMallocAsync(&out->data, total + 16, s));
    PG_CUDA(cudaMallocAsync(&out->offsets, sizeof(unsigned long long) * ((size_t) n + 1), s));
    PG_CUDA(cudaMallocAsync(&out->lens, sizeof(unsigned) * ((size_t) n + 1), s));
    PG_CUDA(cudaMallocAsync(&out->keys, sizeof(unsigned) * ((size_t) n + 1), s));
    if (n) {
        fs_write_kernel<<<NUM_SMS * 8, 256, 0, s>>>(*db, addStop, outLen, outOff, out->data, out->offsets, out->lens, out->keys);
        ctx->launches++;
    }
    PG_CUDA(cudaGetLastError());
    PG_TRY(seqdb_finalize(ctx, out));
    *outDb = out;
    if (d_addStop) *d_addStop = addStop;
    return 0;
}

// ================================================================================================================
// cyclecheck
// ================================================================================================================
// The reference splits the k-mers of a sequence into front / middle / back thirds, sorts each by (k-mer, position) and
// merge-joins front x back, front x middle and middle x back, where only the FIRST occurrence of a k-mer on the left
// side is joined with ALL occurrences on the right (:150-213).  Equivalent without sorting: a hash table per left side
// holding the minimum position of each k-mer; every right-side window looks its k-mer up.  All counts are integers and
// the per-diagonal histogram is order independent, so the result is bit-identical.
//
// Region of the window at position p (:124-141; `pos` is read BEFORE nextKmer advances, i.e. pos = p - 1 as unsigned,
// so the first window lands in the back third):  p == 0 -> back;  p <= third + 1 -> front;  p <= 2 * third + 1 -> middle;
// else back.
constexpr int CC_K_MAX = 31;

struct CcTables {
    unsigned long long *fKey; unsigned *fPos;     // front: k-mer -> minimum position
    unsigned long long *mKey; unsigned *mPos;     // middle
    unsigned *diagHits;                           // 2 * third + 1 counters
    unsigned char *codes;
    unsigned slots;                               // table capacity (power of two), >= 2 * (third + 2)
};

struct WarpTeam {
    __device__ static __forceinline__ void sync() { __syncwarp(); }
    __device__ static __forceinline__ int size() { return 32; }
    __device__ static __forceinline__ int rank() { return threadIdx.x & 31; }
};
struct BlockTeam {
    __device__ static __forceinline__ void sync() { __syncthreads(); }
    __device__ static __forceinline__ int size() { return blockDim.x; }
    __device__ static __forceinline__ int rank() { return threadIdx.x; }
};

__device__ __forceinline__ unsigned cc_slot(unsigned long long k, unsigned mask) { return (unsigned) (mix64(k) >> 32) & mask; }

__device__ __forceinline__ void cc_insert(unsigned long long *keys, unsigned *pos, unsigned mask, unsigned long long k, unsigned p) {
    unsigned slot = cc_slot(k, mask);
    while (true) {
        const unsigned long long old = atomicCAS(&keys[slot], ~0ULL, k);
        if (old == ~0ULL || old == k) break;
        slot = (slot + 1) & mask;
    }
    atomicMin(&pos[slot], p);
}
__device__ __forceinline__ bool cc_lookup(const unsigned long long *keys, const unsigned *pos, unsigned mask, unsigned long long k, unsigned &p) {
    unsigned slot = cc_slot(k, mask);
    while (true) {
        const unsigned long long cur = keys[slot];
        if (cur == k) { p = pos[slot]; return true; }
        if (cur == ~0ULL) return false;
        slot = (slot + 1) & mask;
    }
}

// Indexer::int2index with alphabet 4 (cyclecheck.cpp:94,126): sum code[p+i] * 4^i; codes may be 4 (N), plain arithmetic
__device__ __forceinline__ unsigned long long cc_kmer(const unsigned char *codes, unsigned p, int k) {
    unsigned long long idx = 0;
    for (int i = k - 1; i >= 0; i--) idx = idx * 4ULL + codes[p + i];
    return idx;
}

// whole check of one sequence by a team of threads; returns the split diagonal (0 = none) in every thread
template <class Team>
__device__ unsigned cc_check(const char *seq, unsigned seqLen, int k, const CcTables &t, unsigned *sBest, const unsigned char *a2n) {
    const int rank = Team::rank(), size = Team::size();
    const unsigned third = seqLen / 3;
    // Sequence::mapSequence (Sequence.cpp:476-489)
    if (rank == 0) *sBest = 0xFFFFFFFFu;
    for (unsigned i = rank; i < seqLen; i += size) t.codes[i] = a2n[(unsigned char) seq[i]];
    unsigned tsize = 64;
    while (tsize < 2 * (third + 2)) tsize <<= 1;              // <= t.slots by construction of the caller
    const unsigned mask = tsize - 1;
    for (unsigned i = rank; i < tsize; i += size) { t.fKey[i] = ~0ULL; t.fPos[i] = 0xFFFFFFFFu; t.mKey[i] = ~0ULL; t.mPos[i] = 0xFFFFFFFFu; }
    for (unsigned i = rank; i < 2 * third + 1; i += size) t.diagHits[i] = 0;
    Team::sync();
    unsigned L = seqLen;
    {   // stop at '\n' / '\0' inside the entry (never the case for a well-formed DB)
        unsigned mine = seqLen;
        for (unsigned i = rank; i < seqLen; i += size) { const char ch = seq[i]; if (ch == '\n' || ch == 0) { mine = i; break; } }
        atomicMin(sBest, mine);
        Team::sync();
        L = min(*sBest, seqLen);
        Team::sync();
        if (rank == 0) *sBest = 0xFFFFFFFFu;
    }
    if (L < (unsigned) k) { Team::sync(); return 0; }
    const unsigned nWin = L - (unsigned) k + 1;
    // left sides: first occurrence (minimum position) of every k-mer of the front and of the middle third
    for (unsigned p = 1 + rank; p < nWin; p += size) {
        if (p <= third + 1) cc_insert(t.fKey, t.fPos, mask, cc_kmer(t.codes, p, k), p);
        else if (p <= 2 * third + 1) cc_insert(t.mKey, t.mPos, mask, cc_kmer(t.codes, p, k), p);
    }
    Team::sync();
    // right sides: middle windows against the front table, back windows against both
    for (unsigned p = rank; p < nWin; p += size) {
        if (p >= 1 && p <= third + 1) continue;
        const bool back = (p == 0) || (p > 2 * third + 1);
        const unsigned long long km = cc_kmer(t.codes, p, k);
        unsigned lp;
        if (cc_lookup(t.fKey, t.fPos, mask, km, lp)) {
            const int diag = (int) (p - lp);
            if (diag >= (int) third) atomicAdd(&t.diagHits[diag - (int) third], 1u);
        }
        if (back && cc_lookup(t.mKey, t.mPos, mask, km, lp)) {
            const int diag = (int) (p - lp);
            if (diag >= (int) third) atomicAdd(&t.diagHits[diag - (int) third], 1u);
        }
    }
    Team::sync();
    // hit rate of the diagonal bands, smallest qualifying d wins (:238-262)
    unsigned mineBest = 0xFFFFFFFFu;
    for (unsigned d = rank; d < 2 * third; d += size) {
        const unsigned hd = t.diagHits[d];
        if (hd == 0) continue;
        const unsigned diag = d + third;
        const unsigned diaglen = seqLen - diag;
        const unsigned gapwindow = (unsigned) ((double) diaglen * 0.01);
        const int lowerS = (int) (d - gapwindow);
        const unsigned lower = lowerS > 0 ? (unsigned) lowerS : 0u;
        const unsigned upper = min(d + gapwindow, 2 * third);
        unsigned band = 0;
        for (unsigned i = lower; i <= upper; i++) { const unsigned h = t.diagHits[i]; if (h <= hd) band += h; }
        const unsigned long long denom = (unsigned long long) diaglen - (unsigned long long) k + 1ULL;    // size_t arithmetic, may wrap
        const float rate = (float) band / (float) denom;
        if ((double) rate > 0.2) { mineBest = d; break; }
    }
    if (mineBest != 0xFFFFFFFFu) atomicMin(sBest, mineBest);
    Team::sync();
    const unsigned best = *sBest;
    Team::sync();
    return best == 0xFFFFFFFFu ? 0u : best + third;
}

// SLOTS-slot tables in shared memory: sequences with third + 2 <= SLOTS / 2
template <int SLOTS>
struct CcSmall {
    static constexpr int MAX_THIRD = SLOTS / 2 - 2;
    static constexpr int MAX_LEN = 3 * MAX_THIRD + 2;
    static constexpr int DIAG = 2 * MAX_THIRD + 1;
    static constexpr int CODES = (MAX_LEN + 15) & ~15;
    static constexpr size_t PER_WARP = (size_t) SLOTS * 8 * 2 + (size_t) SLOTS * 4 * 2 + (((size_t) DIAG * 4 + 15) & ~(size_t) 15) + CODES + 16;
};

__constant__ unsigned char c_cc_a2n[256];

template <int SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) cc_warp_kernel(const pg_seqdb db, unsigned lenLo, unsigned lenHi, unsigned maxSeqLen, int k,
                                                             unsigned *__restrict__ split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef CcSmall<SLOTS> S;
    const int w = threadIdx.x >> 5;
    unsigned char *base = smem_raw + (size_t) w * S::PER_WARP;
    CcTables t;
    t.fKey = reinterpret_cast<unsigned long long *>(base);
    t.mKey = t.fKey + SLOTS;
    t.fPos = reinterpret_cast<unsigned *>(t.mKey + SLOTS);
    t.mPos = t.fPos + SLOTS;
    t.diagHits = t.mPos + SLOTS;
    t.codes = base + (size_t) SLOTS * 24 + (((size_t) S::DIAG * 4 + 15) & ~(size_t) 15);
    unsigned *sBest = reinterpret_cast<unsigned *>(t.codes + S::CODES);
    t.slots = SLOTS;
    const unsigned n = (unsigned) db.n;
    const unsigned warpsTotal = gridDim.x * WARPS;
    for (unsigned i = blockIdx.x * WARPS + w; i < n; i += warpsTotal) {
        const unsigned seqLen = db.lens[i] - 2;
        if (seqLen < lenLo || seqLen > lenHi) continue;            // another instance's class
        unsigned res = 0;
        if (seqLen < maxSeqLen && seqLen >= (unsigned) k)          // :107-112 (too long: skipped, not reported)
            res = cc_check<WarpTeam>(db.data + db.offsets[i], seqLen, k, t, sBest, c_cc_a2n);
        if ((threadIdx.x & 31) == 0) split[i] = res;
    }
}

__global__ void cc_list_kernel(const unsigned *__restrict__ lens, unsigned n, unsigned lenLo, unsigned *__restrict__ list, unsigned *__restrict__ count) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (lens[i] - 2 >= lenLo) list[atomicAdd(count, 1u)] = i;
}

// long sequences: one CTA per sequence, tables in global scratch
__global__ void __launch_bounds__(256) cc_block_kernel(const pg_seqdb db, const unsigned *__restrict__ list, const unsigned *__restrict__ listCount,
                                                       unsigned *__restrict__ ticket, unsigned maxSeqLen, int k, unsigned slots, unsigned maxLen,
                                                       unsigned char *__restrict__ scratch, size_t scratchPerBlock, unsigned *__restrict__ split) {
    __shared__ unsigned sBest, sItem;
    unsigned char *base = scratch + (size_t) blockIdx.x * scratchPerBlock;
    CcTables t;
    t.fKey = reinterpret_cast<unsigned long long *>(base);
    t.mKey = t.fKey + slots;
    t.fPos = reinterpret_cast<unsigned *>(t.mKey + slots);
    t.mPos = t.fPos + slots;
    t.diagHits = t.mPos + slots;
    t.codes = reinterpret_cast<unsigned char *>(t.diagHits + (2 * (size_t) (maxLen / 3) + 4));
    t.slots = slots;
    const unsigned nList = *listCount;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sItem = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned li = sItem;
        if (li >= nList) break;
        const unsigned i = list[li];
        const unsigned seqLen = db.lens[i] - 2;
        unsigned res = 0;
        if (seqLen < maxSeqLen && seqLen >= (unsigned) k)
            res = cc_check<BlockTeam>(db.data + db.offsets[i], seqLen, k, t, &sBest, c_cc_a2n);
        if (threadIdx.x == 0) split[i] = res;
    }
}

int cc_run(Context *ctx, const pg_seqdb *db, int maxSeqLen, int k, unsigned **d_split) {
    cudaStream_t s = ctx->stream;
    PG_CHECK(db->dbtype == PG_DBTYPE_NUCLEOTIDES, "cyclecheck: only nucleotide sequence DBs are supported (cyclecheck.cpp:49-54)");
    PG_CHECK(k >= 2 && k <= CC_K_MAX, "cyclecheck: k-mer size must be in [2, 31]");
    PG_CHECK(maxSeqLen > 0, "cyclecheck: --max-seq-len must be positive");
    const unsigned n = (unsigned) db->n;
    PG_CUDA(cudaMemcpyToSymbolAsync(c_cc_a2n, PG_NT_AA2NUM, 256, 0, cudaMemcpyHostToDevice, s));
    typedef CcSmall<256> S0;
    typedef CcSmall<1024> S1;
    constexpr int W0 = 8, W1 = 4;
    const unsigned bigLo = (unsigned) S1::MAX_LEN + 1;
    const bool haveBig = db->max_seq_len >= bigLo;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t) 15; return r; };
    const size_t oSplit = take(sizeof(unsigned) * ((size_t) n + 1));
    const size_t oList = take(sizeof(unsigned) * ((size_t) n + 1));
    const size_t oCnt = take(64);
    PG_TRY(ctx->nextWork.reserve(o));
    unsigned char *bb = ctx->nextWork.as<unsigned char>();
    unsigned *split = (unsigned *) (bb + oSplit), *list = (unsigned *) (bb + oList), *cnt = (unsigned *) (bb + oCnt);
    PG_CUDA(cudaMemsetAsync(split, 0, sizeof(unsigned) * ((size_t) n + 1), s));
    PG_CUDA(cudaMemsetAsync(cnt, 0, 64, s));
    if (n) {
        static std::atomic<unsigned long long> attrDev{0};
        if (first_use_on_device(attrDev)) {
            PG_CUDA(cudaFuncSetAttribute(cc_warp_kernel<256, W0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (S0::PER_WARP * W0)));
            PG_CUDA(cudaFuncSetAttribute(cc_warp_kernel<1024, W1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (S1::PER_WARP * W1)));
        }
        const unsigned blocks0 = std::min<unsigned>((n + W0 - 1) / W0, NUM_SMS * 16);
        cc_warp_kernel<256, W0><<<blocks0, W0 * 32, S0::PER_WARP * W0, s>>>(*db, 0u, (unsigned) S0::MAX_LEN, (unsigned) maxSeqLen, k, split);
        ctx->launches++;
        if (db->max_seq_len > (unsigned) S0::MAX_LEN) {
            const unsigned blocks1 = std::min<unsigned>((n + W1 - 1) / W1, NUM_SMS * 8);
            cc_warp_kernel<1024, W1><<<blocks1, W1 * 32, S1::PER_WARP * W1, s>>>(*db, (unsigned) S0::MAX_LEN + 1, (unsigned) S1::MAX_LEN, (unsigned) maxSeqLen, k, split);
            ctx->launches++;
        }
        if (haveBig) {
            cc_list_kernel<<<NUM_SMS * 4, 256, 0, s>>>(db->lens, n, bigLo, list, cnt);
            ctx->launches++;
            unsigned nBig = 0;
            PG_TRY(read_back(ctx, &nBig, cnt, sizeof(nBig)));
            // only sequences below --max-seq-len are checked: the tables never need more than that
            const unsigned maxLen = std::min<unsigned>(db->max_seq_len, (unsigned) maxSeqLen);
            unsigned slots = 64;
            while (slots < 2 * (maxLen / 3 + 2)) slots <<= 1;
            const size_t perBlock = (((size_t) slots * 24 + (2 * (size_t) (maxLen / 3) + 4) * 4 + (size_t) maxLen + 64) + 255) & ~(size_t) 255;
            const unsigned blocks = std::max(1u, std::min<unsigned>(nBig, NUM_SMS * 2));
            PG_TRY(ctx->scratch.reserve(perBlock * blocks));
            cc_block_kernel<<<blocks, 256, 0, s>>>(*db, list, cnt, cnt + 4, (unsigned) maxSeqLen, k, slots, maxLen, ctx->scratch.as<unsigned char>(), perBlock, split);
            ctx->launches++;
        }
    }
    PG_CUDA(cudaGetLastError());
    *d_split = split;
    return 0;
}

}  // namespace pg
