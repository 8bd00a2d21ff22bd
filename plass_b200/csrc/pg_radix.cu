// pg_radix.cu -- see pg_radix.cuh.
#include "pg_radix.cuh"

#include <algorithm>
#include <cstdlib>

namespace pg {

namespace {

// EXT = false: the two digit kinds of the single-GPU path (plain bits, bits of mix64); EXT = true adds the two kinds of
// the multi-GPU path (owner interval, rebased key).  Separate instances: the extra cases cost the hot kernels ~12 %.
template <bool EXT>
__device__ __forceinline__ unsigned digit_of_t(const Rec &r, const DigitPass p) {
    unsigned long long w = p.word ? r.w1 : r.w0;
    if (EXT) {
        if (p.hashed == 2) {
            const unsigned key = (unsigned) (r.w0 >> 32);
            unsigned d = 0;
            for (unsigned i = 1; i < p.auxN; i++) d += (key >= __ldg(p.aux + i)) ? 1u : 0u;
            return d;
        }
        if (p.hashed == 3) { w = (w >> 32) - p.hashMask; return (unsigned) (w >> p.shift) & p.mask; }
    }
    if (p.hashed) w = mix64(w & p.hashMask);
    return (unsigned) (w >> p.shift) & p.mask;
}
__device__ __forceinline__ unsigned digit_of(const Rec &r, const DigitPass p) { return digit_of_t<true>(r, p); }

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- histograms of all passes in one read -------------------------------------------------------
template <bool EXT>
__global__ void __launch_bounds__(512) radix_hist_kernel(const Rec *__restrict__ in, unsigned long long n, RadixPlan plan,
                                                         unsigned long long *__restrict__ ghist, int stride) {
    extern __shared__ unsigned sh[];          // npasses x stride counters (stride = 256, or 512 / 1024 with wide digits)
    const int np = plan.npasses;
    for (int i = threadIdx.x; i < np * stride; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned long long gstride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(in) + i);
        Rec r;
        r.w0 = ((unsigned long long) raw.y << 32) | raw.x;
        r.w1 = ((unsigned long long) raw.w << 32) | raw.z;
#pragma unroll 4
        for (int p = 0; p < np; p++) atomicAdd(&sh[p * stride + digit_of_t<EXT>(r, plan.pass[p])], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < np * stride; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], (unsigned long long) sh[i]);
}

// histograms counted by the producer of the records (256 bins per pass): digit d of a pass with a narrower mask collects the
// bins congruent to d
__global__ void radix_fold_hist_kernel(const unsigned long long *__restrict__ pre, unsigned long long *__restrict__ ghist, RadixPlan plan) {
    const int p = blockIdx.x, d = threadIdx.x;
    const unsigned mask = plan.pass[p].mask;
    unsigned long long c = 0;
    if ((unsigned) d <= mask) for (unsigned b = (unsigned) d; b < 256u; b += mask + 1u) c += pre[p * 256 + b];
    ghist[p * 256 + d] = c;
}

// exclusive scan of each pass's bins -> bases[p][portion 0][stride]
__global__ void radix_scan_kernel(const unsigned long long *__restrict__ ghist, unsigned long long *__restrict__ bases,
                                  int portionsPlus1, int stride) {
    __shared__ unsigned long long s[1024];
    const int p = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < stride; i += blockDim.x) s[i] = ghist[p * stride + i];
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < stride; i++) { unsigned long long c = s[i]; s[i] = run; run += c; }
    }
    __syncthreads();
    for (int i = t; i < stride; i += blockDim.x) bases[((size_t) p * portionsPlus1) * stride + i] = s[i];
}

// ---- one pass over one portion ------------------------------------------------------------------
constexpr unsigned FLAG_AGG = 1u, FLAG_INC = 2u;

// PEER: the digit runs do not go to one output array but to dstBase[digit] (a byte address per digit, possibly in ANOTHER GPU's
// memory, mapped through CUDA IPC over NVLink) + the record's rank within the digit.  This is the partition pass and the
// all-to-all of the multi-GPU exchanges in one kernel: the 16-byte stores of a run leave the SM as NVLink writes to the owner's
// receive buffer instead of HBM writes followed by a separate send.  gbase0 = the digit bases of portion 0 (rank 0 of a digit).
template <int ITEMS, int MINBLOCKS, bool EXT, bool BOUNDS, bool PEER = false>
__global__ void __launch_bounds__(RADIX_THREADS, MINBLOCKS) radix_scatter_kernel(
    const Rec *__restrict__ in, Rec *__restrict__ out, unsigned long long portionStart, unsigned long long portionEnd,
    DigitPass dp, const unsigned long long *__restrict__ gbase, unsigned long long *__restrict__ gbaseNext,
    unsigned *status, unsigned *ticket, unsigned numTiles, RadixBounds bo,
    const unsigned long long *__restrict__ dstBase = nullptr, const unsigned long long *__restrict__ gbase0 = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec *tileRecs = reinterpret_cast<Rec *>(smem_raw);
    __shared__ unsigned short warpCnt[RADIX_THREADS / 32][256];
    __shared__ unsigned digitStart[256];
    __shared__ long long goff[256];
    __shared__ unsigned warpTotals[RADIX_THREADS / 32];
    __shared__ unsigned sTile;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    for (int i = tid; i < (RADIX_THREADS / 32) * 256; i += RADIX_THREADS) (&warpCnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = sTile;
    const unsigned long long tileBase = portionStart + (unsigned long long) tile * (RADIX_THREADS * ITEMS);
    const unsigned count = (unsigned) min((unsigned long long) (RADIX_THREADS * ITEMS), portionEnd - tileBase);

    // load: warp w owns items [w*512, (w+1)*512), round r covers 32 consecutive records
    Rec rec[ITEMS];
    unsigned short rank[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
        if (idx < count) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(in + tileBase) + idx);
            rec[r].w0 = ((unsigned long long) raw.y << 32) | raw.x;
            rec[r].w1 = ((unsigned long long) raw.w << 32) | raw.z;
        } else {
            rec[r].w0 = 0; rec[r].w1 = 0;
        }
    }
    // rank: warp-private counters + match_any
    const unsigned ltMask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
        const bool valid = idx < count;
        const unsigned d = valid ? digit_of_t<EXT>(rec[r], dp) : 256u;
        const unsigned m = __match_any_sync(0xFFFFFFFFu, d);
        unsigned c = 0;
        if (valid) c = warpCnt[w][d];
        rank[r] = (unsigned short) (c + __popc(m & ltMask));
        __syncwarp();
        if (valid && (m & ltMask) == 0) warpCnt[w][d] = (unsigned short) (c + __popc(m));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, total count
    unsigned cnt;
    {
        unsigned run = 0;
#pragma unroll
        for (int ww = 0; ww < RADIX_THREADS / 32; ww++) {
            const unsigned c = warpCnt[ww][tid];
            warpCnt[ww][tid] = (unsigned short) run;
            run += c;
        }
        cnt = run;
    }
    // publish the aggregate early so successors can look back while we keep working
    if (tile == 0) st_volatile_u32(&status[tid], (cnt << 2) | FLAG_INC);
    else st_volatile_u32(&status[(size_t) tile * 256 + tid], (cnt << 2) | FLAG_AGG);
    // block exclusive scan of cnt over the 256 digits
    {
        unsigned v = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (lane >= o) v += nb;
        }
        if (lane == 31) warpTotals[w] = v;
        __syncthreads();
        unsigned woff = 0;
#pragma unroll
        for (int ww = 0; ww < RADIX_THREADS / 32; ww++) woff += (ww < w) ? warpTotals[ww] : 0u;
        digitStart[tid] = woff + v - cnt;
    }
    __syncthreads();
    // reorder through shared memory
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
        if (idx < count) {
            // (keeping the digit from the ranking step in registers / next to the record was measured: no gain, the pass is
            // bound by the latency of its dependent phases, not by recomputing the hash)
            const unsigned d = digit_of_t<EXT>(rec[r], dp);
            const unsigned pos = digitStart[d] + warpCnt[w][d] + rank[r];
            tileRecs[pos] = rec[r];
        }
    }
    // decoupled look-back: exclusive prefix of this digit over all earlier tiles of the portion
    {
        unsigned long long prev = 0;
        if (tile > 0) {
            // walk back over the predecessors' status words, 8 independent loads at a time (memory-level
            // parallelism: the walk costs ~1/8 of the L2 round trips of a one-by-one walk)
            long long ll = (long long) tile - 1;
            bool done = false;
            while (!done) {
                unsigned v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const long long idx = ll - u;
                    v[u] = (idx >= 0) ? ld_volatile_u32(&status[(size_t) idx * 256 + tid]) : FLAG_INC;   // before tile 0: prefix 0
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (done) break;
                    const unsigned f = v[u] & 3u;
                    if (f == 0u) break;                       // not published yet: reload from here
                    prev += v[u] >> 2;
                    ll--;
                    if (f == FLAG_INC) done = true;
                }
            }
            st_volatile_u32(&status[(size_t) tile * 256 + tid], ((unsigned) (prev + cnt) << 2) | FLAG_INC);
        }
        const unsigned long long base = gbase[tid];
        if (PEER) goff[tid] = (long long) dstBase[tid] + ((long long) (base - gbase0[tid] + prev) - (long long) digitStart[tid]) * (long long) sizeof(Rec);
        else goff[tid] = (long long) (base + prev) - (long long) digitStart[tid];
        if (tile == numTiles - 1) gbaseNext[tid] = base + prev + cnt;
    }
    __syncthreads();
    if (PEER) {
        for (unsigned j = tid; j < count; j += RADIX_THREADS) {
            const Rec r = tileRecs[j];
            const unsigned d = digit_of_t<EXT>(r, dp);
            uint4 raw;
            raw.x = (unsigned) r.w0; raw.y = (unsigned) (r.w0 >> 32); raw.z = (unsigned) r.w1; raw.w = (unsigned) (r.w1 >> 32);
            *reinterpret_cast<uint4 *>(goff[d] + (long long) j * (long long) sizeof(Rec)) = raw;
        }
    } else if (!BOUNDS) {
        for (unsigned j = tid; j < count; j += RADIX_THREADS) {
            const Rec r = tileRecs[j];
            const unsigned d = digit_of_t<EXT>(r, dp);
            uint4 raw;
            raw.x = (unsigned) r.w0; raw.y = (unsigned) (r.w0 >> 32); raw.z = (unsigned) r.w1; raw.w = (unsigned) (r.w1 >> 32);
            reinterpret_cast<uint4 *>(out)[goff[d] + (long long) j] = raw;
        }
    } else {
        // last pass of a hashed-bucket partition: the store loop hashes every record anyway (its digit is a slice of
        // mix64), so the bucket id -- all sorted bits of the same hash -- is free, and so are the bucket boundaries:
        // neighbours in the reordered tile are neighbours in the output as long as they share the digit, and the bucket id
        // contains the digit.
        const int lane = tid & 31;
        unsigned long long localMin = ~0ULL;
        if (bo.kind == 1) {
            // segments by the high half of w0 (sort #2: the representative): first / one-past-last output index of every value and
            // the smallest low half (target) per value, the latter combined per warp before it touches memory
            for (unsigned j0 = 0; j0 < count; j0 += RADIX_THREADS) {
                const unsigned j = j0 + tid;
                const bool valid = j < count;
                unsigned key = 0xFFFFFFFFu, low = 0xFFFFFFFFu;
                Rec r; r.w0 = 0; r.w1 = 0;
                if (valid) { r = tileRecs[j]; key = (unsigned) (r.w0 >> 32); low = (unsigned) r.w0; }
                unsigned kp = __shfl_up_sync(0xFFFFFFFFu, key, 1), kn = __shfl_down_sync(0xFFFFFFFFu, key, 1);
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
                const unsigned mt = __reduce_min_sync(peers, low);
                if (valid) {
                    if (lane == 0) kp = j > 0 ? (unsigned) (tileRecs[j - 1].w0 >> 32) : 0xFFFFFFFFu;
                    if (lane == 31) kn = j + 1 < count ? (unsigned) (tileRecs[j + 1].w0 >> 32) : 0xFFFFFFFFu;
                    if (j + 1 >= count) kn = 0xFFFFFFFFu;
                    const unsigned d = digit_of_t<EXT>(r, dp);
                    // neighbours with the same key share the digit; a key change inside the tile is a segment boundary only if the
                    // neighbour is also adjacent in the output, which an equal digit guarantees -- and a differing digit implies a
                    // differing key anyway
                    const unsigned long long g = (unsigned long long) (goff[d] + (long long) j);
                    uint4 raw;
                    raw.x = (unsigned) r.w0; raw.y = (unsigned) (r.w0 >> 32); raw.z = (unsigned) r.w1; raw.w = (unsigned) (r.w1 >> 32);
                    reinterpret_cast<uint4 *>(out)[g] = raw;
                    if (kp != key) atomicMin(&bo.start[key], g);
                    if (kn != key) atomicMax(&bo.end[key], g + 1ULL);
                    if (lane == __ffs(peers) - 1) atomicMin(&bo.minLow[key], mt);
                }
            }
            return;
        }
        for (unsigned j0 = 0; j0 < count; j0 += RADIX_THREADS) {
            const unsigned j = j0 + tid;
            const bool valid = j < count;
            unsigned long long hb = ~0ULL;
            Rec r; r.w0 = 0; r.w1 = 0;
            if (valid) {
                r = tileRecs[j];
                const unsigned long long k = r.w0 & bo.hashMask;
                localMin = min(localMin, k);
                hb = mix64(k) & (unsigned long long) bo.bucketMask;
            }
            unsigned long long bp = __shfl_up_sync(0xFFFFFFFFu, hb, 1), bn = __shfl_down_sync(0xFFFFFFFFu, hb, 1);
            if (valid) {
                if (lane == 0) bp = j > 0 ? (mix64(tileRecs[j - 1].w0 & bo.hashMask) & (unsigned long long) bo.bucketMask) : ~0ULL;
                if (lane == 31) bn = j + 1 < count ? (mix64(tileRecs[j + 1].w0 & bo.hashMask) & (unsigned long long) bo.bucketMask) : ~0ULL;
                if (j + 1 >= count) bn = ~0ULL;
                const unsigned d = (unsigned) (hb >> dp.shift) & dp.mask;
                const unsigned long long g = (unsigned long long) (goff[d] + (long long) j);
                uint4 raw;
                raw.x = (unsigned) r.w0; raw.y = (unsigned) (r.w0 >> 32); raw.z = (unsigned) r.w1; raw.w = (unsigned) (r.w1 >> 32);
                reinterpret_cast<uint4 *>(out)[g] = raw;
                if (bp != hb) atomicMin(&bo.start[(unsigned) hb], g);
                if (bn != hb) atomicMax(&bo.end[(unsigned) hb], g + 1ULL);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) localMin = min(localMin, __shfl_xor_sync(0xFFFFFFFFu, localMin, o));
        if (lane == 0 && localMin != ~0ULL) atomicMin(bo.minKey, localMin);
    }
}

// Wide-digit instance (BITS = 9 / 10: 512 / 1024 bins, DPT = 2 / 4 digits per thread).  Same structure as the 256-bin
// kernel above; the per-digit start inside the tile is folded into the warp-private counters (a tile holds < 65536
// records) so that shared memory stays at 16 B x tile + 2 B x 8 x BINS + 8 B x BINS (72 KB for 1024 bins: 3 CTAs / SM).
template <int ITEMS, int MINBLOCKS, int BITS>
__global__ void __launch_bounds__(RADIX_THREADS, MINBLOCKS) radix_scatter_wide_kernel(
    const Rec *__restrict__ in, Rec *__restrict__ out, unsigned long long portionStart, unsigned long long portionEnd,
    DigitPass dp, const unsigned long long *__restrict__ gbase, unsigned long long *__restrict__ gbaseNext,
    unsigned *status, unsigned *ticket, unsigned numTiles) {
    constexpr int BINS = 1 << BITS, DPT = BINS / RADIX_THREADS, WARPS = RADIX_THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec *tileRecs = reinterpret_cast<Rec *>(smem_raw);
    unsigned short *tileDig = reinterpret_cast<unsigned short *>(smem_raw + (size_t) RADIX_THREADS * ITEMS * sizeof(Rec));
    __shared__ unsigned short warpCnt[WARPS][BINS];
    __shared__ long long goff[BINS];
    __shared__ unsigned warpTotals[WARPS];
    __shared__ unsigned sTile;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * BINS; i += RADIX_THREADS) (&warpCnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = sTile;
    const unsigned long long tileBase = portionStart + (unsigned long long) tile * (RADIX_THREADS * ITEMS);
    const unsigned count = (unsigned) min((unsigned long long) (RADIX_THREADS * ITEMS), portionEnd - tileBase);

    Rec rec[ITEMS];
    unsigned short rank[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
        if (idx < count) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(in + tileBase) + idx);
            rec[r].w0 = ((unsigned long long) raw.y << 32) | raw.x;
            rec[r].w1 = ((unsigned long long) raw.w << 32) | raw.z;
        } else {
            rec[r].w0 = 0; rec[r].w1 = 0;
        }
    }
    const unsigned ltMask = (1u << lane) - 1u;
    unsigned dpack[(ITEMS + 1) / 2];
#pragma unroll
    for (int r = 0; r < (ITEMS + 1) / 2; r++) dpack[r] = 0;
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
        const bool valid = idx < count;
        const unsigned d = valid ? digit_of(rec[r], dp) : (unsigned) BINS;
        if (valid) dpack[r >> 1] |= d << ((r & 1) * 16);
        const unsigned m = __match_any_sync(0xFFFFFFFFu, d);
        unsigned c = 0;
        if (valid) c = warpCnt[w][d];
        rank[r] = (unsigned short) (c + __popc(m & ltMask));
        __syncwarp();
        if (valid && (m & ltMask) == 0) warpCnt[w][d] = (unsigned short) (c + __popc(m));
        __syncwarp();
    }
    __syncthreads();
    // thread tid owns the digits j * 256 + tid
    unsigned cnt[DPT], dStart[DPT];
#pragma unroll
    for (int j = 0; j < DPT; j++) {
        const int d = j * RADIX_THREADS + tid;
        unsigned run = 0;
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) {
            const unsigned c = warpCnt[ww][d];
            warpCnt[ww][d] = (unsigned short) run;
            run += c;
        }
        cnt[j] = run;
        if (tile == 0) st_volatile_u32(&status[d], (run << 2) | FLAG_INC);
        else st_volatile_u32(&status[(size_t) tile * BINS + d], (run << 2) | FLAG_AGG);
    }
    // exclusive scan over all BINS digits in digit order
    {
        unsigned carry = 0;
#pragma unroll
        for (int j = 0; j < DPT; j++) {
            unsigned v = cnt[j];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= o) v += nb;
            }
            if (lane == 31) warpTotals[w] = v;
            __syncthreads();
            unsigned woff = 0, total = 0;
#pragma unroll
            for (int ww = 0; ww < WARPS; ww++) { const unsigned t = warpTotals[ww]; woff += (ww < w) ? t : 0u; total += t; }
            dStart[j] = carry + woff + v - cnt[j];
            carry += total;
            __syncthreads();
        }
    }
    // fold the digit start into the warp-private prefixes: position in the tile = warpCnt[w][d] + rank
#pragma unroll
    for (int j = 0; j < DPT; j++) {
        const int d = j * RADIX_THREADS + tid;
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) warpCnt[ww][d] = (unsigned short) (warpCnt[ww][d] + dStart[j]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
        if (idx < count) {
            const unsigned d = (dpack[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu;
            const unsigned pos = (unsigned) warpCnt[w][d] + rank[r];
            tileRecs[pos] = rec[r];
            tileDig[pos] = (unsigned short) d;
        }
    }
    // decoupled look-back per owned digit
#pragma unroll
    for (int j = 0; j < DPT; j++) {
        const int d = j * RADIX_THREADS + tid;
        unsigned long long prev = 0;
        if (tile > 0) {
            long long ll = (long long) tile - 1;
            bool done = false;
            while (!done) {
                unsigned v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const long long idx = ll - u;
                    v[u] = (idx >= 0) ? ld_volatile_u32(&status[(size_t) idx * BINS + d]) : FLAG_INC;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (done) break;
                    const unsigned f = v[u] & 3u;
                    if (f == 0u) break;
                    prev += v[u] >> 2;
                    ll--;
                    if (f == FLAG_INC) done = true;
                }
            }
            st_volatile_u32(&status[(size_t) tile * BINS + d], ((unsigned) (prev + cnt[j]) << 2) | FLAG_INC);
        }
        const unsigned long long base = gbase[d];
        goff[d] = (long long) (base + prev) - (long long) dStart[j];
        if (tile == numTiles - 1) gbaseNext[d] = base + prev + cnt[j];
    }
    __syncthreads();
    for (unsigned j = tid; j < count; j += RADIX_THREADS) {
        const Rec r = tileRecs[j];
        const unsigned d = tileDig[j];
        uint4 raw;
        raw.x = (unsigned) r.w0; raw.y = (unsigned) (r.w0 >> 32); raw.z = (unsigned) r.w1; raw.w = (unsigned) (r.w1 >> 32);
        reinterpret_cast<uint4 *>(out)[goff[d] + (long long) j] = raw;
    }
}


// ---- persistent bulk-copy (TMA) pass ------------------------------------------------------------------------------
// One CTA keeps taking tiles (ticket order, so the decoupled look-back never waits on a tile that has not started).  The
// records of tile i+1 are fetched by ONE bulk asynchronous copy (cp.async.bulk.shared.global, completion counted on an
// mbarrier) while tile i is ranked, so the load latency that the register-tile kernel exposes once per tile is off the
// critical path; after the in-place reorder each digit run is one contiguous piece of shared memory AND of the output,
// and the thread that owns the digit hands it to the copy engine as one bulk shared -> global copy.  STAGES buffers:
// with 3, the stores of tile i-1 may still be draining while tile i is ranked and tile i+1 is loading.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int THREADS, int MINB, int ITEMS, int STAGES, bool BOUNDS>
__global__ void __launch_bounds__(THREADS, MINB) radix_scatter_tma_kernel(
    const Rec *__restrict__ in, Rec *__restrict__ out, unsigned long long portionStart, unsigned long long portionEnd,
    DigitPass dp, const unsigned long long *__restrict__ gbase, unsigned long long *__restrict__ gbaseNext,
    unsigned *status, unsigned *ticket, unsigned numTiles, RadixBounds bo, unsigned staggerCycles) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];          // STAGES x TILE records
    __shared__ unsigned short warpCnt[WARPS][256];
    __shared__ unsigned digitStart[256];
    __shared__ long long goff[BOUNDS ? 256 : 1];
    __shared__ unsigned warpTotals[8];
    __shared__ unsigned sTile[STAGES];
    __shared__ __align__(8) unsigned long long bar[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool digitThread = tid < 256;                                  // thread d < 256 owns digit d
    const unsigned ltMask = (1u << lane) - 1u;
    auto bufOf = [&](int st) { return reinterpret_cast<Rec *>(smem_raw) + (size_t) st * TILE; };
    auto tileCount = [&](unsigned tile) {
        const unsigned long long tb = portionStart + (unsigned long long) tile * TILE;
        return (unsigned) min((unsigned long long) TILE, portionEnd - tb);
    };
    // one thread: take the next ticket and start its load into stage st
    auto fetch = [&](int st) {
        const unsigned t = atomicAdd(ticket, 1u);
        sTile[st] = t;
        if (t < numTiles) {
            const unsigned bytes = tileCount(t) * (unsigned) sizeof(Rec);
            fence_async_smem();                                       // generic-proxy accesses to the stage before the async write
            mbar_expect_tx(&bar[st], bytes);
            const unsigned char *src = reinterpret_cast<const unsigned char *>(in + portionStart + (unsigned long long) t * TILE);
            unsigned char *dst = reinterpret_cast<unsigned char *>(bufOf(st));
            // pieces of at most 16 KB: several copies in flight per tile
            for (unsigned o = 0; o < bytes; o += 16384u) bulk_g2s(dst + o, src + o, min(16384u, bytes - o), &bar[st]);
        }
    };
    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // All CTAs of a persistent grid start together; left alone they would run in lockstep, every tile's look-back would find
    // the ~300 tiles in flight all in the "aggregate only" state and walk back over a whole wave of status words.  Spreading
    // the first tickets over one tile period keeps the tiles in flight at evenly spread phases (as a grid of short-lived
    // CTAs is by itself), so a look-back finds an inclusive prefix within a few tiles.
    if (tid == 0) {
        const long long wait = (long long) ((unsigned long long) staggerCycles * blockIdx.x / gridDim.x);
        const long long t0 = clock64();
        while (clock64() - t0 < wait) { }
        fetch(0);
    }
    unsigned long long localMin = ~0ULL;
    __syncthreads();

    for (unsigned it = 0;; it++) {
        const int cur = (int) (it % STAGES);
        const unsigned tile = sTile[cur];
        if (tile >= numTiles) break;
        const unsigned count = tileCount(tile);
        Rec *buf = bufOf(cur);
        // the stage the next tile goes to was last used by tile it+1-STAGES, whose bulk stores have been waited for
        if (tid == 0) fetch((int) ((it + 1) % STAGES));
        for (int i = tid; i < WARPS * 128; i += THREADS) reinterpret_cast<unsigned *>(&warpCnt[0][0])[i] = 0;
        while (!mbar_try_wait(&bar[cur], (it / STAGES) & 1u)) { }
        __syncthreads();

        // records of this thread: warp w owns [w * ITEMS * 32, ...), round r covers 32 consecutive records
        Rec rec[ITEMS];
        unsigned short rank[ITEMS];
        unsigned dpack[(ITEMS + 3) / 4];
#pragma unroll
        for (int r = 0; r < (ITEMS + 3) / 4; r++) dpack[r] = 0;
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
            if (idx < count) {
                const uint4 raw = *reinterpret_cast<const uint4 *>(buf + idx);
                rec[r].w0 = ((unsigned long long) raw.y << 32) | raw.x;
                rec[r].w1 = ((unsigned long long) raw.w << 32) | raw.z;
            } else {
                rec[r].w0 = 0; rec[r].w1 = 0;
            }
        }
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
            const bool valid = idx < count;
            const unsigned d = valid ? digit_of(rec[r], dp) : 256u;
            if (valid) dpack[r >> 2] |= d << ((r & 3) * 8);
            const unsigned m = __match_any_sync(0xFFFFFFFFu, d);
            unsigned c = 0;
            if (valid) c = warpCnt[w][d];
            rank[r] = (unsigned short) (c + __popc(m & ltMask));
            __syncwarp();
            if (valid && (m & ltMask) == 0) warpCnt[w][d] = (unsigned short) (c + __popc(m));
            __syncwarp();
        }
        __syncthreads();
        // per digit: exclusive prefix over warps, total count
        unsigned cnt = 0, myStart = 0;
        if (digitThread) {
            unsigned run = 0;
#pragma unroll
            for (int ww = 0; ww < WARPS; ww++) {
                const unsigned c = warpCnt[ww][tid];
                warpCnt[ww][tid] = (unsigned short) run;
                run += c;
            }
            cnt = run;
            if (tile == 0) st_volatile_u32(&status[tid], (cnt << 2) | FLAG_INC);
            else st_volatile_u32(&status[(size_t) tile * 256 + tid], (cnt << 2) | FLAG_AGG);
        }
        {
            unsigned v = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= o) v += nb;
            }
            if (digitThread && lane == 31) warpTotals[w] = v;
            __syncthreads();
            if (digitThread) {
                unsigned woff = 0;
#pragma unroll
                for (int ww = 0; ww < 8; ww++) woff += (ww < w) ? warpTotals[ww] : 0u;
                myStart = woff + v - cnt;
                digitStart[tid] = myStart;
            }
        }
        __syncthreads();
        // in-place reorder: every thread holds its records in registers since before the two barriers above
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const unsigned idx = w * (ITEMS * 32) + r * 32 + lane;
            if (idx < count) {
                const unsigned d = (dpack[r >> 2] >> ((r & 3) * 8)) & 0xFFu;
                const unsigned pos = digitStart[d] + warpCnt[w][d] + rank[r];
                uint4 raw;
                raw.x = (unsigned) rec[r].w0; raw.y = (unsigned) (rec[r].w0 >> 32); raw.z = (unsigned) rec[r].w1; raw.w = (unsigned) (rec[r].w1 >> 32);
                *reinterpret_cast<uint4 *>(buf + pos) = raw;
            }
        }
        fence_async_smem();                       // the reordered tile must be visible to the copy engine
        // decoupled look-back: exclusive prefix of this digit over all earlier tiles of the portion
        unsigned long long myGlobal = 0;
        if (digitThread) {
            unsigned long long prev = 0;
            if (tile > 0) {
                long long ll = (long long) tile - 1;
                bool done = false;
                while (!done) {
                    unsigned v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const long long idx = ll - u;
                        v[u] = (idx >= 0) ? ld_volatile_u32(&status[(size_t) idx * 256 + tid]) : FLAG_INC;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (done) break;
                        const unsigned f = v[u] & 3u;
                        if (f == 0u) break;
                        prev += v[u] >> 2;
                        ll--;
                        if (f == FLAG_INC) done = true;
                    }
                }
                st_volatile_u32(&status[(size_t) tile * 256 + tid], ((unsigned) (prev + cnt) << 2) | FLAG_INC);
            }
            const unsigned long long base = gbase[tid];
            myGlobal = base + prev;
            if (BOUNDS) goff[tid] = (long long) myGlobal - (long long) myStart;
            if (tile == numTiles - 1) gbaseNext[tid] = base + prev + cnt;
        }
        __syncthreads();
        // one bulk copy per digit run
        if (digitThread && cnt) bulk_s2g(out + myGlobal, buf + myStart, cnt * (unsigned) sizeof(Rec));
        bulk_commit();
        if (BOUNDS) {
            // bucket boundaries + smallest key from the reordered tile (neighbours in the tile are neighbours in the output
            // as long as they share the digit, and the bucket id contains the digit)
            for (unsigned j0 = 0; j0 < count; j0 += THREADS) {
                const unsigned j = j0 + tid;
                const bool valid = j < count;
                unsigned long long hb = ~0ULL;
                if (valid) {
                    const unsigned long long k = buf[j].w0 & bo.hashMask;
                    localMin = min(localMin, k);
                    hb = mix64(k) & (unsigned long long) bo.bucketMask;
                }
                unsigned long long bp = __shfl_up_sync(0xFFFFFFFFu, hb, 1), bn = __shfl_down_sync(0xFFFFFFFFu, hb, 1);
                if (valid) {
                    if (lane == 0) bp = j > 0 ? (mix64(buf[j - 1].w0 & bo.hashMask) & (unsigned long long) bo.bucketMask) : ~0ULL;
                    if (lane == 31) bn = j + 1 < count ? (mix64(buf[j + 1].w0 & bo.hashMask) & (unsigned long long) bo.bucketMask) : ~0ULL;
                    if (j + 1 >= count) bn = ~0ULL;
                    const unsigned d = (unsigned) (hb >> dp.shift) & dp.mask;
                    const unsigned long long g = (unsigned long long) (goff[d] + (long long) j);
                    if (bp != hb) atomicMin(&bo.start[(unsigned) hb], g);
                    if (bn != hb) atomicMax(&bo.end[(unsigned) hb], g + 1ULL);
                }
            }
        }
        // the stage that the NEXT iteration's fetch targets must have been read out by the copy engine
        bulk_wait_read<STAGES - 2>();
        __syncthreads();
    }
    bulk_wait_read<0>();
    if (BOUNDS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) localMin = min(localMin, __shfl_xor_sync(0xFFFFFFFFu, localMin, o));
        if (lane == 0 && localMin != ~0ULL) atomicMin(bo.minKey, localMin);
    }
}

static int g_items = 12;   // records per thread of the scatter kernel (8 / 12 / 16), see radix_set_items
static inline unsigned long long tile_records() { return (unsigned long long) RADIX_THREADS * g_items; }
constexpr unsigned long long PORTION_RECORDS = ((1ull << 30) / 12288 - 1) * 12288;   // look-back prefix < 2^30; multiple of every tile size

inline unsigned long long num_portions(uint64_t n) { return n == 0 ? 1 : (n + PORTION_RECORDS - 1) / PORTION_RECORDS; }
inline unsigned long long max_tiles(uint64_t n) {
    const unsigned long long per = n < PORTION_RECORDS ? n : PORTION_RECORDS;
    return (per + 2048 - 1) / 2048 + 1;                 // sized for the smallest tile
}

}  // namespace

void radix_set_items(int items) { g_items = (items == 8 || items == 16) ? items : 12; }

// 0: register-tile kernel; 1 / 2 / 3: bulk-copy kernel variants (see tma_tile)
static int g_mode = -1;
int radix_get_mode() {
    if (g_mode < 0) {
        g_mode = 0;      // measured: the register-tile kernel is still the fastest on the real record streams (profiles/r2_summary_a.md)
        if (const char *e = getenv("PLASS_B200_RADIX_MODE")) { const int m = atoi(e); if (m >= 0 && m <= 3) g_mode = m; }
    }
    return g_mode;
}
void radix_set_mode(int mode) { g_mode = (mode >= 0 && mode <= 3) ? mode : 0; }

bool radix_emits_bounds(const RadixPlan &plan) {
    if (plan.npasses == 0 || (radix_get_mode() == 0 && g_items != 12)) return false;
    for (int p = 0; p < plan.npasses; p++) if (plan.pass[p].mask > 255u || plan.pass[p].hashed != 1) return false;
    return true;
}

bool radix_emits_segments(const RadixPlan &plan) {
    if (plan.npasses == 0 || radix_get_mode() != 0 || g_items != 12) return false;
    for (int p = 0; p < plan.npasses; p++) if (plan.pass[p].mask > 255u || plan.pass[p].hashed == 1 || plan.pass[p].hashed == 2 || plan.pass[p].word != 0) return false;
    return true;
}

template <int THREADS, int MINB, int ITEMS, int STAGES>
static int launch_tma(const Rec *src, Rec *dst, unsigned long long ps, unsigned long long pe, const DigitPass &dp, const unsigned long long *gb,
                      unsigned long long *gbNext, unsigned *status, unsigned *ticket, const RadixBounds *bounds, cudaStream_t stream) {
    constexpr int TILE = THREADS * ITEMS;
    const int smem = STAGES * TILE * (int) sizeof(Rec);
    static std::atomic<unsigned long long> attrDev[2];
    const unsigned tiles = (unsigned) ((pe - ps + TILE - 1) / TILE);
    const unsigned grid = std::min<unsigned>(tiles, (unsigned) NUM_SMS * (unsigned) MINB);
    // one tile period: 2 x TILE x 16 B per CTA at this SM's share of ~5 TB/s (read + write), in SM cycles at ~1.9 GHz
    static int stagger = -1;
    if (stagger < 0) { stagger = 11000; if (const char *e = getenv("PLASS_B200_RADIX_STAGGER")) stagger = std::max(0, atoi(e)); }
    const unsigned staggerCycles = tiles > grid ? (unsigned) stagger : 0u;
    if (bounds) {
        if (first_use_on_device(attrDev[1])) PG_CUDA(cudaFuncSetAttribute(radix_scatter_tma_kernel<THREADS, MINB, ITEMS, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        radix_scatter_tma_kernel<THREADS, MINB, ITEMS, STAGES, true><<<grid, THREADS, smem, stream>>>(src, dst, ps, pe, dp, gb, gbNext, status, ticket, tiles, *bounds, staggerCycles);
    } else {
        if (first_use_on_device(attrDev[0])) PG_CUDA(cudaFuncSetAttribute(radix_scatter_tma_kernel<THREADS, MINB, ITEMS, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        radix_scatter_tma_kernel<THREADS, MINB, ITEMS, STAGES, false><<<grid, THREADS, smem, stream>>>(src, dst, ps, pe, dp, gb, gbNext, status, ticket, tiles, RadixBounds(), staggerCycles);
    }
    return 0;
}

// tile size of the bulk-copy variants: mode 1 = 512 threads x 6 records, 2 stages, 2 CTAs / SM (32 warps);
// mode 2 = 256 threads x 8 records, 2 stages, 3 CTAs / SM (24 warps); mode 3 = 512 threads x 4 records, 3 stages, 2 CTAs / SM
static unsigned tma_tile(int mode) { return mode == 1 ? 3072u : 2048u; }

void plan_add_bits(RadixPlan &plan, int word, int lo, int hi) {
    for (int b = lo; b < hi; b += 8) {
        const int bits = (hi - b) < 8 ? (hi - b) : 8;
        DigitPass &p = plan.pass[plan.npasses++];
        p.word = word; p.shift = b; p.mask = (1u << bits) - 1u; p.hashed = 0; p.hashMask = 0; p.aux = nullptr; p.auxN = 0;
    }
}

void plan_add_interval(RadixPlan &plan, const unsigned *deviceBounds, unsigned n) {
    DigitPass &p = plan.pass[plan.npasses++];
    p.word = 0; p.shift = 0; p.mask = 255u; p.hashed = 2; p.hashMask = 0; p.aux = deviceBounds; p.auxN = n;
}

void plan_add_rebased_high_bits(RadixPlan &plan, unsigned long long sub, int nbits) {
    for (int b = 0; b < nbits; b += 8) {
        const int bits = (nbits - b) < 8 ? (nbits - b) : 8;
        DigitPass &p = plan.pass[plan.npasses++];
        p.word = 0; p.shift = b; p.mask = (1u << bits) - 1u; p.hashed = 3; p.hashMask = sub; p.aux = nullptr; p.auxN = 0;
    }
}

void plan_add_hash_bits(RadixPlan &plan, unsigned long long hashMask, int lo, int hi) {
    for (int b = lo; b < hi; b += 8) {
        const int bits = (hi - b) < 8 ? (hi - b) : 8;
        DigitPass &p = plan.pass[plan.npasses++];
        p.word = 0; p.shift = b; p.mask = (1u << bits) - 1u; p.hashed = 1; p.hashMask = hashMask; p.aux = nullptr; p.auxN = 0;
    }
}

static int plan_stride(const RadixPlan &plan) {
    unsigned mx = 255;
    for (int p = 0; p < plan.npasses; p++) mx = plan.pass[p].mask > mx ? plan.pass[p].mask : mx;
    return mx > 511 ? 1024 : (mx > 255 ? 512 : 256);
}

void plan_add_bits_w(RadixPlan &plan, int word, int lo, int hi, int digitBits) {
    const int total = hi - lo;
    if (total <= 0) return;
    const int nd = (total + digitBits - 1) / digitBits;
    int b = lo;
    for (int i = 0; i < nd; i++) {
        const int bits = (total - (b - lo) + (nd - i) - 1) / (nd - i);      // spread the bits evenly over the digits
        DigitPass &p = plan.pass[plan.npasses++];
        p.word = word; p.shift = b; p.mask = (1u << bits) - 1u; p.hashed = 0; p.hashMask = 0; p.aux = nullptr; p.auxN = 0;
        b += bits;
    }
}

void plan_add_hash_bits_w(RadixPlan &plan, unsigned long long hashMask, int lo, int hi, int digitBits) {
    const int first = plan.npasses;
    plan_add_bits_w(plan, 0, lo, hi, digitBits);
    for (int p = first; p < plan.npasses; p++) { plan.pass[p].hashed = 1; plan.pass[p].hashMask = hashMask; }
}

size_t radix_workspace_bytes(uint64_t n, int maxDigitBits) {
    const size_t stride = maxDigitBits > 9 ? 1024 : (maxDigitBits > 8 ? 512 : 256);
    const size_t hist = sizeof(unsigned long long) * RADIX_MAX_PASSES * stride;
    const size_t bases = sizeof(unsigned long long) * RADIX_MAX_PASSES * (num_portions(n) + 1) * stride;
    const size_t status = sizeof(unsigned) * (max_tiles(n) * stride + 64);
    return hist + bases + status + 1024;
}

// ---- one pass whose output goes to per-digit destinations (fused partition + exchange, see the PEER kernel) ---------------
static void peer_ws_layout(void *workspace, uint64_t n, unsigned long long **ghist, unsigned long long **bases, unsigned **status, size_t *statusWords) {
    unsigned char *ws = (unsigned char *) workspace;
    *ghist = (unsigned long long *) ws;
    *bases = *ghist + (size_t) RADIX_MAX_PASSES * 256;
    *status = (unsigned *) (*bases + (size_t) RADIX_MAX_PASSES * (num_portions(n) + 1) * 256);
    *statusWords = max_tiles(n) * (size_t) 256 + 64;
}

int radix_pass_histogram(const Rec *a, uint64_t n, const DigitPass &pass, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                         unsigned long long *hostHist, uint64_t *launches) {
    PG_CHECK(workspace_bytes >= radix_workspace_bytes(n, 8) && pass.mask <= 255u, "radix_pass_histogram: workspace too small / digit wider than 8 bits");
    unsigned long long *ghist, *bases; unsigned *status; size_t statusWords;
    peer_ws_layout(workspace, n, &ghist, &bases, &status, &statusWords);
    PG_CUDA(cudaMemsetAsync(ghist, 0, sizeof(unsigned long long) * 256, stream));
    RadixPlan plan; plan.npasses = 1; plan.pass[0] = pass;
    if (n) {
        int histBlocks = (int) std::min<unsigned long long>((n + 512ull * 16 - 1) / (512ull * 16), (unsigned long long) NUM_SMS * 4);
        if (pass.hashed >= 2) radix_hist_kernel<true><<<histBlocks, 512, 256 * sizeof(unsigned), stream>>>(a, n, plan, ghist, 256);
        else radix_hist_kernel<false><<<histBlocks, 512, 256 * sizeof(unsigned), stream>>>(a, n, plan, ghist, 256);
    }
    radix_scan_kernel<<<1, 256, 0, stream>>>(ghist, bases, (int) (num_portions(n) + 1), 256);
    if (launches) *launches += 2;
    PG_CUDA(cudaMemcpyAsync(hostHist, ghist, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost, stream));
    PG_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

int radix_scatter_peer(const Rec *a, uint64_t n, const DigitPass &pass, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                       const unsigned long long *d_dstBase, uint64_t *launches) {
    if (n == 0) return 0;
    PG_CHECK(workspace_bytes >= radix_workspace_bytes(n, 8) && pass.mask <= 255u, "radix_scatter_peer: workspace too small / digit wider than 8 bits");
    unsigned long long *ghist, *bases; unsigned *status; size_t statusWords;
    peer_ws_layout(workspace, n, &ghist, &bases, &status, &statusWords);
    static std::atomic<unsigned long long> attrDev{0};
    const int dynSmem = 3072 * (int) sizeof(Rec);
    if (first_use_on_device(attrDev)) {
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<12, 3, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dynSmem));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<12, 3, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dynSmem));
    }
    const unsigned long long portions = num_portions(n);
    for (unsigned long long q = 0; q < portions; q++) {
        const unsigned long long ps = q * PORTION_RECORDS;
        const unsigned long long pe = (ps + PORTION_RECORDS < n) ? ps + PORTION_RECORDS : n;
        const unsigned tiles = (unsigned) ((pe - ps + 3072 - 1) / 3072);
        PG_CUDA(cudaMemsetAsync(status, 0, sizeof(unsigned) * ((size_t) tiles * 256 + 64), stream));
        unsigned *ticket = status + statusWords - 32;
        PG_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), stream));
        unsigned long long *gb = bases + q * 256;
        if (pass.hashed >= 2)
            radix_scatter_kernel<12, 3, true, false, true><<<tiles, RADIX_THREADS, dynSmem, stream>>>(a, nullptr, ps, pe, pass, gb, gb + 256, status, ticket, tiles, RadixBounds(), d_dstBase, bases);
        else
            radix_scatter_kernel<12, 3, false, false, true><<<tiles, RADIX_THREADS, dynSmem, stream>>>(a, nullptr, ps, pe, pass, gb, gb + 256, status, ticket, tiles, RadixBounds(), d_dstBase, bases);
        if (launches) *launches += 1;
    }
    PG_CUDA(cudaGetLastError());
    return 0;
}

int radix_sort(Rec *a, Rec *b, uint64_t n, const RadixPlan &plan, void *workspace, size_t workspace_bytes,
               cudaStream_t stream, Rec **sorted, uint64_t *launches, cudaEvent_t evScatterBegin, cudaEvent_t evScatterEnd, const RadixBounds *bounds,
               const unsigned long long *preHist) {
    *sorted = a;
    if (n == 0 || plan.npasses == 0) return 0;
    PG_CHECK(plan.npasses <= RADIX_MAX_PASSES, "radix_sort: too many passes");
    const int stride = plan_stride(plan);
    PG_CHECK(workspace_bytes >= radix_workspace_bytes(n, stride == 1024 ? 10 : (stride == 512 ? 9 : 8)), "radix_sort: workspace too small");
    PG_CHECK((size_t) plan.npasses * stride * sizeof(unsigned) <= 48 * 1024, "radix_sort: too many wide passes for one histogram launch");
    static std::atomic<unsigned long long> attrDev{0};
    const int dynSmem = (int) (tile_records() * sizeof(Rec));
    const int dynSmemWide = (int) (tile_records() * (sizeof(Rec) + 2));    // wide digits: two bytes
    if (first_use_on_device(attrDev)) {
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<16, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int) sizeof(Rec)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<12, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * (int) sizeof(Rec)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<12, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * (int) sizeof(Rec)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<12, 3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * (int) sizeof(Rec)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<12, 3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * (int) sizeof(Rec)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<8, 4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * (int) sizeof(Rec)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_wide_kernel<16, 2, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int) (sizeof(Rec) + 2)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_wide_kernel<12, 3, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * (int) (sizeof(Rec) + 2)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_wide_kernel<16, 2, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int) (sizeof(Rec) + 2)));
        PG_CUDA(cudaFuncSetAttribute(radix_scatter_wide_kernel<12, 3, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * (int) (sizeof(Rec) + 2)));
    }
    const unsigned long long portions = num_portions(n);
    unsigned char *ws = (unsigned char *) workspace;
    unsigned long long *ghist = (unsigned long long *) ws;
    unsigned long long *bases = ghist + (size_t) RADIX_MAX_PASSES * stride;
    unsigned *status = (unsigned *) (bases + (size_t) RADIX_MAX_PASSES * (portions + 1) * stride);
    const size_t statusWords = max_tiles(n) * (size_t) stride + 64;

    PG_CUDA(cudaMemsetAsync(ghist, 0, sizeof(unsigned long long) * RADIX_MAX_PASSES * stride, stream));
    int histBlocks = (int) ((n + 512ull * 16 - 1) / (512ull * 16));
    if (histBlocks > NUM_SMS * 4) histBlocks = NUM_SMS * 4;
    if (histBlocks < 1) histBlocks = 1;
    bool ext = false;
    for (int p = 0; p < plan.npasses; p++) ext = ext || plan.pass[p].hashed >= 2;
    bool usePre = preHist != nullptr && stride == 256 && plan.npasses <= 3;
    for (int p = 0; p < plan.npasses && usePre; p++) {
        const DigitPass &dp = plan.pass[p];
        usePre = dp.hashed == 1 && dp.word == 0 && dp.shift == 8 * p && dp.mask <= 255u && ((dp.mask + 1u) & dp.mask) == 0u;
    }
    if (usePre) radix_fold_hist_kernel<<<plan.npasses, 256, 0, stream>>>(preHist, ghist, plan);
    else if (ext) radix_hist_kernel<true><<<histBlocks, 512, (size_t) plan.npasses * stride * sizeof(unsigned), stream>>>(a, n, plan, ghist, stride);
    else radix_hist_kernel<false><<<histBlocks, 512, (size_t) plan.npasses * stride * sizeof(unsigned), stream>>>(a, n, plan, ghist, stride);
    radix_scan_kernel<<<plan.npasses, 256, 0, stream>>>(ghist, bases, (int) (portions + 1), stride);
    if (launches) *launches += 2;

    Rec *src = a, *dst = b;
    if (evScatterBegin) cudaEventRecord(evScatterBegin, stream);
    for (int p = 0; p < plan.npasses; p++) {
        for (unsigned long long q = 0; q < portions; q++) {
            const unsigned long long ps = q * PORTION_RECORDS;
            const unsigned long long pe = (ps + PORTION_RECORDS < n) ? ps + PORTION_RECORDS : n;
            unsigned tiles = (unsigned) ((pe - ps + tile_records() - 1) / tile_records());
            const int bins = plan.pass[p].mask > 511 ? 1024 : (plan.pass[p].mask > 255 ? 512 : 256);
            const int mode = radix_get_mode();
            if (bins == 256 && mode != 0) tiles = (unsigned) ((pe - ps + tma_tile(mode) - 1) / tma_tile(mode));
            PG_CUDA(cudaMemsetAsync(status, 0, sizeof(unsigned) * ((size_t) tiles * bins + 64), stream));
            unsigned *ticket = status + statusWords - 32;
            PG_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), stream));
            unsigned long long *gb = bases + ((size_t) p * (portions + 1) + q) * stride;
            if (bins == 256 && mode != 0) {
                const RadixBounds *bp = (bounds && bounds->kind == 0 && p == plan.npasses - 1 && radix_emits_bounds(plan)) ? bounds : nullptr;
                if (mode == 2) PG_TRY((launch_tma<256, 3, 8, 2>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, bp, stream)));
                else if (mode == 3) PG_TRY((launch_tma<512, 2, 4, 3>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, bp, stream)));
                else PG_TRY((launch_tma<512, 2, 6, 2>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, bp, stream)));
            } else if (bins == 512) {
                if (g_items == 16) radix_scatter_wide_kernel<16, 2, 9><<<tiles, RADIX_THREADS, dynSmemWide, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles);
                else { PG_CHECK(g_items == 12, "radix_sort: wide digits need 12 or 16 records per thread"); radix_scatter_wide_kernel<12, 3, 9><<<tiles, RADIX_THREADS, dynSmemWide, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles); }
            } else if (bins == 1024) {
                if (g_items == 16) radix_scatter_wide_kernel<16, 2, 10><<<tiles, RADIX_THREADS, dynSmemWide, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles);
                else { PG_CHECK(g_items == 12, "radix_sort: wide digits need 12 or 16 records per thread"); radix_scatter_wide_kernel<12, 3, 10><<<tiles, RADIX_THREADS, dynSmemWide, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles); }
            } else if (plan.pass[p].hashed >= 2) {
                PG_CHECK(g_items == 12, "radix_sort: interval / rebased digits need 12 records per thread");
                if (bounds && bounds->kind == 1 && p == plan.npasses - 1 && radix_emits_segments(plan))
                    radix_scatter_kernel<12, 3, true, true><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, *bounds);
                else
                    radix_scatter_kernel<12, 3, true, false><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, RadixBounds());
            } else if (bounds && bounds->kind == 1 && p == plan.npasses - 1 && radix_emits_segments(plan)) {
                radix_scatter_kernel<12, 3, false, true><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, *bounds);
            } else if (bounds && bounds->kind == 0 && p == plan.npasses - 1 && radix_emits_bounds(plan) && g_items == 12) {
                radix_scatter_kernel<12, 3, false, true><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, *bounds);
            } else if (g_items == 16) radix_scatter_kernel<16, 2, false, false><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, RadixBounds());
            else if (g_items == 12) radix_scatter_kernel<12, 3, false, false><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, RadixBounds());
            else radix_scatter_kernel<8, 4, false, false><<<tiles, RADIX_THREADS, dynSmem, stream>>>(src, dst, ps, pe, plan.pass[p], gb, gb + stride, status, ticket, tiles, RadixBounds());
            if (launches) *launches += 1;
        }
        Rec *t = src; src = dst; dst = t;
    }
    if (evScatterEnd) cudaEventRecord(evScatterEnd, stream);
    PG_CUDA(cudaGetLastError());
    *sorted = src;
    return 0;
}

}  // namespace pg
