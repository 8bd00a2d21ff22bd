// pg_scan.cu -- see pg_scan.cuh.
#include "pg_scan.cuh"

namespace pg {
namespace {
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *sWarp, unsigned long long &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long nb = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += nb; }
    if (lane == 31) sWarp[w] = inc;
    __syncthreads();
    unsigned long long woff = 0, tot = 0;
    for (int ww = 0; ww < SCAN_THREADS / 32; ww++) { if (ww < w) woff += sWarp[ww]; tot += sWarp[ww]; }
    total = tot;
    __syncthreads();
    return woff + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(const unsigned *__restrict__ in, unsigned long long *__restrict__ out,
                                                                  unsigned long long n, unsigned long long *__restrict__ blockSums) {
    __shared__ unsigned long long sWarp[SCAN_THREADS / 32];
    const unsigned long long base = (unsigned long long) blockIdx.x * SCAN_TILE + (unsigned long long) threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0u; mine += v[i]; }
    unsigned long long total;
    unsigned long long off = block_exclusive_scan(mine, sWarp, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = off; off += v[i]; }
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(unsigned long long *__restrict__ sums, unsigned long long nBlocks,
                                                                 unsigned long long *__restrict__ total) {
    __shared__ unsigned long long sWarp[SCAN_THREADS / 32];
    __shared__ unsigned long long sCarry;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    for (unsigned long long b = 0; b < nBlocks; b += SCAN_THREADS) {
        const unsigned long long i = b + threadIdx.x;
        const unsigned long long c = (i < nBlocks) ? sums[i] : 0ULL;
        unsigned long long tot;
        const unsigned long long off = block_exclusive_scan(c, sWarp, tot);
        const unsigned long long carry = sCarry;
        if (i < nBlocks) sums[i] = carry + off;
        __syncthreads();
        if (threadIdx.x == 0) sCarry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = sCarry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(unsigned long long *__restrict__ out, unsigned long long n,
                                                                const unsigned long long *__restrict__ blockSums) {
    const unsigned long long add = blockSums[blockIdx.x];
    const unsigned long long base = (unsigned long long) blockIdx.x * SCAN_TILE + (unsigned long long) threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) out[base + i] += add;
}
}  // namespace

size_t scan_workspace_bytes(uint64_t n) { return sizeof(unsigned long long) * ((n + SCAN_TILE - 1) / SCAN_TILE + 8); }

int exclusive_scan_u32(const unsigned *in, unsigned long long *out, uint64_t n, unsigned long long *d_total,
                       void *ws, size_t wsBytes, cudaStream_t stream, uint64_t *launches) {
    if (n == 0) { PG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream)); return 0; }
    PG_CHECK(wsBytes >= scan_workspace_bytes(n), "exclusive_scan_u32: workspace too small");
    const unsigned long long blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long *sums = (unsigned long long *) ws;
    scan_local_kernel<<<(unsigned) blocks, SCAN_THREADS, 0, stream>>>(in, out, n, sums);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(sums, blocks, d_total);
    scan_add_kernel<<<(unsigned) blocks, SCAN_THREADS, 0, stream>>>(out, n, sums);
    if (launches) *launches += 3;
    PG_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace pg
