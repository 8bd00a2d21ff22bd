// pg_shard.cu -- multi-GPU data plane in C++: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces the reference's split / merge machinery for several workers (lib/mmseqs/src/linclust/kmermatcher.cpp:632-694:
// splits round-robin over MPI ranks, results merged from files by rank 0; rescorediagonal.cpp:399-421) with an in-memory
// exchange:
//
//   rank r extracts the k-mers of ITS SLICE of the sequences (the DB itself is replicated in every HBM);
//   exchange #1  k-mer records  -> the rank owning the k-mer (top bits of mix64(k-mer): equal k-mers meet, which is all
//                                   sort #1 + assignGroup need, so the representatives are those of the unsplit run)
//   sort #1 + assignGroup on the owner -> (rep, target, diagonal) pair records; a histogram of the pairs over the
//                                   representative key space is all-reduced and cut into `world` ranges of equal work
//   exchange #2  pair records   -> the rank owning the representative (the "single all-to-all of candidate pairs")
//   sort #2 + best diagonal, rescorediagonal and the extension for the owned queries; the new sequences of all ranks are
//   all-gathered so that the next iteration starts from a replicated DB again (data/assemble.sh:153 INPUT=assembly_$STEP).
//
// Exchanges are grouped ncclSend / ncclRecv on the context's stream straight between the record buffers of the stages
// (no staging copies, no host round trip except the W x W count matrix that sizes the receive buffers).
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a process that already holds NCCL (PyTorch) shares that copy, the
// CLI gets the system library, and a build without multi-GPU use never needs it.
#include "pg_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace pg {

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    // a copy the process already holds (PyTorch's) first; otherwise the system library, kept local
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    PG_CHECK(h != nullptr, std::string("multi-GPU: cannot load libnccl.so.2 (") + (dlerror() ? dlerror() : "?") + ")");
    auto sym = [&](const char *name) { return dlsym(h, name); };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId)) sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank)) sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy)) sym("ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString)) sym("ncclGetErrorString");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart)) sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd)) sym("ncclGroupEnd");
    g_nccl.Send = (decltype(g_nccl.Send)) sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv)) sym("ncclRecv");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce)) sym("ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather)) sym("ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast)) sym("ncclBroadcast");
    PG_CHECK(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.GetErrorString && g_nccl.GroupStart && g_nccl.GroupEnd &&
                 g_nccl.Send && g_nccl.Recv && g_nccl.AllReduce && g_nccl.AllGather && g_nccl.Broadcast,
             "multi-GPU: libnccl.so.2 lacks a required symbol");
    g_nccl.handle = h;
    return 0;
}

#define PG_NCCL(call)                                                                                                   \
    do {                                                                                                                \
        ncclResult_t r_ = (call);                                                                                       \
        if (r_ != ncclSuccess) {                                                                                        \
            pg::set_error(std::string(#call) + " failed: " + g_nccl.GetErrorString(r_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return 1;                                                                                                   \
        }                                                                                                               \
    } while (0)

inline ncclComm_t comm_of(Context *ctx) { return (ncclComm_t) ctx->comm; }

// device -> host read-back of up to a few KB through the mapped-pinned staging words (1 KB per round)
int read_back_big(Context *ctx, void *host, const void *dev, size_t bytes) {
    for (size_t o = 0; o < bytes; o += 1024) PG_TRY(read_back(ctx, (char *) host + o, (const char *) dev + o, std::min<size_t>(1024, bytes - o)));
    return 0;
}

// All-to-all-v of 16-byte records: send[r] records to rank r from sendBuf (packed in rank order), receive into recvBuf
// (packed in source-rank order).  The count matrix is all-gathered first so that every rank can size its receive side.
int exchange_records(Context *ctx, const Rec *sendBuf, const uint64_t *sendCounts, DevBuf &recvBuf, uint64_t *nRecv, int which) {
    cudaStream_t s = ctx->stream;
    const int W = ctx->world, me = ctx->rank;
    PG_TRY(ctx->commWs.reserve(sizeof(unsigned long long) * (size_t) (W * W + W + 8) + sizeof(unsigned long long) * 8192));
    unsigned long long *d_mine = ctx->commWs.as<unsigned long long>();          // [W]
    unsigned long long *d_all = d_mine + W;                                      // [W x W], row = source rank
    std::vector<unsigned long long> mine(sendCounts, sendCounts + W), all((size_t) W * W);
    PG_CUDA(cudaMemcpyAsync(d_mine, mine.data(), sizeof(unsigned long long) * W, cudaMemcpyHostToDevice, s));
    PG_NCCL(g_nccl.AllGather(d_mine, d_all, (size_t) W, ncclUint64, comm_of(ctx), s));
    PG_TRY(read_back_big(ctx, all.data(), d_all, sizeof(unsigned long long) * (size_t) W * W));
    uint64_t total = 0;
    for (int r = 0; r < W; r++) total += all[(size_t) r * W + me];
    PG_TRY(recvBuf.reserve(sizeof(Rec) * (total + 1)));
    Rec *recv = recvBuf.as<Rec>();
    cudaEventRecord(ctx->evXchg[2 * which], s);
    uint64_t so = 0, ro = 0, sentAway = 0;
    PG_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < W; r++) {
        const uint64_t sc = sendCounts[r], rc = all[(size_t) r * W + me];
        if (r == me) {
            if (sc) PG_CUDA(cudaMemcpyAsync(recv + ro, sendBuf + so, sizeof(Rec) * sc, cudaMemcpyDeviceToDevice, s));
        } else {
            if (sc) PG_NCCL(g_nccl.Send(sendBuf + so, sc * sizeof(Rec), ncclUint8, r, comm_of(ctx), s));
            if (rc) PG_NCCL(g_nccl.Recv(recv + ro, rc * sizeof(Rec), ncclUint8, r, comm_of(ctx), s));
            sentAway += sc;
        }
        so += sc; ro += rc;
    }
    PG_NCCL(g_nccl.GroupEnd());
    cudaEventRecord(ctx->evXchg[2 * which + 1], s);
    ctx->lastExchangeBytes[which] = sentAway * sizeof(Rec);
    *nRecv = total;
    return 0;
}

__global__ void min_kmer_kernel(const Rec *__restrict__ in, unsigned long long n, unsigned long long hashMask, unsigned long long *__restrict__ out) {
    unsigned long long m = ~0ULL;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) m = min(m, in[i].w0 & hashMask);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out, m);
}

// sharded.py:balanced_bounds in C++: cuts the representative key space [0, max_key] into `world` contiguous ranges of
// (nearly) equal work.  hist[b] = pair records (summed over all ranks) whose representative falls into bin
// b = rep * BINS / (max_key + 1); the work of a bin = its pair records + perKeyWeight x its keys (every owned sequence is
// also a query with a self alignment and an output entry).  Cuts fall on bin edges; every rank computes the same bounds
// from the same summed histogram.
void balanced_bounds(const unsigned long long *hist, int bins, unsigned maxKey, int world, double perKeyWeight, unsigned *bounds) {
    const unsigned long long span = (unsigned long long) maxKey + 1ull;
    std::vector<unsigned long long> edges((size_t) bins + 1);
    for (int b = 0; b <= bins; b++) edges[(size_t) b] = ((unsigned long long) b * span + (unsigned long long) bins - 1ull) / (unsigned long long) bins;
    std::vector<double> cum((size_t) bins + 1, 0.0);
    for (int b = 0; b < bins; b++) cum[(size_t) b + 1] = cum[(size_t) b] + (double) hist[b] + perKeyWeight * (double) (edges[(size_t) b + 1] - edges[(size_t) b]);
    bounds[0] = 0;
    for (int r = 1; r < world; r++) {
        const double target = cum[(size_t) bins] * (double) r / (double) world;
        int b = (int) (std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        b = std::min(std::max(b, 0), bins);
        bounds[r] = (unsigned) std::max<unsigned long long>(edges[(size_t) b], bounds[r - 1]);
    }
    bounds[world] = 0xFFFFFFFFu;
}

__global__ void rebase_offsets_kernel(unsigned long long *__restrict__ offsets, unsigned long long lo, unsigned long long hi, unsigned long long base) {
    for (unsigned long long i = lo + (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (unsigned long long) gridDim.x * blockDim.x) offsets[i] += base;
}

}  // namespace

// ---- fused partition + exchange over peer memory -------------------------------------------------------------------
// Every rank owns two receive buffers (xr[0]: k-mer records of exchange #1, xr[1]: pair records of exchange #2) that all other
// ranks map through CUDA IPC.  A sender's partition pass (radix_scatter_peer) writes each digit run straight into the owner's
// buffer over NVLink; a stream-ordered barrier (a one-word all-reduce) tells the owner that every sender's pass has finished.
// Growth is decided from the all-gathered count matrix, which every rank holds identically, so all ranks re-allocate and
// re-exchange handles in the same iteration.
static int stream_barrier(Context *ctx) {
    unsigned long long *d = ctx->commWs.as<unsigned long long>() + 7000;
    PG_NCCL(g_nccl.AllReduce(d, d, 1, ncclUint64, ncclSum, comm_of(ctx), ctx->stream));
    return 0;
}

static void close_peer_mappings(Context *ctx, int which) {
    for (int r = 0; r < ctx->world; r++) {
        if (r != ctx->rank && ctx->peer[which][r]) cudaIpcCloseMemHandle(ctx->peer[which][r]);
        ctx->peer[which][r] = nullptr;
    }
}

// returns 0 and *ok = true when every rank's buffer `which` holds at least needBytes and is mapped everywhere
static int ensure_peer_buffers(Context *ctx, int which, size_t needBytes, bool *ok) {
    *ok = false;
    cudaStream_t s = ctx->stream;
    const int W = ctx->world, me = ctx->rank;
    if (ctx->xr[which].cap >= needBytes && ctx->peer[which][me] == ctx->xr[which].p) { *ok = true; return 0; }
    // (re)allocate: nobody may still be writing into the old buffer -- the previous exchange ended with a barrier, and the
    // owner's own consumers run on this stream
    PG_CUDA(cudaStreamSynchronize(s));
    PG_TRY(stream_barrier(ctx));
    PG_CUDA(cudaStreamSynchronize(s));
    close_peer_mappings(ctx, which);
    ctx->xr[which].release();
    PG_TRY(ctx->xr[which].reserve(needBytes + needBytes / 4));
    cudaIpcMemHandle_t mine;
    PG_CUDA(cudaIpcGetMemHandle(&mine, ctx->xr[which].p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    unsigned char *d_mine = ctx->commWs.as<unsigned char>() + 8 * 8000;          // 64 B, then W x 64 B (past the histogram area)
    unsigned char *d_all = d_mine + 64;
    std::vector<unsigned char> all((size_t) 64 * W);
    PG_CUDA(cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, s));
    PG_NCCL(g_nccl.AllGather(d_mine, d_all, 64, ncclUint8, comm_of(ctx), s));
    PG_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t) 64 * W, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaStreamSynchronize(s));
    unsigned long long good = 1;
    for (int r = 0; r < W; r++) {
        if (r == me) { ctx->peer[which][r] = ctx->xr[which].p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, all.data() + (size_t) 64 * r, 64);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = 0; p = nullptr; }
        ctx->peer[which][r] = p;
    }
    // all ranks or none
    unsigned long long *d_flag = ctx->commWs.as<unsigned long long>() + 7001;
    PG_CUDA(cudaMemcpyAsync(d_flag, &good, sizeof(good), cudaMemcpyHostToDevice, s));
    PG_NCCL(g_nccl.AllReduce(d_flag, d_flag, 1, ncclUint64, ncclMin, comm_of(ctx), s));
    PG_TRY(read_back(ctx, &good, d_flag, sizeof(good)));
    if (!good) { close_peer_mappings(ctx, which); ctx->xr[which].release(); return 0; }
    *ok = true;
    return 0;
}

// One fused partition + exchange: the n records at `src` are partitioned by `pass` (digit -> owner through ownerOfDigit) and
// written into the owners' receive buffers `which`.  On return (stream-ordered) this rank's buffer holds *nRecv records.
// *done = false if peer memory is not available (the caller falls back to NCCL send / recv).
static int exchange_fused(Context *ctx, const Rec *src, uint64_t n, const DigitPass &pass, int nDigits, const int *ownerOfDigit, int which,
                          uint64_t *nRecv, bool *done) {
    *done = false;
    cudaStream_t s = ctx->stream;
    const int W = ctx->world, me = ctx->rank;
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(n)));
    PG_TRY(ctx->commWs.reserve(sizeof(unsigned long long) * 16384));
    unsigned long long h[256];
    PG_TRY(radix_pass_histogram(src, n, pass, ctx->radixWs.p, ctx->radixWs.cap, s, h, &ctx->launches));
    std::vector<unsigned long long> mine((size_t) W, 0), all((size_t) W * W), gbase(257, 0);
    for (int b = 0; b < nDigits; b++) { mine[(size_t) ownerOfDigit[b]] += h[b]; gbase[(size_t) b + 1] = gbase[b] + h[b]; }
    unsigned long long *d_mine = ctx->commWs.as<unsigned long long>();
    unsigned long long *d_all = d_mine + W;
    PG_CUDA(cudaMemcpyAsync(d_mine, mine.data(), sizeof(unsigned long long) * W, cudaMemcpyHostToDevice, s));
    PG_NCCL(g_nccl.AllGather(d_mine, d_all, (size_t) W, ncclUint64, comm_of(ctx), s));
    PG_TRY(read_back_big(ctx, all.data(), d_all, sizeof(unsigned long long) * (size_t) W * W));
    unsigned long long maxRecv = 0;
    std::vector<unsigned long long> recvTotal((size_t) W, 0);
    for (int d = 0; d < W; d++) { for (int r = 0; r < W; r++) recvTotal[d] += all[(size_t) r * W + d]; maxRecv = std::max(maxRecv, recvTotal[d]); }
    bool ok = false;
    PG_TRY(ensure_peer_buffers(ctx, which, sizeof(Rec) * (maxRecv + 1), &ok));
    if (!ok) return 0;
    // destination of every digit: owner's buffer + what the lower ranks send there + this rank's digits below it for that owner
    unsigned long long dst[256];
    std::vector<unsigned long long> firstBase((size_t) W, ~0ull);
    for (int b = 0; b < nDigits; b++) if (firstBase[(size_t) ownerOfDigit[b]] == ~0ull) firstBase[(size_t) ownerOfDigit[b]] = gbase[b];
    for (int b = 0; b < 256; b++) {
        if (b >= nDigits) { dst[b] = 0; continue; }
        const int d = ownerOfDigit[b];
        unsigned long long before = 0;
        for (int r = 0; r < me; r++) before += all[(size_t) r * W + d];
        dst[b] = (unsigned long long) ctx->peer[which][d] + (before + gbase[b] - firstBase[(size_t) d]) * sizeof(Rec);
    }
    unsigned long long *d_dst = ctx->commWs.as<unsigned long long>() + 6000;
    PG_CUDA(cudaMemcpyAsync(d_dst, dst, sizeof(dst), cudaMemcpyHostToDevice, s));
    cudaEventRecord(ctx->evXchg[2 * which], s);
    PG_TRY(radix_scatter_peer(src, n, pass, ctx->radixWs.p, ctx->radixWs.cap, s, d_dst, &ctx->launches));
    PG_TRY(stream_barrier(ctx));
    cudaEventRecord(ctx->evXchg[2 * which + 1], s);
    ctx->lastExchangeBytes[which] = (n - mine[(size_t) me]) * sizeof(Rec);
    *nRecv = recvTotal[(size_t) me];
    *done = true;
    return 0;
}

// the received records of exchange #1 -> the job-wide smallest k-mer (nt: assignGroup's first-group quirk, kmermatcher.cpp:463,
// belongs to exactly one group of the whole job, not to one group per rank)
static int shard_global_min_kmer(Context *ctx, const Rec *recs, uint64_t n, bool nt) {
    ctx->useFirstKmerOverride = false;
    if (!nt) return 0;
    cudaStream_t s = ctx->stream;
    unsigned long long *d_min = ctx->small.as<unsigned long long>() + 52;
    PG_CUDA(cudaMemsetAsync(d_min, 0xFF, sizeof(unsigned long long), s));
    if (n) min_kmer_kernel<<<NUM_SMS * 8, 256, 0, s>>>(recs, n, ~(1ULL << 63), d_min);
    PG_NCCL(g_nccl.AllReduce(d_min, d_min, 1, ncclUint64, ncclMin, comm_of(ctx), s));
    ctx->useFirstKmerOverride = true;
    ctx->launches++;
    return 0;
}

void km_min_kmer_slot(Context *ctx, unsigned long long **d_min) {
    if (ctx->useFirstKmerOverride) *d_min = ctx->small.as<unsigned long long>() + 52;
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_comm_unique_id(void *id) {
    PG_CHECK(id, "pg_comm_unique_id: null argument");
    PG_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) <= PG_COMM_ID_BYTES, "ncclUniqueId does not fit PG_COMM_ID_BYTES");
    ncclUniqueId u;
    PG_NCCL(g_nccl.GetUniqueId(&u));
    memset(id, 0, PG_COMM_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return 0;
}

int pg_comm_init(pg_context *ctx, int rank, int world, const void *id) {
    PG_CHECK(ctx && id, "pg_comm_init: null argument");
    PG_CHECK(world >= 1 && world <= 16 && rank >= 0 && rank < world, "pg_comm_init: need 0 <= rank < world <= 16");
    PG_CHECK(ctx->comm == nullptr, "pg_comm_init: the context already has a communicator");
    PG_TRY(load_nccl());
    cudaSetDevice(ctx->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    PG_NCCL(g_nccl.CommInitRank(&c, world, u, rank));
    ctx->comm = c; ctx->rank = rank; ctx->world = world;
    if (const char *e = getenv("PLASS_B200_SHARD_P2P")) ctx->p2pDisabled = atoi(e) == 0;     // 0: NCCL send / recv exchanges
    for (int i = 0; i < 4; i++) PG_CUDA(cudaEventCreate(&ctx->evXchg[i]));
    return 0;
}

int pg_comm_destroy(pg_context *ctx) {
    PG_CHECK(ctx, "pg_comm_destroy: null argument");
    if (!ctx->comm) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int w = 0; w < 2; w++) { close_peer_mappings(ctx, w); ctx->xr[w].release(); }
    g_nccl.CommDestroy(comm_of(ctx));
    ctx->comm = nullptr; ctx->world = 1; ctx->rank = 0;
    for (int i = 0; i < 4; i++) { if (ctx->evXchg[i]) cudaEventDestroy(ctx->evXchg[i]); ctx->evXchg[i] = nullptr; }
    ctx->commWs.release();
    return 0;
}

// Collective: gives the peer-mapped receive buffers of the fused exchanges back (they are re-created by the next
// pg_shard_iteration).  For callers that need the memory in between, e.g. a single-GPU run of the whole job on one rank.
int pg_shard_release_buffers(pg_context *ctx) {
    PG_CHECK(ctx && ctx->comm, "pg_shard_release_buffers: null argument / no communicator");
    cudaSetDevice(ctx->device);
    PG_TRY(ctx->commWs.reserve(sizeof(unsigned long long) * 16384));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    PG_TRY(stream_barrier(ctx));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int w = 0; w < 2; w++) { close_peer_mappings(ctx, w); ctx->xr[w].release(); }
    PG_TRY(stream_barrier(ctx));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pg_comm_rank(const pg_context *ctx) { return ctx ? ctx->rank : 0; }
int pg_comm_world(const pg_context *ctx) { return ctx ? ctx->world : 1; }

// Replicates a sequence DB that lives on `root` in every rank's HBM (NVLink broadcast of the four arrays).
int pg_shard_broadcast_db(pg_context *ctx, const pg_seqdb *db_on_root, int root, pg_seqdb **out) {
    PG_CHECK(ctx && out && ctx->comm, "pg_shard_broadcast_db: null argument / no communicator");
    PG_CHECK(root >= 0 && root < ctx->world && (ctx->rank != root || db_on_root), "pg_shard_broadcast_db: the root must supply the DB");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    if (ctx->rank == root) PG_TRY(db_ready(ctx, db_on_root));
    PG_TRY(ctx->commWs.reserve(sizeof(unsigned long long) * 8192));
    unsigned long long *d_meta = ctx->commWs.as<unsigned long long>();
    unsigned long long meta[4] = {0, 0, 0, 0};
    if (ctx->rank == root) { meta[0] = db_on_root->n; meta[1] = db_on_root->data_bytes; meta[2] = (unsigned long long) db_on_root->dbtype; }
    PG_CUDA(cudaMemcpyAsync(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    PG_NCCL(g_nccl.Broadcast(d_meta, d_meta, 4, ncclUint64, root, comm_of(ctx), s));
    PG_TRY(read_back(ctx, meta, d_meta, sizeof(meta)));
    pg_seqdb *db = new pg_seqdb();
    db->n = meta[0]; db->data_bytes = meta[1]; db->dbtype = (int) meta[2];
    PG_CUDA(cudaMallocAsync(&db->data, db->data_bytes + 16, s));
    PG_CUDA(cudaMallocAsync(&db->offsets, sizeof(unsigned long long) * (db->n + 1), s));
    PG_CUDA(cudaMallocAsync(&db->lens, sizeof(unsigned) * (db->n + 1), s));
    PG_CUDA(cudaMallocAsync(&db->keys, sizeof(unsigned) * (db->n + 1), s));
    const bool isRoot = ctx->rank == root;
    PG_NCCL(g_nccl.GroupStart());
    PG_NCCL(g_nccl.Broadcast(isRoot ? db_on_root->data : db->data, db->data, db->data_bytes, ncclUint8, root, comm_of(ctx), s));
    PG_NCCL(g_nccl.Broadcast(isRoot ? (void *) db_on_root->offsets : (void *) db->offsets, db->offsets, db->n, ncclUint64, root, comm_of(ctx), s));
    PG_NCCL(g_nccl.Broadcast(isRoot ? db_on_root->lens : db->lens, db->lens, db->n, ncclUint32, root, comm_of(ctx), s));
    PG_NCCL(g_nccl.Broadcast(isRoot ? db_on_root->keys : db->keys, db->keys, db->n, ncclUint32, root, comm_of(ctx), s));
    PG_NCCL(g_nccl.GroupEnd());
    if (seqdb_finalize(ctx, db)) { seqdb_release(db, s); return 1; }
    *out = db;
    return 0;
}

// All-gather of the ranks' DB slices into one replicated DB: rank r holds the entries of a contiguous key range, the
// ranges ascend with the rank (what pg_shard_iteration returns, or what every rank uploaded of a host DB), the result is
// the concatenation in rank order with offsets rebased.
int pg_shard_allgather_db(pg_context *ctx, const pg_seqdb *slice, pg_seqdb **out) {
    PG_CHECK(ctx && slice && out && ctx->comm, "pg_shard_allgather_db: null argument / no communicator");
    cudaSetDevice(ctx->device);
    PG_TRY(db_ready(ctx, slice));
    cudaStream_t s = ctx->stream;
    const int W = ctx->world;
    PG_TRY(ctx->commWs.reserve(sizeof(unsigned long long) * 8192));
    unsigned long long *d_mine = ctx->commWs.as<unsigned long long>(), *d_all = d_mine + 2;
    unsigned long long mine[2] = {slice->n, slice->data_bytes};
    std::vector<unsigned long long> all((size_t) 2 * W);
    PG_CUDA(cudaMemcpyAsync(d_mine, mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    PG_NCCL(g_nccl.AllGather(d_mine, d_all, 2, ncclUint64, comm_of(ctx), s));
    PG_TRY(read_back_big(ctx, all.data(), d_all, sizeof(unsigned long long) * 2 * (size_t) W));
    std::vector<unsigned long long> nOff((size_t) W + 1, 0), bOff((size_t) W + 1, 0);
    for (int r = 0; r < W; r++) { nOff[(size_t) r + 1] = nOff[r] + all[2 * (size_t) r]; bOff[(size_t) r + 1] = bOff[r] + all[2 * (size_t) r + 1]; }
    PG_CHECK(nOff[W] < 0xFFFFFFF0ull, "pg_shard_allgather_db: more than 2^32 sequences");
    pg_seqdb *db = new pg_seqdb();
    db->n = nOff[W]; db->data_bytes = bOff[W]; db->dbtype = slice->dbtype;
    PG_CUDA(cudaMallocAsync(&db->data, db->data_bytes + 16, s));
    PG_CUDA(cudaMallocAsync(&db->offsets, sizeof(unsigned long long) * (db->n + 1), s));
    PG_CUDA(cudaMallocAsync(&db->lens, sizeof(unsigned) * (db->n + 1), s));
    PG_CUDA(cudaMallocAsync(&db->keys, sizeof(unsigned) * (db->n + 1), s));
    PG_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < W; r++) {
        const bool me = r == ctx->rank;
        const unsigned long long n = all[2 * (size_t) r], b = all[2 * (size_t) r + 1];
        if (b) PG_NCCL(g_nccl.Broadcast(me ? slice->data : db->data + bOff[r], db->data + bOff[r], b, ncclUint8, r, comm_of(ctx), s));
        if (n) {
            PG_NCCL(g_nccl.Broadcast(me ? (void *) slice->offsets : (void *) (db->offsets + nOff[r]), db->offsets + nOff[r], n, ncclUint64, r, comm_of(ctx), s));
            PG_NCCL(g_nccl.Broadcast(me ? slice->lens : db->lens + nOff[r], db->lens + nOff[r], n, ncclUint32, r, comm_of(ctx), s));
            PG_NCCL(g_nccl.Broadcast(me ? slice->keys : db->keys + nOff[r], db->keys + nOff[r], n, ncclUint32, r, comm_of(ctx), s));
        }
    }
    PG_NCCL(g_nccl.GroupEnd());
    for (int r = 1; r < W; r++)
        if (all[2 * (size_t) r] && bOff[r]) rebase_offsets_kernel<<<NUM_SMS * 2, 256, 0, s>>>(db->offsets, nOff[r], nOff[(size_t) r + 1], bOff[r]);
    PG_CUDA(cudaGetLastError());
    if (seqdb_finalize(ctx, db)) { seqdb_release(db, s); return 1; }
    *out = db;
    return 0;
}

}  // extern "C"

// pg_shard_iteration with the two exchanges fused into their partition passes (peer memory).  *done = false (nothing changed
// that matters) if peer mappings cannot be had; the caller then runs the NCCL send / recv variant.
static int shard_iteration_fused(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp, const pg_ex_params *ep,
                                 pg_seqdb **out_slice, uint32_t *own_lo, uint32_t *own_hi, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns,
                                 bool *done) {
    *done = false;
    cudaStream_t s = ctx->stream;
    const int W = ctx->world, me = ctx->rank;
    const bool nt = db->dbtype == PG_DBTYPE_NUCLEOTIDES;
    // phase 0: k-mer records of this rank's slice of the sequences; partition by k-mer owner + exchange #1 in one pass
    begin_call(ctx);
    uint64_t nRec = 0, nRecv = 0;
    PG_TRY(km_shard_extract_only(ctx, db, kp, me, W, &nRec));
    end_shard_phase(ctx, true);
    begin_call(ctx);
    {
        RadixPlan plan; plan.npasses = 0;
        plan_add_hash_bits(plan, nt ? ~(1ULL << 63) : ~0ULL, 56, 64);      // the same owner function as km_shard_extract: top byte of mix64
        int owner[256];
        for (int b = 0; b < 256; b++) owner[b] = (int) ((unsigned) b * (unsigned) W / 256u);
        bool ok = false;
        PG_TRY(exchange_fused(ctx, ctx->recA.as<Rec>(), nRec, plan.pass[0], 256, owner, 0, &nRecv, &ok));
        if (!ok) { end_shard_phase(ctx, false); return 0; }
    }
    // phase 1 on the received records: the receive buffer stands in for the first record buffer (no copy) and the extraction's
    // buffer, dead after the scatter, for the second -- the fused path needs three record-sized buffers, not four
    DevBuf bufA = ctx->recA, bufB = ctx->recB;
    void *const r1 = ctx->xr[0].p;
    ctx->recA = ctx->xr[0]; ctx->recB = bufA;
    int rc = shard_global_min_kmer(ctx, ctx->recA.as<Rec>(), nRecv, nt);
    std::vector<uint64_t> hist(PG_SHARD_HIST_BINS);
    if (rc == 0) rc = km_shard_group(ctx, db, kp, ctx->recA.p, nRecv, hist.data());
    ctx->useFirstKmerOverride = false;
    uint64_t nRecv2 = 0;
    if (rc == 0) {
        unsigned long long *d_hist = ctx->commWs.as<unsigned long long>() + 1024;
        cudaMemcpyAsync(d_hist, hist.data(), sizeof(unsigned long long) * PG_SHARD_HIST_BINS, cudaMemcpyHostToDevice, s);
        rc = (g_nccl.AllReduce(d_hist, d_hist, PG_SHARD_HIST_BINS, ncclUint64, ncclSum, comm_of(ctx), s) == ncclSuccess) ? 0 : 1;
        cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * PG_SHARD_HIST_BINS, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
    }
    if (rc == 0) {
        balanced_bounds((const unsigned long long *) hist.data(), PG_SHARD_HIST_BINS, db->max_key, W, 4.0, ctx->lastBounds);
        // partition by the representative's owner + exchange #2 in one pass
        Rec *pairs = nullptr; uint64_t nPairs = 0;
        km_shard_pairs_location(ctx, &pairs, &nPairs);
        unsigned *d_bounds = (unsigned *) (ctx->small.as<unsigned long long>() + 64);
        cudaMemcpyAsync(d_bounds, ctx->lastBounds, sizeof(unsigned) * (W + 1), cudaMemcpyHostToDevice, s);
        RadixPlan plan; plan.npasses = 0;
        plan_add_interval(plan, d_bounds, (unsigned) W);
        int owner[256];
        for (int b = 0; b < 256; b++) owner[b] = b < W ? b : W - 1;
        bool ok = false;
        rc = exchange_fused(ctx, pairs, nPairs, plan.pass[0], W, owner, 1, &nRecv2, &ok);
        if (rc == 0 && !ok) rc = 1, pg::set_error("multi-GPU: peer memory became unavailable between the two exchanges of a step");
    }
    ctx->xr[0] = ctx->recA; bufA = ctx->recB;                 // (the scratch side may have been re-allocated)
    ctx->recA = bufA; ctx->recB = bufB;
    if (rc == 0 && ctx->xr[0].p != r1) { rc = 1; pg::set_error("multi-GPU: the mapped receive buffer of exchange #1 was re-allocated inside the step"); }
    if (rc) return rc;
    end_shard_phase(ctx, false);
    // phase 2: sort #2 + best diagonal + rescore + extension of the owned queries, the pair buffer standing in for recA, the
    // extraction's buffer for recB
    begin_call(ctx);
    const uint32_t lo = ctx->lastBounds[me], hi = ctx->lastBounds[me + 1];
    void *const r2 = ctx->xr[1].p;
    ctx->recA = ctx->xr[1]; ctx->recB = bufA;
    ctx->ownLo = lo; ctx->ownHi = hi;
    pg_hit *dHits = nullptr; uint64_t nH = 0;
    rc = km_shard_reduce(ctx, db, ctx->recA.p, nRecv2, &dHits, &nH);
    if (rc == 0 && hits && n_hits) { rc = hits_to_host_overlapped(ctx, dHits, nH, hits); *n_hits = nH; }
    pg_aln *dAlns = nullptr; uint64_t nA = 0;
    unsigned char *dExt = nullptr;
    if (rc == 0) rc = rs_run(ctx, db, dHits, nH, rp, &dAlns, &nA);
    if (rc == 0 && alns && n_alns) { rc = alns_to_host_overlapped(ctx, dAlns, nA, alns); *n_alns = nA; }
    if (rc == 0) rc = ex_run(ctx, db, dAlns, nA, ep, out_slice, &dExt);
    ctx->ownLo = 0; ctx->ownHi = 0xFFFFFFFFu;
    ctx->xr[1] = ctx->recA; bufA = ctx->recB;
    ctx->recA = bufA; ctx->recB = bufB;
    if (rc == 0 && ctx->xr[1].p != r2) { rc = 1; pg::set_error("multi-GPU: the mapped receive buffer of exchange #2 was re-allocated inside the step"); }
    if (rc != 0) { cudaStreamSynchronize(ctx->copyStream); return rc; }
    cudaFreeAsync(dExt, s);
    end_shard_phase(ctx, false);
    {
        float ms = 0;
        ctx->timings.exchange_ms = 0;
        for (int w = 0; w < 2; w++) {
            if (cudaEventElapsedTime(&ms, ctx->evXchg[2 * w], ctx->evXchg[2 * w + 1]) == cudaSuccess) { ctx->lastExchangeMs[w] = ms; ctx->timings.exchange_ms += ms; }
            else cudaGetLastError();
        }
    }
    if (!ctx->asyncResults) PG_CUDA(cudaStreamSynchronize(ctx->copyStream));
    if (own_lo) *own_lo = lo;
    if (own_hi) *own_hi = hi;
    *done = true;
    return 0;
}

extern "C" {

// One whole assemble iteration over `world` GPUs.  `db` is the replicated sequence DB; out_slice receives the new entries
// of the keys this rank owns ([own_lo, own_hi), ascending with the rank; pg_shard_allgather_db rebuilds the replicated DB
// for the next iteration); hits / alns (optional, pinned host) are this rank's share of pref_N / aln_N.
int pg_shard_iteration(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp, const pg_ex_params *ep,
                       pg_seqdb **out_slice, uint32_t *own_lo, uint32_t *own_hi, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns) {
    PG_CHECK(ctx && db && kp && rp && ep && out_slice && ctx->comm, "pg_shard_iteration: null argument / no communicator (pg_comm_init)");
    PG_TRY(db_ready(ctx, db));
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const int W = ctx->world, me = ctx->rank;
    const bool nt = db->dbtype == PG_DBTYPE_NUCLEOTIDES;
    std::vector<uint64_t> counts((size_t) W);
    if (!ctx->p2pDisabled) {
        bool done = false;
        PG_TRY(shard_iteration_fused(ctx, db, kp, rp, ep, out_slice, own_lo, own_hi, hits, n_hits, alns, n_alns, &done));
        if (done) return 0;
        ctx->p2pDisabled = true;          // peer memory is not available between these devices / processes: NCCL send / recv from now on
    }
    // phase 0: k-mer records of this rank's slice of the sequences, partitioned by the rank owning the k-mer
    begin_call(ctx);
    PG_TRY(km_shard_extract(ctx, db, kp, me, W, counts.data()));
    end_shard_phase(ctx, true);
    // exchange #1 into the record buffer that does not hold the send side
    begin_call(ctx);
    {
        const bool sendInA = (ctx->shardPairs == ctx->recA.as<Rec>());
        DevBuf &recvBuf = sendInA ? ctx->recB : ctx->recA;
        uint64_t nRecv = 0;
        PG_TRY(exchange_records(ctx, ctx->shardPairs, counts.data(), recvBuf, &nRecv, 0));
        if (sendInA) std::swap(ctx->recA, ctx->recB);          // km_group reads its input from recA
        PG_TRY(shard_global_min_kmer(ctx, ctx->recA.as<Rec>(), nRecv, nt));
        // phase 1: sort #1 + assignGroup on the k-mers this rank owns; histogram of the pairs over the representative keys
        PG_CUDA(cudaStreamSynchronize(s));                     // the old send buffer (now recB) may be re-allocated by km_group
        std::vector<uint64_t> hist(PG_SHARD_HIST_BINS);
        PG_TRY(km_shard_group(ctx, db, kp, ctx->recA.p, nRecv, hist.data()));
        ctx->useFirstKmerOverride = false;
        // work per slice of the representative key space, summed over the ranks -> equal-work key ranges
        unsigned long long *d_hist = ctx->commWs.as<unsigned long long>() + 1024;
        PG_CUDA(cudaMemcpyAsync(d_hist, hist.data(), sizeof(unsigned long long) * PG_SHARD_HIST_BINS, cudaMemcpyHostToDevice, s));
        PG_NCCL(g_nccl.AllReduce(d_hist, d_hist, PG_SHARD_HIST_BINS, ncclUint64, ncclSum, comm_of(ctx), s));
        PG_CUDA(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * PG_SHARD_HIST_BINS, cudaMemcpyDeviceToHost, s));
        PG_CUDA(cudaStreamSynchronize(s));
        balanced_bounds((const unsigned long long *) hist.data(), PG_SHARD_HIST_BINS, db->max_key, W, 4.0, ctx->lastBounds);
        PG_TRY(km_shard_route(ctx, W, ctx->lastBounds, counts.data()));
    }
    end_shard_phase(ctx, false);
    // exchange #2: pair records -> owner of the representative; then sort #2 + best diagonal + rescore + extension
    begin_call(ctx);
    const uint32_t lo = ctx->lastBounds[me], hi = ctx->lastBounds[me + 1];
    {
        const bool sendInA = (ctx->shardPairs == ctx->recA.as<Rec>());
        DevBuf &recvBuf = sendInA ? ctx->recB : ctx->recA;
        uint64_t nRecv = 0;
        PG_TRY(exchange_records(ctx, ctx->shardPairs, counts.data(), recvBuf, &nRecv, 1));
        if (sendInA) std::swap(ctx->recA, ctx->recB);
        PG_CUDA(cudaStreamSynchronize(s));
        ctx->ownLo = lo; ctx->ownHi = hi;
        pg_hit *dHits = nullptr; uint64_t nH = 0;
        int rc = km_shard_reduce(ctx, db, ctx->recA.p, nRecv, &dHits, &nH);
        if (rc == 0 && hits && n_hits) { rc = hits_to_host_overlapped(ctx, dHits, nH, hits); *n_hits = nH; }
        pg_aln *dAlns = nullptr; uint64_t nA = 0;
        unsigned char *dExt = nullptr;
        if (rc == 0) rc = rs_run(ctx, db, dHits, nH, rp, &dAlns, &nA);
        if (rc == 0 && alns && n_alns) { rc = alns_to_host_overlapped(ctx, dAlns, nA, alns); *n_alns = nA; }
        if (rc == 0) rc = ex_run(ctx, db, dAlns, nA, ep, out_slice, &dExt);
        ctx->ownLo = 0; ctx->ownHi = 0xFFFFFFFFu;
        if (rc != 0) { cudaStreamSynchronize(ctx->copyStream); return rc; }
        cudaFreeAsync(dExt, s);
    }
    end_shard_phase(ctx, false);
    {
        float ms = 0;
        ctx->timings.exchange_ms = 0;
        for (int w = 0; w < 2; w++) {
            if (cudaEventElapsedTime(&ms, ctx->evXchg[2 * w], ctx->evXchg[2 * w + 1]) == cudaSuccess) { ctx->lastExchangeMs[w] = ms; ctx->timings.exchange_ms += ms; }
            else cudaGetLastError();
        }
    }
    if (!ctx->asyncResults) PG_CUDA(cudaStreamSynchronize(ctx->copyStream));   // pg_set_async_results: the caller waits on a ticket
    if (own_lo) *own_lo = lo;
    if (own_hi) *own_hi = hi;
    return 0;
}

// host logic of the equal-work key ranges, exported for the CPU tests (no device involved)
int pg_shard_balanced_bounds(const uint64_t *hist, int bins, uint32_t max_key, int world, double per_key_weight, uint32_t *bounds) {
    PG_CHECK(hist && bounds && bins > 0 && world >= 1 && world <= 256, "pg_shard_balanced_bounds: bad argument");
    balanced_bounds((const unsigned long long *) hist, bins, max_key, world, per_key_weight, bounds);
    return 0;
}

// per-exchange figures of the last pg_shard_iteration: ms[2], bytes sent to other ranks[2]
int pg_shard_exchange_stats(const pg_context *ctx, float *ms, uint64_t *bytes) {
    PG_CHECK(ctx && ms && bytes, "pg_shard_exchange_stats: null argument");
    for (int w = 0; w < 2; w++) { ms[w] = ctx->lastExchangeMs[w]; bytes[w] = ctx->lastExchangeBytes[w]; }
    return 0;
}

}  // extern "C"
