// pg_api.cu -- the C ABI of libplassgpu.so (declared in include/plassgpu.h).
#include "pg_internal.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

namespace pg {
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }

// Device arrays of sequence DBs come from the stream-ordered pool (cudaMallocAsync; the pool keeps its memory,
// see pg_init), so producing a new DB every iteration does not pay cudaMalloc / cudaFree each time.
void seqdb_release(pg_seqdb *db, cudaStream_t s) {
    if (!db) return;
    if (db->borrowed) { delete db; return; }
    if (db->data) cudaFreeAsync(db->data, s);
    if (db->offsets) cudaFreeAsync(db->offsets, s);
    if (db->lens) cudaFreeAsync(db->lens, s);
    if (db->keys) cudaFreeAsync(db->keys, s);
    delete db;
}

__global__ void read_back_kernel(unsigned *__restrict__ hostMapped, const unsigned *__restrict__ dev, unsigned words) {
    for (unsigned i = threadIdx.x; i < words; i += blockDim.x) hostMapped[i] = dev[i];
}

int read_back_on(Context *ctx, cudaStream_t stream, void *host, const void *dev, size_t bytes) {
    PG_CHECK(bytes <= 1024 && (bytes & 3) == 0, "read_back: at most 1024 bytes, multiple of 4");
    // the auxiliary stream stages through the second half of the mapped block
    unsigned *stage = (unsigned *) ctx->hostStage + (stream == ctx->stream ? 0 : 256);
    read_back_kernel<<<1, 32, 0, stream>>>(stage, (const unsigned *) dev, (unsigned) (bytes / 4));
    PG_CUDA(cudaStreamSynchronize(stream));
    memcpy(host, stage, bytes);
    return 0;
}

int read_back(Context *ctx, void *host, const void *dev, size_t bytes) { return read_back_on(ctx, ctx->stream, host, dev, bytes); }

__global__ void count_nonzero_kernel(const unsigned char *__restrict__ v, unsigned long long n, unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) c += v[i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// max length, residue count, max key, key density -- one small reduction kernel
__global__ void seqdb_stats_kernel(const unsigned *__restrict__ lens, const unsigned *__restrict__ keys, const unsigned long long *__restrict__ offsets,
                                   unsigned long long n,
                                   unsigned long long *__restrict__ out /* [0] sum(len), [1] maxLen, [2] maxKey, [3] nonDense, [4] not ascending, [5] gaps */) {
    unsigned long long sum = 0, mx = 0, mk = 0, nd = 0, na = 0, gaps = 0;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned l = lens[i], k = keys[i];
        sum += l; mx = max(mx, (unsigned long long) l); mk = max(mk, (unsigned long long) k); nd += (k != (unsigned) i);
        if (i > 0 && keys[i - 1] >= k) na++;
        if (i + 1 < n && offsets[i] + l != offsets[i + 1]) gaps++;
    }
    gaps = __reduce_add_sync(0xFFFFFFFFu, (unsigned) min(gaps, 0xFFFFFFFFull));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        nd += __shfl_xor_sync(0xFFFFFFFFu, nd, o);
        na += __shfl_xor_sync(0xFFFFFFFFu, na, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        mk = max(mk, __shfl_xor_sync(0xFFFFFFFFu, mk, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], sum); atomicMax(&out[1], mx); atomicMax(&out[2], mk); atomicAdd(&out[3], nd);
        if (na) atomicAdd(&out[4], na);
        if (gaps) atomicAdd(&out[5], gaps);
    }
}

// Host-supplied prefilter hits / alignments name sequences by key: every key must exist in the DB before a kernel uses
// find_id()'s answer as an index (a prefilter DB of another sequence DB would otherwise read out of bounds).
template <class T, class KA, class KB>
__global__ void validate_keys_kernel(const T *__restrict__ items, unsigned long long n, const unsigned *__restrict__ keys, unsigned nKeys,
                                     KA keyA, KB keyB, unsigned *__restrict__ bad) {
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x) {
        const T it = items[i];
        if (find_id(keys, nKeys, keyA(it)) == 0xFFFFFFFFu || find_id(keys, nKeys, keyB(it)) == 0xFFFFFFFFu) atomicAdd(bad, 1u);
    }
}
struct HitRep { __device__ unsigned operator()(const pg_hit &h) const { return h.rep; } };
struct HitTarget { __device__ unsigned operator()(const pg_hit &h) const { return h.target; } };
struct AlnQuery { __device__ unsigned operator()(const pg_aln &a) const { return a.query; } };
struct AlnTarget { __device__ unsigned operator()(const pg_aln &a) const { return a.target; } };

template <class T, class KA, class KB>
static int validate_keys(Context *ctx, const pg_seqdb *db, const T *d_items, uint64_t n, KA ka, KB kb, const char *what) {
    if (n == 0) return 0;
    PG_TRY(ctx->small.reserve(4096));
    unsigned *d_bad = (unsigned *) (ctx->small.as<unsigned long long>() + 48);
    PG_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), ctx->stream));
    validate_keys_kernel<<<NUM_SMS * 8, 256, 0, ctx->stream>>>(d_items, n, db->keys, (unsigned) db->n, ka, kb, d_bad);
    unsigned bad = 0;
    PG_TRY(read_back(ctx, &bad, d_bad, sizeof(bad)));
    PG_CHECK(bad == 0, std::string(what) + ": " + std::to_string(bad) + " records name a key that is not in the sequence DB");
    return 0;
}

int seqdb_finalize(Context *ctx, pg_seqdb *db) {
    PG_TRY(ctx->small.reserve(4096));
    unsigned long long *d = ctx->small.as<unsigned long long>() + 16;
    PG_CUDA(cudaMemsetAsync(d, 0, 48, ctx->stream));
    if (db->n) seqdb_stats_kernel<<<NUM_SMS * 2, 256, 0, ctx->stream>>>(db->lens, db->keys, db->offsets, db->n, d);
    unsigned long long h[6];
    PG_TRY(read_back(ctx, h, d, sizeof(h)));
    db->residues = (double) h[0] - 2.0 * (double) db->n;        // DBReader::getAminoAcidDBSize (DBReader.cpp:537-546)
    db->max_seq_len = h[1] >= 2 ? (unsigned) (h[1] - 2) : 0;
    db->max_key = (unsigned) h[2];
    db->dense_keys = (h[3] == 0);
    db->contiguous = (h[5] == 0);
    PG_CHECK(h[4] == 0, "sequence DB: keys must be strictly ascending (index order of a sequence DB)");
    PG_CHECK(db->max_key < 0xFFFFFFF0u, "sequence DB: keys >= 2^32 - 16 are reserved");
    return 0;
}

// First use of a DB whose upload was only enqueued (pg_seqdb_upload_async): order the main stream after the copies and
// compute the statistics every stage relies on.
int db_ready(Context *ctx, const pg_seqdb *cdb) {
    pg_seqdb *db = const_cast<pg_seqdb *>(cdb);
    if (!db || !db->uploadPending) return 0;
    cudaSetDevice(ctx->device);
    PG_CUDA(cudaStreamWaitEvent(ctx->stream, db->evReady, 0));
    db->uploadPending = false;
    cudaEventDestroy(db->evReady);
    db->evReady = nullptr;
    return seqdb_finalize(ctx, db);
}

// Pinned host blocks are expensive to create (cudaMallocHost pins pages, ~0.3 s/GB), so blocks released
// with pg_free_host() are kept in a small process-wide cache and handed out again.
static std::mutex g_pinMutex;
static std::map<void *, size_t> g_pinLive;            // blocks owned by the caller
static std::multimap<size_t, void *> g_pinCache;      // released blocks, by size
static size_t g_pinCachedBytes = 0;
static const size_t PIN_CACHE_LIMIT = 96ull << 30;     // two steps' results in flight at 50 M reads are 2 x 21 GB

int alloc_pinned(size_t bytes, void **out) {
    if (bytes < 64) bytes = 64;
    {
        std::lock_guard<std::mutex> lk(g_pinMutex);
        auto it = g_pinCache.lower_bound(bytes);
        if (it != g_pinCache.end() && it->first <= bytes * 2 + (1 << 20)) {
            *out = it->second;
            g_pinLive[it->second] = it->first;
            g_pinCachedBytes -= it->first;
            g_pinCache.erase(it);
            return 0;
        }
    }
    void *p = nullptr;
    const size_t want = bytes + bytes / 8;            // head-room so that slightly larger results reuse the block
    PG_CUDA(cudaMallocHost(&p, want));
    std::lock_guard<std::mutex> lk(g_pinMutex);
    g_pinLive[p] = want;
    *out = p;
    return 0;
}
static void release_pinned(void *p) {
    std::lock_guard<std::mutex> lk(g_pinMutex);
    auto it = g_pinLive.find(p);
    if (it == g_pinLive.end()) { cudaFreeHost(p); return; }
    const size_t sz = it->second;
    g_pinLive.erase(it);
    if (g_pinCachedBytes + sz <= PIN_CACHE_LIMIT) { g_pinCache.emplace(sz, p); g_pinCachedBytes += sz; }
    else cudaFreeHost(p);
}

template <class T>
static int to_host(cudaStream_t s, const T *d, uint64_t n, T **out) {
    T *h = nullptr;
    PG_TRY(alloc_pinned(sizeof(T) * (n + 1), (void **) &h));
    if (n) PG_CUDA(cudaMemcpyAsync(h, d, sizeof(T) * n, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaStreamSynchronize(s));
    *out = h;
    return 0;
}

// Device -> host copy of a finished stage's result on the context's copy stream, so that the transfer runs under the
// kernels of the following stage (the result buffers ctx->hits / ctx->alns are not written again within the call).
// The caller synchronises ctx->copyStream before it hands the host array out.
template <class T>
static int to_host_overlapped(Context *ctx, const T *d, uint64_t n, T **out, cudaEvent_t copied = nullptr) {
    T *h = nullptr;
    PG_TRY(alloc_pinned(sizeof(T) * (n + 1), (void **) &h));
    *out = h;
    if (n == 0) return 0;
    PG_CUDA(cudaEventRecord(ctx->evCopyReady, ctx->stream));
    PG_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evCopyReady, 0));
    PG_CUDA(cudaMemcpyAsync(h, d, sizeof(T) * n, cudaMemcpyDeviceToHost, ctx->copyStream));
    if (copied) PG_CUDA(cudaEventRecord(copied, ctx->copyStream));
    return 0;
}

static void collect_timings(Context *ctx) {
    pg_timings &t = ctx->timings;
    auto el = [&](int a, int b) { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) != cudaSuccess) { ms = 0; cudaGetLastError(); } return ms; };
    cudaStreamSynchronize(ctx->stream);
    if (ctx->tExtract) t.extract_ms = el(EV_KM_BEGIN, EV_EXTRACT_END);
    if (ctx->tGroup) {
        t.sort1_ms = el(EV_SORT1_BEGIN, EV_SORT1_END);
        t.sort1_scatter_ms = el(EV_SCATTER1_BEGIN, EV_SCATTER1_END);
        t.group_ms = el(EV_SORT1_END, EV_GROUP_END);
    }
    if (ctx->tReduce) {
        t.sort2_ms = el(EV_GROUP_END, EV_SORT2_END);
        t.reduce_ms = el(EV_SORT2_END, EV_REDUCE_END);
    }
    if (ctx->rsRan) t.rescore_ms = el(EV_RS_BEGIN, EV_RS_END);
    if (ctx->exRan) t.extend_ms = el(EV_EX_BEGIN, EV_EX_END);
    t.total_ms = el(EV_TOTAL_BEGIN, EV_TOTAL_END);
    t.kernel_launches = ctx->launches;
}

void begin_call(Context *ctx) {
    cudaSetDevice(ctx->device);
    ctx->tExtract = ctx->tGroup = ctx->tReduce = ctx->rsRan = ctx->exRan = false;
    ctx->launches = 0;
    ctx->rsOut = nullptr; ctx->rsCnt = nullptr; ctx->rsOff = nullptr;
    memset(&ctx->timings, 0, sizeof(ctx->timings));
    cudaEventRecord(ctx->ev[EV_TOTAL_BEGIN], ctx->stream);
}
void end_call(Context *ctx) {
    cudaEventRecord(ctx->ev[EV_TOTAL_END], ctx->stream);
    collect_timings(ctx);
}
// multi-GPU: one step is several calls; the reported timings are the sums over its phases
void end_shard_phase(Context *ctx, bool first) {
    end_call(ctx);
    pg_timings &a = ctx->shardAcc;
    const pg_timings &t = ctx->timings;
    if (first) memset(&a, 0, sizeof(a));
    a.extract_ms += t.extract_ms; a.sort1_ms += t.sort1_ms; a.group_ms += t.group_ms; a.sort2_ms += t.sort2_ms; a.reduce_ms += t.reduce_ms;
    a.rescore_ms += t.rescore_ms; a.extend_ms += t.extend_ms; a.total_ms += t.total_ms; a.sort1_scatter_ms += t.sort1_scatter_ms;
    a.kernel_launches += t.kernel_launches; a.exchange_ms += t.exchange_ms;
    if (t.n_kmer_records) a.n_kmer_records = t.n_kmer_records;
    if (t.n_pair_records) a.n_pair_records = t.n_pair_records;
    if (t.sort1_bytes) a.sort1_bytes = t.sort1_bytes;
    if (t.sort1_passes) a.sort1_passes = t.sort1_passes;
    if (t.n_hits) a.n_hits = t.n_hits;
    if (t.n_alns) a.n_alns = t.n_alns;
    if (t.n_extended) a.n_extended = t.n_extended;
    ctx->timings = a;
}
// (the events guard the device buffers against the next call's writes while an asynchronous copy still reads them)
int hits_to_host_overlapped(Context *ctx, const pg_hit *d, uint64_t n, pg_hit **out) { return to_host_overlapped(ctx, d, n, out, ctx->evHitsCopied); }
int alns_to_host_overlapped(Context *ctx, const pg_aln *d, uint64_t n, pg_aln **out) { return to_host_overlapped(ctx, d, n, out, ctx->evAlnsCopied); }
}  // namespace pg

using namespace pg;

extern "C" {

const char *pg_last_error(void) { return g_error.c_str(); }

int pg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pg_init(int device, pg_context **out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("pg_init: no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
        cudaGetLastError();
        return 1;
    }
    PG_CHECK(device >= 0 && device < n, "pg_init: device index out of range");
    PG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PG_CUDA(cudaGetDeviceProperties(&prop, device));
    PG_CHECK(prop.major == 10, "pg_init: this build targets sm_100a (B200) only");
    pg_context *ctx = new pg_context();
    ctx->device = device;
    ctx->deviceMemBytes = prop.totalGlobalMem;
    PG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    PG_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    PG_CUDA(cudaStreamCreateWithFlags(&ctx->h2dStream, cudaStreamNonBlocking));
    PG_CUDA(cudaEventCreateWithFlags(&ctx->evCopyReady, cudaEventDisableTiming));
    PG_CUDA(cudaEventCreateWithFlags(&ctx->evHitsCopied, cudaEventDisableTiming));
    PG_CUDA(cudaEventCreateWithFlags(&ctx->evAlnsCopied, cudaEventDisableTiming));
    for (int i = 0; i < 16; i++) PG_CUDA(cudaEventCreateWithFlags(&ctx->evTicket[i], cudaEventDisableTiming));
    PG_CUDA(cudaHostAlloc((void **) &ctx->hostStage, 2048, cudaHostAllocMapped | cudaHostAllocPortable));
    {   // high priority: its small kernels get SM slots as soon as CTAs of the main stream's big kernel retire
        int lowest = 0, greatest = 0;
        PG_CUDA(cudaDeviceGetStreamPriorityRange(&lowest, &greatest));
        PG_CUDA(cudaStreamCreateWithPriority(&ctx->auxStream, cudaStreamNonBlocking, greatest));
    }
    PG_CUDA(cudaEventCreateWithFlags(&ctx->evAuxFork, cudaEventDisableTiming));
    PG_CUDA(cudaEventCreateWithFlags(&ctx->evAuxJoin, cudaEventDisableTiming));
    {
        cudaMemPool_t pool;
        PG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;                 // never give memory back to the driver between iterations
        PG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (int i = 0; i < EV_COUNT; i++) PG_CUDA(cudaEventCreate(&ctx->ev[i]));
    memset(&ctx->timings, 0, sizeof(ctx->timings));
    if (const char *e = getenv("PLASS_B200_NO_SCRATCH_ALIAS")) ctx->noScratchAlias = atoi(e) != 0;
    if (const char *e = getenv("PLASS_B200_NO_PREHIST")) ctx->noPreHist = atoi(e) != 0;     // A/B: histogram sweep in front of sort #1
    if (const char *e = getenv("PLASS_B200_BUCKET_TARGET")) { const int b = atoi(e); if (b >= 64 && b <= 1200) ctx->bucketTarget = (unsigned) b; }
    if (const char *e = getenv("PLASS_B200_DIGIT_BITS")) { const int b = atoi(e); if (b >= 8 && b <= 10) ctx->digitBits = b; }
    *out = ctx;
    return 0;
}

void pg_destroy(pg_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) pg_comm_destroy(ctx);
    DevBuf *bufs[] = {&ctx->small, &ctx->lists, &ctx->recA, &ctx->recB, &ctx->radixWs, &ctx->scratch, &ctx->blockCounts, &ctx->hits,
                      &ctx->alnAll, &ctx->alns, &ctx->flags, &ctx->exWork, &ctx->exSegs, &ctx->exMeta, &ctx->exLists, &ctx->ntTab, &ctx->buckets, &ctx->buckets2, &ctx->wideTabs, &ctx->nextWork, &ctx->orfInfo, &ctx->pairAcc, &ctx->spill, &ctx->commWs};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < EV_COUNT; i++) cudaEventDestroy(ctx->ev[i]);
    cudaStreamSynchronize(ctx->copyStream);
    cudaEventDestroy(ctx->evCopyReady);
    cudaEventDestroy(ctx->evHitsCopied); cudaEventDestroy(ctx->evAlnsCopied);
    for (int i = 0; i < 16; i++) cudaEventDestroy(ctx->evTicket[i]);
    cudaFreeHost(ctx->hostStage);
    cudaStreamSynchronize(ctx->auxStream);
    cudaEventDestroy(ctx->evAuxFork); cudaEventDestroy(ctx->evAuxJoin);
    cudaStreamDestroy(ctx->auxStream);
    cudaStreamDestroy(ctx->copyStream);
    cudaStreamSynchronize(ctx->h2dStream);
    cudaStreamDestroy(ctx->h2dStream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Gives the context's cached device buffers (record buffers, scratch, result staging) back to the driver; they are
// re-created on demand.  For callers that alternate between very different problem sizes on one context.
int pg_release_workspace(pg_context *ctx) {
    PG_CHECK(ctx, "pg_release_workspace: null argument");
    cudaSetDevice(ctx->device);
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->copyStream));
    PG_CUDA(cudaStreamSynchronize(ctx->auxStream));
    DevBuf *bufs[] = {&ctx->lists, &ctx->recA, &ctx->recB, &ctx->radixWs, &ctx->scratch, &ctx->blockCounts, &ctx->hits,
                      &ctx->alnAll, &ctx->alns, &ctx->flags, &ctx->exWork, &ctx->exSegs, &ctx->exMeta, &ctx->exLists, &ctx->buckets, &ctx->buckets2, &ctx->wideTabs,
                      &ctx->nextWork, &ctx->orfInfo, &ctx->pairAcc, &ctx->spill};
    for (DevBuf *b : bufs) b->release();
    ctx->rsOut = nullptr; ctx->rsCnt = nullptr; ctx->rsOff = nullptr; ctx->shardPairs = nullptr; ctx->shardPairCount = 0;
    return 0;
}

int pg_get_timings(const pg_context *ctx, pg_timings *out) {
    PG_CHECK(ctx && out, "pg_get_timings: null argument");
    *out = ctx->timings;
    return 0;
}

int pg_seqdb_upload(pg_context *ctx, const pg_seqdb_view *v, pg_seqdb **out) {
    PG_CHECK(ctx && v && out, "pg_seqdb_upload: null argument");
    PG_CHECK(v->dbtype == PG_DBTYPE_AMINO_ACIDS || v->dbtype == PG_DBTYPE_NUCLEOTIDES, "pg_seqdb_upload: dbtype must be amino acids (0) or nucleotides (1)");
    PG_CHECK(v->n < 0xFFFFFFF0ull, "pg_seqdb_upload: more than 2^32 sequences");
    cudaSetDevice(ctx->device);
    pg_seqdb *db = new pg_seqdb();
    db->n = v->n; db->data_bytes = v->data_bytes; db->dbtype = v->dbtype;
    PG_CUDA(cudaMallocAsync(&db->data, v->data_bytes + 16, ctx->stream));
    PG_CUDA(cudaMallocAsync(&db->offsets, sizeof(unsigned long long) * (v->n + 1), ctx->stream));
    PG_CUDA(cudaMallocAsync(&db->lens, sizeof(unsigned) * (v->n + 1), ctx->stream));
    PG_CUDA(cudaMallocAsync(&db->keys, sizeof(unsigned) * (v->n + 1), ctx->stream));
    PG_CUDA(cudaMemcpyAsync(db->data, v->data, v->data_bytes, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(db->offsets, v->offsets, sizeof(unsigned long long) * v->n, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(db->lens, v->lens, sizeof(unsigned) * v->n, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(db->keys, v->keys, sizeof(unsigned) * v->n, cudaMemcpyHostToDevice, ctx->stream));
    if (seqdb_finalize(ctx, db)) { seqdb_release(db, ctx->stream); return 1; }
    *out = db;
    return 0;
}

int pg_seqdb_upload_async(pg_context *ctx, const pg_seqdb_view *v, pg_seqdb **out) {
    PG_CHECK(ctx && v && out, "pg_seqdb_upload_async: null argument");
    PG_CHECK(v->dbtype == PG_DBTYPE_AMINO_ACIDS || v->dbtype == PG_DBTYPE_NUCLEOTIDES, "pg_seqdb_upload_async: dbtype must be amino acids (0) or nucleotides (1)");
    PG_CHECK(v->n < 0xFFFFFFF0ull, "pg_seqdb_upload_async: more than 2^32 sequences");
    cudaSetDevice(ctx->device);
    cudaStream_t hs = ctx->h2dStream;
    pg_seqdb *db = new pg_seqdb();
    db->n = v->n; db->data_bytes = v->data_bytes; db->dbtype = v->dbtype;
    PG_CUDA(cudaMallocAsync(&db->data, v->data_bytes + 16, hs));
    PG_CUDA(cudaMallocAsync(&db->offsets, sizeof(unsigned long long) * (v->n + 1), hs));
    PG_CUDA(cudaMallocAsync(&db->lens, sizeof(unsigned) * (v->n + 1), hs));
    PG_CUDA(cudaMallocAsync(&db->keys, sizeof(unsigned) * (v->n + 1), hs));
    PG_CUDA(cudaMemcpyAsync(db->data, v->data, v->data_bytes, cudaMemcpyHostToDevice, hs));
    PG_CUDA(cudaMemcpyAsync(db->offsets, v->offsets, sizeof(unsigned long long) * v->n, cudaMemcpyHostToDevice, hs));
    PG_CUDA(cudaMemcpyAsync(db->lens, v->lens, sizeof(unsigned) * v->n, cudaMemcpyHostToDevice, hs));
    PG_CUDA(cudaMemcpyAsync(db->keys, v->keys, sizeof(unsigned) * v->n, cudaMemcpyHostToDevice, hs));
    PG_CUDA(cudaEventCreateWithFlags(&db->evReady, cudaEventDisableTiming));
    PG_CUDA(cudaEventRecord(db->evReady, hs));
    db->uploadPending = true;
    *out = db;
    return 0;
}

int pg_seqdb_adopt(pg_context *ctx, const pg_seqdb_view *v, pg_seqdb **out) {
    PG_CHECK(ctx && v && out, "pg_seqdb_adopt: null argument");
    PG_CHECK(v->dbtype == PG_DBTYPE_AMINO_ACIDS || v->dbtype == PG_DBTYPE_NUCLEOTIDES, "pg_seqdb_adopt: dbtype must be amino acids (0) or nucleotides (1)");
    PG_CHECK(v->n < 0xFFFFFFF0ull, "pg_seqdb_adopt: more than 2^32 sequences");
    cudaSetDevice(ctx->device);
    cudaPointerAttributes at;
    PG_CHECK(cudaPointerGetAttributes(&at, v->data) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device == ctx->device,
             "pg_seqdb_adopt: the arrays must live in this context's device memory");
    pg_seqdb *db = new pg_seqdb();
    db->n = v->n; db->data_bytes = v->data_bytes; db->dbtype = v->dbtype; db->borrowed = true;
    db->data = (char *) v->data; db->offsets = (unsigned long long *) v->offsets; db->lens = (unsigned *) v->lens; db->keys = (unsigned *) v->keys;
    if (seqdb_finalize(ctx, db)) { delete db; return 1; }
    *out = db;
    return 0;
}

int pg_seqdb_download(pg_context *ctx, const pg_seqdb *db, char **data, uint64_t *data_bytes, uint64_t **offsets,
                      uint32_t **lens, uint32_t **keys, uint64_t *n) {
    PG_CHECK(ctx && db, "pg_seqdb_download: null argument");
    PG_TRY(db_ready(ctx, db));
    cudaSetDevice(ctx->device);
    if (ctx->asyncResults) {
        // enqueue only: the arrays are complete after pg_results_wait on a ticket taken after this call.  The DB's device
        // arrays are then released on the copy stream (pg_seqdb_free), i.e. after these copies.
        PG_TRY(to_host_overlapped(ctx, db->data, db->data_bytes, data));
        PG_TRY(to_host_overlapped(ctx, (const uint64_t *) db->offsets, db->n, offsets));
        PG_TRY(to_host_overlapped(ctx, db->lens, db->n, lens));
        PG_TRY(to_host_overlapped(ctx, db->keys, db->n, keys));
        const_cast<pg_seqdb *>(db)->downloadPending = true;
        *data_bytes = db->data_bytes; *n = db->n;
        return 0;
    }
    PG_TRY(to_host(ctx->stream, db->data, db->data_bytes, data));
    PG_TRY(to_host(ctx->stream, (const uint64_t *) db->offsets, db->n, offsets));
    PG_TRY(to_host(ctx->stream, db->lens, db->n, lens));
    PG_TRY(to_host(ctx->stream, db->keys, db->n, keys));
    *data_bytes = db->data_bytes; *n = db->n;
    return 0;
}

uint64_t pg_seqdb_size(const pg_seqdb *db) { return db ? db->n : 0; }

void pg_seqdb_free(pg_context *ctx, pg_seqdb *db) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (db && db->uploadPending) {            // uploaded but never used: release after the copies
        cudaStreamWaitEvent(ctx->stream, db->evReady, 0);
        cudaEventDestroy(db->evReady);
        db->evReady = nullptr; db->uploadPending = false;
    }
    // stream-ordered release: after the kernels of the main stream, or after the pending download on the copy stream
    // (every consumer on the main stream was enqueued before the download)
    seqdb_release(db, (db && db->downloadPending) ? ctx->copyStream : ctx->stream);
}

int pg_kmermatch(pg_context *ctx, const pg_seqdb *db, const pg_km_params *p, pg_hit **hits, uint64_t *n_hits) {
    PG_CHECK(ctx && db && p && hits && n_hits, "pg_kmermatch: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    pg_hit *d = nullptr; uint64_t n = 0;
    PG_TRY(km_run(ctx, db, p, &d, &n));
    PG_TRY(to_host(ctx->stream, d, n, hits));
    *n_hits = n;
    end_call(ctx);
    return 0;
}

int pg_rescore(pg_context *ctx, const pg_seqdb *db, const pg_hit *hits, uint64_t n_hits, const pg_rs_params *p,
               pg_aln **alns, uint64_t *n_alns) {
    PG_CHECK(ctx && db && p && alns && n_alns && (hits || n_hits == 0), "pg_rescore: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    PG_TRY(ctx->hits.reserve(sizeof(pg_hit) * (n_hits + 1)));
    if (n_hits) PG_CUDA(cudaMemcpyAsync(ctx->hits.p, hits, sizeof(pg_hit) * n_hits, cudaMemcpyHostToDevice, ctx->stream));
    for (uint64_t i = 1; i < n_hits; i++)
        PG_CHECK(hits[i - 1].rep <= hits[i].rep, "pg_rescore: hits must be ordered by rep (prefilter DB order)");
    PG_TRY(validate_keys(ctx, db, ctx->hits.as<pg_hit>(), n_hits, HitRep(), HitTarget(), "pg_rescore"));
    pg_aln *d = nullptr; uint64_t n = 0;
    PG_TRY(rs_run(ctx, db, ctx->hits.as<pg_hit>(), n_hits, p, &d, &n));
    PG_TRY(to_host(ctx->stream, d, n, alns));
    *n_alns = n;
    end_call(ctx);
    return 0;
}

int pg_extend(pg_context *ctx, const pg_seqdb *db, const pg_aln *alns, uint64_t n_alns, const pg_ex_params *p,
              pg_seqdb **out_db, uint8_t **extended) {
    PG_CHECK(ctx && db && p && out_db && (alns || n_alns == 0), "pg_extend: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    PG_TRY(ctx->alns.reserve(sizeof(pg_aln) * (n_alns + 1)));
    if (n_alns) PG_CUDA(cudaMemcpyAsync(ctx->alns.p, alns, sizeof(pg_aln) * n_alns, cudaMemcpyHostToDevice, ctx->stream));
    for (uint64_t i = 1; i < n_alns; i++)
        PG_CHECK(alns[i - 1].query <= alns[i].query, "pg_extend: alignments must be ordered by query");
    PG_TRY(validate_keys(ctx, db, ctx->alns.as<pg_aln>(), n_alns, AlnQuery(), AlnTarget(), "pg_extend"));
    unsigned char *dExt = nullptr;
    PG_TRY(ex_run(ctx, db, ctx->alns.as<pg_aln>(), n_alns, p, out_db, &dExt));
    if (extended) PG_TRY(to_host(ctx->stream, dExt, (*out_db)->n, extended));
    cudaFreeAsync(dExt, ctx->stream);
    end_call(ctx);
    return 0;
}

int pg_assemble_iteration(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp,
                          const pg_ex_params *ep, pg_seqdb **out_db,
                          pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns) {
    PG_CHECK(ctx && db && kp && rp && ep && out_db, "pg_assemble_iteration: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    pg_hit *dHits = nullptr; uint64_t nH = 0;
    PG_TRY(km_run(ctx, db, kp, &dHits, &nH));
    if (hits && n_hits) { PG_TRY(to_host_overlapped(ctx, dHits, nH, hits, ctx->evHitsCopied)); *n_hits = nH; }        // copied while rescorediagonal runs
    pg_aln *dAlns = nullptr; uint64_t nA = 0;
    PG_TRY(rs_run(ctx, db, dHits, nH, rp, &dAlns, &nA));
    if (alns && n_alns) { PG_TRY(to_host_overlapped(ctx, dAlns, nA, alns, ctx->evAlnsCopied)); *n_alns = nA; }        // copied while the extension runs
    unsigned char *dExt = nullptr;
    PG_TRY(ex_run(ctx, db, dAlns, nA, ep, out_db, &dExt));
    {   // number of new contigs, for the statistics
        unsigned long long *d = ctx->small.as<unsigned long long>() + 40;
        PG_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
        if ((*out_db)->n) count_nonzero_kernel<<<NUM_SMS * 2, 256, 0, ctx->stream>>>(dExt, (*out_db)->n, d);
        unsigned long long c = 0;
        PG_TRY(read_back(ctx, &c, d, sizeof(c)));
        ctx->timings.n_extended = c;
    }
    cudaFreeAsync(dExt, ctx->stream);
    end_call(ctx);
    if (!ctx->asyncResults) PG_CUDA(cudaStreamSynchronize(ctx->copyStream));
    return 0;
}

// plass STEP 0 in one call (data/assemble.sh:88-151 with STEP = 0): kmermatcher -> rescorediagonal -> findassemblystart ->
// kmermatcher -> rescorediagonal -> assembleresults, everything between the input DB and assembly_0 stays in HBM.
int pg_assemble_step0(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp, const pg_ex_params *ep,
                      pg_seqdb **corrected_db, pg_seqdb **out_db, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns) {
    PG_CHECK(ctx && db && kp && rp && ep && out_db, "pg_assemble_step0: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    pg_hit *dHits = nullptr; uint64_t nH = 0;
    pg_aln *dAlns = nullptr; uint64_t nA = 0;
    PG_TRY(km_run(ctx, db, kp, &dHits, &nH));                       // pref_0
    PG_TRY(rs_run(ctx, db, dHits, nH, rp, &dAlns, &nA));            // aln_0
    pg_seqdb *corr = nullptr;
    PG_TRY(fs_run(ctx, db, dAlns, nA, &corr, nullptr));             // corrected_seqs
    const uint64_t launchesFirst = ctx->launches;
    int rc = km_run(ctx, corr, kp, &dHits, &nH);                    // pref_corrected_0
    if (rc == 0 && hits && n_hits) { rc = to_host_overlapped(ctx, dHits, nH, hits, ctx->evHitsCopied); *n_hits = nH; }
    if (rc == 0) rc = rs_run(ctx, corr, dHits, nH, rp, &dAlns, &nA);    // aln_corrected_0
    if (rc == 0 && alns && n_alns) { rc = to_host_overlapped(ctx, dAlns, nA, alns, ctx->evAlnsCopied); *n_alns = nA; }
    unsigned char *dExt = nullptr;
    if (rc == 0) rc = ex_run(ctx, corr, dAlns, nA, ep, out_db, &dExt);  // assembly_0
    if (rc != 0) { cudaStreamSynchronize(ctx->copyStream); seqdb_release(corr, ctx->stream); return rc; }
    cudaFreeAsync(dExt, ctx->stream);
    (void) launchesFirst;
    end_call(ctx);
    if (!ctx->asyncResults) PG_CUDA(cudaStreamSynchronize(ctx->copyStream));
    if (corrected_db) *corrected_db = corr; else seqdb_release(corr, ctx->stream);
    return 0;
}

int pg_findassemblystart(pg_context *ctx, const pg_seqdb *db, const pg_aln *alns, uint64_t n_alns, pg_seqdb **out_db, int32_t **add_stop) {
    PG_CHECK(ctx && db && out_db && (alns || n_alns == 0), "pg_findassemblystart: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    PG_TRY(ctx->alns.reserve(sizeof(pg_aln) * (n_alns + 1)));
    if (n_alns) PG_CUDA(cudaMemcpyAsync(ctx->alns.p, alns, sizeof(pg_aln) * n_alns, cudaMemcpyHostToDevice, ctx->stream));
    for (uint64_t i = 1; i < n_alns; i++)
        PG_CHECK(alns[i - 1].query <= alns[i].query, "pg_findassemblystart: alignments must be ordered by query");
    PG_TRY(validate_keys(ctx, db, ctx->alns.as<pg_aln>(), n_alns, AlnQuery(), AlnTarget(), "pg_findassemblystart"));
    int *dStop = nullptr;
    PG_TRY(fs_run(ctx, db, ctx->alns.as<pg_aln>(), n_alns, out_db, &dStop));
    if (add_stop) PG_TRY(to_host(ctx->stream, dStop, db->n, add_stop));
    end_call(ctx);
    return 0;
}

int pg_extractorfs(pg_context *ctx, const pg_seqdb *db, const pg_orf_params *p, int translate, pg_seqdb **out_db, uint32_t **orf_info) {
    PG_CHECK(ctx && db && p && out_db, "pg_extractorfs: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    unsigned *dInfo = nullptr;
    PG_TRY(orf_run(ctx, db, p, translate, out_db, &dInfo));
    if (orf_info) PG_TRY(to_host(ctx->stream, dInfo, 4 * (*out_db)->n, orf_info));
    end_call(ctx);
    return 0;
}

int pg_translatenucs(pg_context *ctx, const pg_seqdb *db, const uint8_t *flags, int translation_table, pg_seqdb **out_db) {
    PG_CHECK(ctx && db && out_db, "pg_translatenucs: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    unsigned char *dFlags = nullptr;
    if (flags && db->n) {
        PG_TRY(ctx->flags.reserve(db->n + 16));
        dFlags = ctx->flags.as<unsigned char>();
        PG_CUDA(cudaMemcpyAsync(dFlags, flags, db->n, cudaMemcpyHostToDevice, ctx->stream));
    }
    PG_TRY(tn_run(ctx, db, dFlags, translation_table, out_db));
    end_call(ctx);
    return 0;
}

// concatdbs of two sequence DBs without --preserve-keys (lib/mmseqs/src/util/concatdbs.cpp): a's entries keep their
// order and get the keys 0..a.n-1, b's follow with a.n..a.n+b.n-1
__global__ void concat_meta_kernel(const unsigned long long *__restrict__ aOff, const unsigned *__restrict__ aLen, unsigned long long an,
                                   const unsigned long long *__restrict__ bOff, const unsigned *__restrict__ bLen, unsigned long long bn,
                                   unsigned long long aBytes, unsigned long long *__restrict__ oOff, unsigned *__restrict__ oLen, unsigned *__restrict__ oKey) {
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < an + bn; i += (unsigned long long) gridDim.x * blockDim.x) {
        if (i < an) { oOff[i] = aOff[i]; oLen[i] = aLen[i]; }
        else { oOff[i] = aBytes + bOff[i - an]; oLen[i] = bLen[i - an]; }
        oKey[i] = (unsigned) i;
    }
}

int pg_seqdb_concat(pg_context *ctx, const pg_seqdb *a, const pg_seqdb *b, pg_seqdb **out_db) {
    PG_CHECK(ctx && a && b && out_db, "pg_seqdb_concat: null argument");
    PG_CHECK(a->dbtype == b->dbtype, "pg_seqdb_concat: the two DBs must have the same type");
    PG_TRY(db_ready(ctx, a));
    PG_TRY(db_ready(ctx, b));
    begin_call(ctx);
    cudaStream_t s = ctx->stream;
    const uint64_t n = a->n + b->n;
    PG_CHECK(n < 0xFFFFFFF0ull, "pg_seqdb_concat: more than 2^32 sequences");
    pg_seqdb *out = new pg_seqdb();
    out->n = n; out->data_bytes = a->data_bytes + b->data_bytes; out->dbtype = a->dbtype;
    PG_CUDA(cudaMallocAsync(&out->data, out->data_bytes + 16, s));
    PG_CUDA(cudaMallocAsync(&out->offsets, sizeof(unsigned long long) * (n + 1), s));
    PG_CUDA(cudaMallocAsync(&out->lens, sizeof(unsigned) * (n + 1), s));
    PG_CUDA(cudaMallocAsync(&out->keys, sizeof(unsigned) * (n + 1), s));
    if (a->data_bytes) PG_CUDA(cudaMemcpyAsync(out->data, a->data, a->data_bytes, cudaMemcpyDeviceToDevice, s));
    if (b->data_bytes) PG_CUDA(cudaMemcpyAsync(out->data + a->data_bytes, b->data, b->data_bytes, cudaMemcpyDeviceToDevice, s));
    if (n) {
        concat_meta_kernel<<<NUM_SMS * 4, 256, 0, s>>>(a->offsets, a->lens, a->n, b->offsets, b->lens, b->n, a->data_bytes, out->offsets, out->lens, out->keys);
        ctx->launches++;
    }
    PG_CUDA(cudaGetLastError());
    PG_TRY(seqdb_finalize(ctx, out));
    end_call(ctx);
    *out_db = out;
    return 0;
}

int pg_cyclecheck(pg_context *ctx, const pg_seqdb *db, int max_seq_len, uint32_t **split) {
    PG_CHECK(ctx && db && split, "pg_cyclecheck: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    unsigned *d = nullptr;
    PG_TRY(cc_run(ctx, db, max_seq_len, 22 /* setCycleCheckDefaults, cyclecheck.cpp:26-29 */, &d));
    PG_TRY(to_host(ctx->stream, d, db->n, split));
    end_call(ctx);
    return 0;
}

int pg_set_async_results(pg_context *ctx, int on) {
    PG_CHECK(ctx, "pg_set_async_results: null argument");
    cudaSetDevice(ctx->device);
    if (!on) PG_CUDA(cudaStreamSynchronize(ctx->copyStream));
    ctx->asyncResults = on != 0;
    return 0;
}

int pg_results_ticket(pg_context *ctx, uint64_t *ticket) {
    PG_CHECK(ctx && ticket, "pg_results_ticket: null argument");
    cudaSetDevice(ctx->device);
    const uint64_t t = ctx->nextTicket++;
    PG_CUDA(cudaEventRecord(ctx->evTicket[t % 16], ctx->copyStream));
    *ticket = t;
    return 0;
}

int pg_results_wait(pg_context *ctx, uint64_t ticket) {
    PG_CHECK(ctx, "pg_results_wait: null argument");
    PG_CHECK(ticket < ctx->nextTicket, "pg_results_wait: unknown ticket");
    cudaSetDevice(ctx->device);
    // a slot that was recorded again later marks a later position of the same stream: waiting for it covers the ticket
    PG_CUDA(cudaEventSynchronize(ctx->evTicket[ticket % 16]));
    return 0;
}

void pg_free_host(void *p) { if (p) release_pinned(p); }

/* for a DB from pg_seqdb_upload_async the key statistics exist after its first use in a compute call */
uint32_t pg_seqdb_max_key(const pg_seqdb *db) { return db ? db->max_key : 0; }

void pg_shard_owner_range(uint32_t max_key, int rank, int world, uint32_t *lo, uint32_t *hi) {
    const unsigned long long per = ((unsigned long long) max_key + (unsigned long long) world) / (unsigned long long) world;   // = keysPerRank of tag_owner_kernel
    *lo = (uint32_t) std::min<unsigned long long>(per * (unsigned long long) rank, 0xFFFFFFFFull);
    *hi = (rank == world - 1) ? 0xFFFFFFFFu : (uint32_t) std::min<unsigned long long>(per * (unsigned long long) (rank + 1), 0xFFFFFFFFull);
}

int pg_shard_pairs(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, int world, uint64_t *counts) {
    PG_CHECK(ctx && db && kp && counts, "pg_shard_pairs: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    PG_TRY(km_shard_pairs(ctx, db, kp, world, counts));
    end_shard_phase(ctx, true);
    return 0;
}

int pg_shard_extract(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, int rank, int world, uint64_t *counts) {
    PG_CHECK(ctx && db && kp && counts, "pg_shard_extract: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    PG_TRY(km_shard_extract(ctx, db, kp, rank, world, counts));
    end_shard_phase(ctx, true);
    return 0;
}

int pg_shard_group(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const void *device_records, uint64_t n_records,
                   uint64_t *rep_hist) {
    PG_CHECK(ctx && db && kp && rep_hist && (device_records || n_records == 0), "pg_shard_group: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    PG_TRY(km_shard_group(ctx, db, kp, device_records, n_records, rep_hist));
    end_shard_phase(ctx, false);
    return 0;
}

int pg_shard_route(pg_context *ctx, int world, const uint32_t *bounds, uint64_t *counts) {
    PG_CHECK(ctx && bounds && counts, "pg_shard_route: null argument");
    begin_call(ctx);
    PG_TRY(km_shard_route(ctx, world, bounds, counts));
    end_shard_phase(ctx, false);
    return 0;
}

int pg_shard_export(pg_context *ctx, void *device_dst, uint64_t n_records) {
    PG_CHECK(ctx && (device_dst || n_records == 0), "pg_shard_export: null argument");
    PG_CHECK(n_records == ctx->shardPairCount, "pg_shard_export: record count does not match the preceding phase");
    cudaSetDevice(ctx->device);
    if (n_records) PG_CUDA(cudaMemcpyAsync(device_dst, ctx->shardPairs, sizeof(Rec) * n_records, cudaMemcpyDeviceToDevice, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pg_shard_finish(pg_context *ctx, const pg_seqdb *db, const void *device_pairs, uint64_t n_pairs,
                    uint32_t own_lo, uint32_t own_hi, const pg_rs_params *rp, const pg_ex_params *ep,
                    pg_seqdb **out_db, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns) {
    PG_CHECK(ctx && db && rp && ep && out_db && (device_pairs || n_pairs == 0), "pg_shard_finish: null argument");
    PG_TRY(db_ready(ctx, db));
    begin_call(ctx);
    ctx->ownLo = own_lo; ctx->ownHi = own_hi;
    pg_hit *dHits = nullptr; uint64_t nH = 0;
    int rc = km_shard_reduce(ctx, db, device_pairs, n_pairs, &dHits, &nH);
    if (rc == 0 && hits && n_hits) { rc = to_host_overlapped(ctx, dHits, nH, hits); *n_hits = nH; }
    pg_aln *dAlns = nullptr; uint64_t nA = 0;
    unsigned char *dExt = nullptr;
    if (rc == 0) rc = rs_run(ctx, db, dHits, nH, rp, &dAlns, &nA);
    if (rc == 0 && alns && n_alns) { rc = to_host_overlapped(ctx, dAlns, nA, alns); *n_alns = nA; }
    if (rc == 0) rc = ex_run(ctx, db, dAlns, nA, ep, out_db, &dExt);
    ctx->ownLo = 0; ctx->ownHi = 0xFFFFFFFFu;
    if (rc != 0) { cudaStreamSynchronize(ctx->copyStream); return rc; }
    cudaFreeAsync(dExt, ctx->stream);
    end_shard_phase(ctx, false);
    PG_CUDA(cudaStreamSynchronize(ctx->copyStream));
    return 0;
}

int pg_set_split_memory_limit(pg_context *ctx, uint64_t bytes) {
    PG_CHECK(ctx, "pg_set_split_memory_limit: null argument");
    ctx->memLimit = bytes;
    return 0;
}

int pg_debug_no_scratch_alias(pg_context *ctx, int on) { if (!ctx) return 1; ctx->noScratchAlias = on != 0; return 0; }

// tests: run the kmermatcher stage in exactly n hash-range splits (0 = decide from the memory limit)
int pg_debug_force_splits(pg_context *ctx, unsigned n) { if (!ctx) return 1; ctx->forceSplits = n; return 0; }

// tests: force the full-sort group path (1) or allow the bucketed hash join (0)
int pg_debug_force_full_sort(pg_context *ctx, int on) { if (!ctx) return 1; ctx->forceFullSort = on != 0; return 0; }

// 256-bin radix pass: 0 register-tile kernel, 1 bulk-copy kernel (3072-record tiles, 2 stages), 2 bulk-copy kernel (2048, 3 stages)
int pg_debug_set_radix_mode(int mode) { radix_set_mode(mode); return 0; }
int pg_debug_get_radix_mode(void) { return radix_get_mode(); }

// radix digit width of the fast-path sorts: 8 (256-bin passes), 9 or 10 (wide-digit kernel, fewer passes)
int pg_debug_set_digit_bits(pg_context *ctx, int bits) {
    PG_CHECK(ctx && bits >= 8 && bits <= 10, "pg_debug_set_digit_bits: 8, 9 or 10");
    ctx->digitBits = bits;
    return 0;
}

// micro-benchmark of the radix sort on device-resident pseudo-random records: returns ms per scatter pass
int pg_debug_radix_bench_w(pg_context *ctx, uint64_t n, int items, int passes, int digit_bits, float *ms_per_pass);
int pg_debug_radix_bench(pg_context *ctx, uint64_t n, int items, int passes, float *ms_per_pass) {
    return pg_debug_radix_bench_w(ctx, n, items, passes, 8, ms_per_pass);
}
int pg_debug_radix_bench_w(pg_context *ctx, uint64_t n, int items, int passes, int digit_bits, float *ms_per_pass) {
    PG_CHECK(ctx && ms_per_pass, "pg_debug_radix_bench: null argument");
    PG_CHECK(digit_bits >= 8 && digit_bits <= 10, "pg_debug_radix_bench: digit bits must be 8, 9 or 10");
    cudaSetDevice(ctx->device);
    radix_set_items(items);
    RadixPlan plan; plan.npasses = 0;
    if (digit_bits > 8) plan_add_bits_w(plan, 0, 0, digit_bits * passes, digit_bits);
    else plan_add_bits(plan, 0, 0, 8 * passes);
    PG_TRY(ctx->recA.reserve(sizeof(Rec) * (n + 1)));
    PG_TRY(ctx->recB.reserve(sizeof(Rec) * (n + 1)));
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(n, digit_bits)));
    std::vector<unsigned long long> seed(1 << 20);
    unsigned long long x = 88172645463325252ull;
    for (auto &v : seed) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x; }
    for (uint64_t off = 0; off < n * 2; off += seed.size()) {
        const uint64_t c = std::min<uint64_t>(seed.size(), n * 2 - off);
        PG_CUDA(cudaMemcpyAsync((unsigned long long *) ctx->recA.p + off, seed.data(), c * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    Rec *sorted = nullptr; uint64_t launches = 0;
    for (int rep = 0; rep < 2; rep++)
        PG_TRY(radix_sort(ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), n, plan, ctx->radixWs.p, ctx->radixWs.cap, ctx->stream, &sorted, &launches,
                          ctx->ev[EV_SCATTER1_BEGIN], ctx->ev[EV_SCATTER1_END]));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    PG_CUDA(cudaEventElapsedTime(&ms, ctx->ev[EV_SCATTER1_BEGIN], ctx->ev[EV_SCATTER1_END]));
    *ms_per_pass = ms / passes;
    radix_set_items(12);
    return 0;
}

// ---- diagnostics used by the tests (not part of the drop-in surface) --------------------------------
int pg_debug_radix_sort(pg_context *ctx, uint64_t *recs /* n x 2 u64, in place */, uint64_t n, const int *word, const int *lo, const int *hi, int nRanges) {
    PG_CHECK(ctx && recs, "pg_debug_radix_sort: null argument");
    cudaSetDevice(ctx->device);
    RadixPlan plan; plan.npasses = 0;
    for (int i = 0; i < nRanges; i++) {
        if (ctx->digitBits > 8) plan_add_bits_w(plan, word[i], lo[i], hi[i], ctx->digitBits);
        else plan_add_bits(plan, word[i], lo[i], hi[i]);
    }
    PG_TRY(ctx->recA.reserve(sizeof(Rec) * (n + 1)));
    PG_TRY(ctx->recB.reserve(sizeof(Rec) * (n + 1)));
    PG_TRY(ctx->radixWs.reserve(radix_workspace_bytes(n, ctx->digitBits)));
    PG_CUDA(cudaMemcpyAsync(ctx->recA.p, recs, sizeof(Rec) * n, cudaMemcpyHostToDevice, ctx->stream));
    Rec *sorted = nullptr;
    uint64_t launches = 0;
    PG_TRY(radix_sort(ctx->recA.as<Rec>(), ctx->recB.as<Rec>(), n, plan, ctx->radixWs.p, ctx->radixWs.cap, ctx->stream, &sorted, &launches));
    PG_CUDA(cudaMemcpyAsync(recs, sorted, sizeof(Rec) * n, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
