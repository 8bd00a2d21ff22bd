// pg_internal.cuh -- library-internal state shared by the stage implementations.
#pragma once
#include "pg_common.cuh"
#include "pg_radix.cuh"

namespace pg {

enum {
    EV_KM_BEGIN = 0, EV_EXTRACT_END, EV_SORT1_BEGIN, EV_SORT1_END, EV_GROUP_END, EV_SORT2_END, EV_REDUCE_END,
    EV_SCATTER1_BEGIN, EV_SCATTER1_END, EV_RS_BEGIN, EV_RS_END, EV_EX_BEGIN, EV_EX_END, EV_TOTAL_BEGIN, EV_TOTAL_END, EV_COUNT
};

}  // namespace pg

// One per process / GPU.
struct pg_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;         // device -> host copies of finished stage results, overlapped with the next stage
    cudaStream_t h2dStream = nullptr;          // pg_seqdb_upload_async: the next input travels while the current iteration computes
    cudaEvent_t evCopyReady = nullptr;
    // asynchronous results (pg_set_async_results): the device -> host copies are only enqueued; tickets mark positions
    // of the copy stream.  evHitsCopied / evAlnsCopied guard the result buffers against the next call's writes.
    bool asyncResults = false;
    cudaEvent_t evHitsCopied = nullptr, evAlnsCopied = nullptr;
    cudaEvent_t evTicket[16];
    uint64_t nextTicket = 0;
    cudaStream_t auxStream = nullptr;          // independent kernel work that runs next to the main stream (extension: heap rounds)
    cudaEvent_t evAuxFork = nullptr, evAuxJoin = nullptr;
    unsigned long long *hostStage = nullptr;   // mapped pinned words: the kernels' small results are read back through here (pg::read_back)
    cudaEvent_t ev[pg::EV_COUNT];
    pg::DevBuf small, lists, recA, recB, radixWs, scratch, blockCounts, hits, alnAll, alns, flags, exWork, exSegs, exMeta, exLists, ntTab, buckets, buckets2, wideTabs, nextWork, orfInfo;
    unsigned bucketTarget = 512;  // average records per bucket of the partial-key partition (<= 700: small hash-join instance first)
    int digitBits = 8;            // radix digit width of the two fast-path sorts (8: 256 bins; 9 / 10: wide-digit kernel)
    bool forceFullSort = false;   // tests: take the 8-pass sort + group_kernel path instead of the bucketed hash join
    // --split-memory-limit (pg_set_split_memory_limit): bound on the two k-mer record buffers; splitDiv = number of equal
    // hash-range splits the running kmermatcher call uses (1 = no split), pairAcc collects the splits' pair records
    bool noScratchAlias = false;  // diagnostics: keep the later stages' scratch out of the record buffers
    uint64_t memLimit = 0;
    uint64_t deviceMemBytes = 0;  // total device memory (pg_init)
    uint64_t kmerTotalHint = 0;   // computeKmerCount + 1 of the whole DB, handed from km_choose_splits to km_extract
    unsigned splitDiv = 1;
    unsigned forceSplits = 0;     // tests: use exactly this many splits
    pg::DevBuf pairAcc, spill;
    unsigned ntTabN = 0;
    bool pairsInA = false;
    bool tExtract = false, tGroup = false, tReduce = false, rsRan = false, exRan = false;   // which stages recorded their events in this call
    pg_timings shardAcc;                       // multi-GPU: stage times accumulated over the phases of one step
    unsigned seqLo = 0, seqHi = 0xFFFFFFFFu;   // multi-GPU: extraction restricted to the sequences with index in [seqLo, seqHi)
    uint64_t launches = 0;
    pg_timings timings;
    uint64_t nHits = 0, nAlns = 0;
    // rescorediagonal leaves, per sequence index, the number of alignment lines and their first position in its output;
    // the extension of the same call reuses them instead of re-deriving the ranges from the alignment array
    const pg_aln *rsOut = nullptr;
    const unsigned *rsCnt = nullptr;
    const unsigned long long *rsOff = nullptr;
    // multi-GPU: this rank owns the queries / representatives with key in [ownLo, ownHi)
    unsigned ownLo = 0, ownHi = 0xFFFFFFFFu;
    pg::Rec *shardPairs = nullptr;
    uint64_t shardPairCount = 0;
    // multi-GPU data plane (pg_shard.cu): one process per GPU, NCCL communicator over NVLink / NVSwitch
    void *comm = nullptr;                      // ncclComm_t
    int rank = 0, world = 1;
    cudaEvent_t evXchg[4] = {nullptr, nullptr, nullptr, nullptr};   // begin / end of the two exchanges of a step
    pg::DevBuf commWs;                         // counts matrix, histogram, bounds
    unsigned long long firstKmerOverride = 0;  // nt: the job-wide smallest k-mer (all-reduced), see hash_group_kernel
    bool useFirstKmerOverride = false;
    // fused partition + exchange over peer memory (CUDA IPC): this rank's two receive buffers (k-mer records, pair records) and
    // every rank's mapping of them
    pg::DevBuf xr[2];
    void *peer[2][16] = {};
    bool p2pDisabled = false;
    bool extractOnly = false;                  // km_extract: do not reserve the sort's second record buffer
    // histograms of sort #1's partition digits, counted by the extraction kernels (3 x 256 bins); valid for the km_group call
    // that follows the extraction directly
    pg::DevBuf preHist;
    bool preHistValid = false;
    uint64_t preHistRecords = 0;
    bool noPreHist = false;                    // debugging / tests: always run the histogram sweep
    float lastExchangeMs[2] = {0, 0};
    uint64_t lastExchangeBytes[2] = {0, 0};
    uint32_t lastBounds[257];
};

namespace pg {
typedef pg_context Context;

struct KmConst;
// kmermatcher stages (pg_kmermatch.cu)
int km_run(Context *ctx, const pg_seqdb *db, const pg_km_params *p, pg_hit **d_hits, uint64_t *nHits);
int km_shard_pairs(Context *ctx, const pg_seqdb *db, const pg_km_params *p, int world, uint64_t *counts);
int km_shard_extract(Context *ctx, const pg_seqdb *db, const pg_km_params *p, int rank, int world, uint64_t *counts);
int km_shard_group(Context *ctx, const pg_seqdb *db, const pg_km_params *p, const void *d_records, uint64_t nRecords, uint64_t *hist);
int km_shard_extract_only(Context *ctx, const pg_seqdb *db, const pg_km_params *p, int rank, int world, uint64_t *nRecords);
void km_shard_pairs_location(Context *ctx, Rec **pairs, uint64_t *n);
int km_shard_route(Context *ctx, int world, const unsigned *bounds, uint64_t *counts);
void km_equal_key_bounds(unsigned max_key, int world, unsigned *bounds);
int km_shard_reduce(Context *ctx, const pg_seqdb *db, const void *d_pairs, uint64_t nPairs, pg_hit **d_hits, uint64_t *nHits);
// rescorediagonal (pg_rescore.cu): d_hits sorted by (rep,target); result device array in ctx->alns
int rs_run(Context *ctx, const pg_seqdb *db, const pg_hit *d_hits, uint64_t nHits, const pg_rs_params *p, pg_aln **d_alns, uint64_t *nAlns);
// extension (pg_extend.cu): d_alns sorted by query; produces a new device DB
int ex_run(Context *ctx, const pg_seqdb *db, const pg_aln *d_alns, uint64_t nAlns, const pg_ex_params *p, pg_seqdb **out, unsigned char **d_extended);
// findassemblystart / cyclecheck (pg_next.cu); the returned device arrays live in ctx->nextWork until the next call
int fs_run(Context *ctx, const pg_seqdb *db, const pg_aln *d_alns, uint64_t nAlns, pg_seqdb **out, int **d_addStop);
int cc_run(Context *ctx, const pg_seqdb *db, int maxSeqLen, int k, unsigned **d_split);
// extractorfs (+ translatenucs --add-orf-stop 1) (pg_orf.cu); d_info (4 words per fragment) lives in ctx->orfInfo
int orf_run(Context *ctx, const pg_seqdb *db, const pg_orf_params *p, int translate, pg_seqdb **out, unsigned **d_info);
int tn_run(Context *ctx, const pg_seqdb *db, const unsigned char *d_flags, int translationTable, pg_seqdb **out);
int db_ready(Context *ctx, const pg_seqdb *db);   // completes a pg_seqdb_upload_async at the DB's first use
int seqdb_finalize(Context *ctx, pg_seqdb *db);   // computes max_seq_len / residues / dense_keys on the device
void seqdb_release(pg_seqdb *db, cudaStream_t s);
// Small device -> host read-back (counters, totals) that does NOT go through the copy engine: a one-warp kernel stores
// the words into mapped pinned memory and the stream is synchronised.  A cudaMemcpyAsync would queue behind the large
// result transfers that run on ctx->copyStream underneath the following stage.  bytes <= 1024, multiple of 4.
int read_back(Context *ctx, void *host, const void *dev, size_t bytes);
int read_back_on(Context *ctx, cudaStream_t stream, void *host, const void *dev, size_t bytes);   // same on another stream of the context
int alloc_pinned(size_t bytes, void **out);        // pooled pinned host memory, released with pg_free_host
// call bracketing shared by pg_api.cu and pg_shard.cu
void begin_call(Context *ctx);
void end_call(Context *ctx);
void end_shard_phase(Context *ctx, bool first);
int hits_to_host_overlapped(Context *ctx, const pg_hit *d, uint64_t n, pg_hit **out);
int alns_to_host_overlapped(Context *ctx, const pg_aln *d, uint64_t n, pg_aln **out);
void km_min_kmer_slot(Context *ctx, unsigned long long **d_min);
}  // namespace pg
