/* plassgpu.h -- C ABI of the B200-native assemble-iteration hot path (libplassgpu.so).
 *
 * Drop-in boundary: these entry points are what the reference's four hot-path commands bind to when
 * the CPU implementations are swapped out (see INTEGRATION.md for the Command-table stub):
 *
 *   pg_kmermatch   replaces  kmermatcherInner<T>           lib/mmseqs/src/linclust/kmermatcher.cpp:589-733
 *                            (fillKmerPositionArray :77-385, SORT_PARALLEL :408-412,427-431,
 *                             assignGroup :450-559, writeKmerMatcherResult :809-924)
 *   pg_rescore     replaces  doRescorediagonal             lib/mmseqs/src/alignment/rescorediagonal.cpp:45-379
 *   pg_extend      replaces  doassembly / doNuclAssembly   src/assembler/assembleresult.cpp:110-356,
 *                                                          src/assembler/nuclassembleresult.cpp:144-398
 *   pg_seqdb_*     replaces  DBReader<unsigned int>::getData/getSeqLen/getDbKey views
 *                                                          lib/mmseqs/src/commons/DBReader.h:151-236
 *
 * Plain C: pointers + sizes, no C++ or torch types.  All functions return 0 on success, non-zero on
 * error (pg_last_error() gives the message).  There is no CPU fallback: without a CUDA device every
 * compute entry point fails.  Arrays returned through `**out` parameters are pinned host memory owned
 * by the caller, released with pg_free_host().
 */
#ifndef PLASSGPU_H
#define PLASSGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_DBTYPE_AMINO_ACIDS 0   /* Parameters::DBTYPE_AMINO_ACIDS  (lib/mmseqs/src/commons/Parameters.h:65-79) */
#define PG_DBTYPE_NUCLEOTIDES 1   /* Parameters::DBTYPE_NUCLEOTIDES */

typedef struct pg_context pg_context; /* one per process/GPU: device, stream, workspace, (optional) NCCL comm */
typedef struct pg_seqdb pg_seqdb;     /* a sequence DB resident in HBM */

/* Host view of a sequence DB exactly as DBReader exposes it: entry i = data[offsets[i] .. +lens[i]) =
 * residues + '\n' + '\0' (getSeqLen = lens[i]-2, DBReader.h:192-213); keys ascending (index order). */
typedef struct {
    const char *data;
    uint64_t data_bytes;
    const uint64_t *offsets;
    const uint32_t *lens;
    const uint32_t *keys;
    uint64_t n;
    int dbtype;
} pg_seqdb_view;

/* kmermatcher flags that reach the kernels (Parameters.cpp:871-892).  Unsupported settings
 * (--mask 1, --spaced-kmer-mode 1, --adjust-kmer-len 1, custom --sub-mat) are rejected by the caller. */
typedef struct {
    int kmer_size;               /* -k */
    int alph_size;               /* --alph-size (aa: 13 reduced / 21 full; nt: 5) */
    int kmers_per_seq;           /* --kmer-per-seq */
    float kmers_per_seq_scale;   /* --kmer-per-seq-scale */
    int hash_shift;              /* --hash-shift */
    int include_only_extendable; /* --include-only-extendable */
    int ignore_multi_kmer;       /* --ignore-multi-kmer */
    int cov_mode;                /* --cov-mode */
    float cov_thr;               /* -c */
    uint32_t hash_start;         /* inclusive 16-bit hash range of this shard (setupKmerSplits, kmermatcher.cpp:736-778) */
    uint32_t hash_end;           /* 0..65535 = everything */
} pg_km_params;

/* One prefilter hit line of the block of `rep`: "target \t score \t diag" (QueryMatcher.h:35-51,114-126).
 * score < 0 <=> reverse strand (nt, DBTYPE_PREFILTER_REV_RES). Ordered by (rep, target). */
typedef struct {
    uint32_t rep;
    uint32_t target;
    int32_t score;
    int32_t diag; /* (short) of the 16-bit diagonal, as printed */
} pg_hit;

typedef struct {          /* Parameters.cpp:422-439 */
    int rescore_mode;     /* --rescore-mode, only 3 (END_TO_END) */
    float seq_id_thr;     /* --min-seq-id */
    double eval_thr;      /* -e */
    int cov_mode;         /* --cov-mode */
    float cov_thr;        /* -c */
    int aln_len_thr;      /* --min-aln-len */
    int seq_id_mode;      /* --seq-id-mode */
} pg_rs_params;

/* One accepted alignment = the printed fields of Matcher::result_t (Matcher.h:32-91, Matcher.cpp:323-370).
 * Ordered by (query, prefilter order); the self alignment of every query comes first. */
typedef struct {
    uint32_t query;
    uint32_t target;
    int32_t bits;
    float seq_id;
    double evalue;
    int32_t q_start, q_end, q_len;
    int32_t db_start, db_end, db_len;
} pg_aln;

typedef struct {          /* src/commons/LocalParameters.h:96-102 */
    float seq_id_thr;     /* --min-seq-id */
    int max_seq_len;      /* --max-seq-len */
    int keep_target;      /* --keep-target */
    int rescore_mode;     /* --rescore-mode, only 3 */
} pg_ex_params;

/* Device time (CUDA events on the library's stream) of the stages of the last call, in ms. */
typedef struct {
    float extract_ms, sort1_ms, group_ms, sort2_ms, reduce_ms;   /* kmermatcher */
    float rescore_ms, extend_ms, exchange_ms, total_ms;
    uint64_t n_kmer_records;   /* records entering sort #1 (Sum m) */
    uint64_t n_pair_records;   /* records entering sort #2 */
    uint64_t n_hits;           /* prefilter hit lines (h) */
    uint64_t n_alns;           /* accepted alignments */
    uint64_t n_extended;       /* sequences that became contigs */
    uint64_t kernel_launches;  /* kernels launched by the last call */
    uint64_t sort1_bytes;      /* algorithmic bytes of sort #1: one read + one write of every record */
    float sort1_scatter_ms;    /* device time of the radix scatter launches of sort #1 (events around them) */
    uint32_t sort1_passes;     /* number of scatter launches in sort #1 */
    uint32_t splits;           /* hash-range splits of the kmermatcher stage (1 = everything at once) */
    uint64_t spilled_records;  /* k-mer records of buckets beyond the shared-memory hash join (sorted and grouped apart) */
} pg_timings;

const char *pg_last_error(void);
int pg_device_count(void);
int pg_init(int device, pg_context **ctx);
void pg_destroy(pg_context *ctx);
int pg_get_timings(const pg_context *ctx, pg_timings *out);
/* frees the context's cached device buffers (re-created on demand) */
int pg_release_workspace(pg_context *ctx);

int pg_seqdb_upload(pg_context *ctx, const pg_seqdb_view *view, pg_seqdb **db);
/* Same, but only ENQUEUES the copies on the context's upload stream and returns: the transfer of the next input runs
 * underneath the kernels of the current call.  `view`'s arrays must be pinned host memory and stay alive until the DB
 * has been used by (or waited for in) a compute call, which orders itself after the copies. */
int pg_seqdb_upload_async(pg_context *ctx, const pg_seqdb_view *view, pg_seqdb **db);
/* Same, but the view's four arrays are DEVICE pointers owned by the caller (not copied; pg_seqdb_free releases only
 * the handle; data needs 16 readable bytes past data_bytes).  Multi-GPU: every rank copies its slice of the DB to
 * its GPU, the slices are all-gathered over NVLink, and the gathered arrays are adopted (plass_b200/sharded.py). */
int pg_seqdb_adopt(pg_context *ctx, const pg_seqdb_view *device_view, pg_seqdb **db);
/* Copies a device-resident DB back: all four arrays are pinned host buffers (pg_free_host). */
int pg_seqdb_download(pg_context *ctx, const pg_seqdb *db, char **data, uint64_t *data_bytes, uint64_t **offsets,
                      uint32_t **lens, uint32_t **keys, uint64_t *n);
uint64_t pg_seqdb_size(const pg_seqdb *db);
void pg_seqdb_free(pg_context *ctx, pg_seqdb *db);

/* The three steps with HOST inputs/outputs (what the command shims call). */
int pg_kmermatch(pg_context *ctx, const pg_seqdb *db, const pg_km_params *p, pg_hit **hits, uint64_t *n_hits);
int pg_rescore(pg_context *ctx, const pg_seqdb *db, const pg_hit *hits, uint64_t n_hits, const pg_rs_params *p,
               pg_aln **alns, uint64_t *n_alns);
int pg_extend(pg_context *ctx, const pg_seqdb *db, const pg_aln *alns, uint64_t n_alns, const pg_ex_params *p,
              pg_seqdb **out_db, uint8_t **extended);

/* --split-memory-limit (kmermatcherInner, kmermatcher.cpp:608-624: splits = ceil(totalKmers * 16 B / limit);
 * setupKmerSplits :736-778; merge :926-1104).  `bytes` bounds the two k-mer record buffers of the kmermatcher stage
 * (0 = 90 % of the device memory that is free when the stage starts).  If the records do not fit, the 16-bit hash space
 * is cut into 2, 4, 8 ... equal ranges (pg_km_params.hash_start / hash_end per split); every split extracts, sorts and
 * groups only its k-mers, the (rep, target, diagonal) pair records of all splits are collected and reduced TOGETHER, so
 * the hits equal the unsplit run's (the reference's own merge pre-aggregates with a saturating 8-bit score and is not
 * bit-identical to its unsplit path).  pg_timings.splits reports the number used. */
int pg_set_split_memory_limit(pg_context *ctx, uint64_t bytes);

/* One whole assemble iteration kept in HBM: kmermatcher -> rescorediagonal -> (nucl)assembleresults.
 * `out_db` is the next iteration's input.  If hits/alns pointers are non-NULL the intermediate
 * results are also copied to pinned host arrays (for writing pref_N / aln_N). */
int pg_assemble_iteration(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp,
                          const pg_ex_params *ep, pg_seqdb **out_db,
                          pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns);

/* The two per-iteration helpers next to the hot path.
 *
 * pg_findassemblystart  replaces findassemblystart, src/assembler/findassemblystart.cpp:35-176 (plass STEP 0 only,
 *                       data/assemble.sh:108-117): `alns` = the alignment DB of `db` against itself, ordered by query.
 *                       For every query with an 'M': if >= 20 % of {query, aligned targets} carry "*M" at the projected
 *                       position, each member's sequence is cut to "*" + residues[M position ..]; out_db holds every
 *                       sequence (changed or not).  add_stop (optional, pinned, one int32 per sequence) = the cut
 *                       position or -1.
 * pg_cyclecheck         replaces cyclecheck, src/assembler/cyclecheck.cpp:71-274 (every penguin iteration,
 *                       data/nuclassemble.sh:19-60), k = 22: split[i] (pinned, one uint32 per sequence) = the diagonal
 *                       at which sequence i repeats itself (circular / terminally redundant), 0 = not reported.  The
 *                       caller writes the reported sequences (--chop-cycle: the first split[i] residues). */
int pg_findassemblystart(pg_context *ctx, const pg_seqdb *db, const pg_aln *alns, uint64_t n_alns, pg_seqdb **out_db, int32_t **add_stop);
/* plass STEP 0 fused (data/assemble.sh:88-151 with STEP = 0): kmermatcher -> rescorediagonal -> findassemblystart ->
 * kmermatcher -> rescorediagonal -> assembleresults without leaving HBM.  corrected_db (optional) = corrected_seqs,
 * out_db = assembly_0; hits / alns (optional, pinned) = pref_corrected_0 / aln_corrected_0. */
int pg_assemble_step0(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp, const pg_ex_params *ep,
                      pg_seqdb **corrected_db, pg_seqdb **out_db, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns);
int pg_cyclecheck(pg_context *ctx, const pg_seqdb *db, int max_seq_len, uint32_t **split);

/* Six-frame ORF extraction of the reads, the step upstream of the first iteration.
 *
 * pg_extractorfs   replaces extractorfs, lib/mmseqs/src/util/extractorfs.cpp:20-159 (Orf::findForward,
 *                  lib/mmseqs/src/commons/Orf.cpp:192-330) and, with translate != 0, the translatenucs --add-orf-stop 1 that
 *                  follows it in data/assemble.sh:41-66 (lib/mmseqs/src/util/translatenucs.cpp:14-128): out_db holds the
 *                  fragments (nucleotides, or amino acids framed by '*' where the ORF has a real start / stop) with keys
 *                  0..n-1 in (read, emission) order, i.e. the renumbered DB the reference writes; orf_info (optional,
 *                  pinned, 4 words per fragment) = {read key, fromPos, toPos, incompleteStart | incompleteEnd << 1}, the
 *                  fields of the ORF header DB (Orf::writeOrfHeader, Orf.cpp:445-462).
 * pg_seqdb_concat  replaces concatdbs without --preserve-keys on two sequence DBs (data/assemble.sh:68-77): a's entries
 *                  keep the keys 0..a.n-1, b's follow. */
typedef struct {               /* extractorfs flags (Parameters.cpp, extractorfs list) */
    int min_length;            /* --min-length (codons) */
    int max_length;            /* --max-length */
    int max_gaps;              /* --max-gaps */
    int contig_start_mode;     /* --contig-start-mode */
    int contig_end_mode;       /* --contig-end-mode */
    int orf_start_mode;        /* --orf-start-mode */
    unsigned forward_frames;   /* bit mask 1 | 2 | 4 = frames 1,2,3 (Orf::getFrames, Orf.h:17-35) */
    unsigned reverse_frames;
    int translation_table;     /* --translation-table, only 1 */
    int use_all_table_starts;  /* --use-all-table-starts */
} pg_orf_params;
int pg_extractorfs(pg_context *ctx, const pg_seqdb *db, const pg_orf_params *p, int translate, pg_seqdb **out_db, uint32_t **orf_info);
int pg_seqdb_concat(pg_context *ctx, const pg_seqdb *a, const pg_seqdb *b, pg_seqdb **out_db);
/* pg_translatenucs replaces translatenucs (lib/mmseqs/src/util/translatenucs.cpp:14-128) on any nucleotide DB: flags
 * (optional, host, one byte per sequence) bit 0 = '*' in front, bit 1 = '*' behind unless the last residue is one
 * (--add-orf-stop 1: from the ORF header of the sequence); entries of fewer than three bytes are dropped, keys are kept. */
int pg_translatenucs(pg_context *ctx, const pg_seqdb *db, const uint8_t *flags, int translation_table, pg_seqdb **out_db);

/* Asynchronous result transfer.  The reference workflow writes pref_N / aln_N / assembly_N to disk while the next
 * iteration's input already sits in HBM; with pg_set_async_results(ctx, 1) pg_assemble_iteration (hits, alns) and
 * pg_seqdb_download only ENQUEUE their device->host copies on the context's copy stream and return, so that the
 * transfers of iteration N run underneath iteration N+1.  The returned host arrays may be read after
 * pg_results_wait(ctx, t) for a ticket t taken (pg_results_ticket) after the calls that produced them.  A DB whose
 * download is pending may be released with pg_seqdb_free right away (the release is ordered after the copy).
 * pg_set_async_results(ctx, 0) drains the copy stream and restores the blocking behaviour. */
int pg_set_async_results(pg_context *ctx, int on);
int pg_results_ticket(pg_context *ctx, uint64_t *ticket);
int pg_results_wait(pg_context *ctx, uint64_t ticket);

/* Multi-GPU, one process per GPU.  The sequence DB is replicated in every HBM; one iteration is
 * split around its two exchange steps, which the caller performs on the device buffers (NCCL all-to-all through
 * torch.distributed in bench.py / plass_b200/sharded.py):
 *
 *   pg_shard_extract  k-mer extraction (fillKmerPositionArray, kmermatcher.cpp:77-385) of this rank's slice of the
 *                     SEQUENCES; the 16-byte k-mer records are left on the device partitioned by the rank that owns
 *                     the k-mer (a fixed hash of the k-mer; equal k-mers meet on one rank, which is all that sort #1
 *                     + assignGroup need).  counts[world] (host) = records per destination rank.
 *   -- all-to-all #1 (k-mer records) --
 *   pg_shard_group    received records -> sort #1 + assignGroup (kmermatcher.cpp:408-559); the (rep, target,
 *                     diagonal) pair records stay on the device; rep_hist[PG_SHARD_HIST_BINS] (host) = number of
 *                     pair records per slice of the representative key space, bin = rep * BINS / (max_key + 1).
 *   pg_shard_route    representatives are owned in contiguous key ranges, rank r owning [bounds[r], bounds[r+1])
 *                     (bounds[0] = 0, bounds[world] = 0xFFFFFFFF).  Representatives are the longest / lowest-id
 *                     members of their groups, so the work per key is heavily skewed towards low ids: the caller
 *                     sums rep_hist over the ranks and cuts the key space into ranges of equal work
 *                     (plass_b200/sharded.py:balanced_bounds).  The pair records are partitioned by owner;
 *                     counts[world] as above.
 *   -- all-to-all #2 (pair records) --
 *   pg_shard_finish   received pairs -> sort #2 + best diagonal (kmermatcher.cpp:427-431, :809-924) ->
 *                     rescorediagonal -> (nucl)assembleresults for the queries with key in [own_lo, own_hi);
 *                     out_db holds only those sequences.  hits / alns (optional) are copied to pinned host arrays.
 *   pg_shard_export   copies the records the preceding phase left behind into a caller-supplied DEVICE buffer
 *                     (the all-to-all send buffer).
 *
 *   pg_shard_pairs    single-exchange variant = the reference's memory-split mechanism (kmermatcher.cpp:736-778):
 *                     every rank extracts from ALL sequences only the k-mers whose 16-bit hash lies in
 *                     [kp->hash_start, kp->hash_end], groups them and leaves the pair records partitioned by
 *                     representative owner (equal key ranges, pg_shard_owner_range); followed by all-to-all #2 and pg_shard_finish.  Extraction work is
 *                     replicated world times, so it is the better choice only for world <= 2.
 *
 * pg_get_timings after pg_shard_finish reports the sums over the phases of the step.
 */
int pg_shard_extract(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, int rank, int world, uint64_t *counts);
#define PG_SHARD_HIST_BINS 4096
int pg_shard_group(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const void *device_records, uint64_t n_records,
                   uint64_t *rep_hist);
int pg_shard_route(pg_context *ctx, int world, const uint32_t *bounds, uint64_t *counts);
int pg_shard_pairs(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, int world, uint64_t *counts);
int pg_shard_export(pg_context *ctx, void *device_dst, uint64_t n_records);
int pg_shard_finish(pg_context *ctx, const pg_seqdb *db, const void *device_pairs, uint64_t n_pairs,
                    uint32_t own_lo, uint32_t own_hi, const pg_rs_params *rp, const pg_ex_params *ep,
                    pg_seqdb **out_db, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns);
/* The same decomposition driven entirely from C++ over NCCL (one process per GPU; plass_b200/csrc/pg_shard.cu):
 *
 *   pg_comm_unique_id / pg_comm_init   ncclGetUniqueId on one rank, the PG_COMM_ID_BYTES bytes reach the others by any means
 *                                      (a file, MPI, torch.distributed), then ncclCommInitRank on every rank's context.
 *   pg_shard_broadcast_db              replicates a DB held by `root` in every rank's HBM over NVLink.
 *   pg_shard_iteration                 extraction of this rank's slice -> exchange #1 (k-mer records to the k-mer owner)
 *                                      -> sort #1 + assignGroup -> all-reduced work histogram -> equal-work key ranges ->
 *                                      exchange #2 (candidate pairs to the owner of the representative) -> sort #2 + best
 *                                      diagonal -> rescorediagonal -> extension of the owned queries.  Each exchange is
 *                                      FUSED into the partition pass in front of it: the pass stores every digit run
 *                                      straight into the owner's receive buffer over NVLink (CUDA IPC peer mappings,
 *                                      stream-ordered barrier); where peer memory cannot be mapped, or with
 *                                      PLASS_B200_SHARD_P2P=0, the partition pass is followed by grouped ncclSend / ncclRecv.
 *                                      The only host round trips are the W x W count matrices and the work histogram.
 *                                      out_slice = the new entries of the keys in [*own_lo, *own_hi).
 *   pg_shard_allgather_db              concatenates the ranks' slices (ascending key ranges in rank order) into a replicated
 *                                      DB: the next iteration's input (data/assemble.sh:153), or the upload of a host DB of
 *                                      which every rank copied only its slice over PCIe. */
#define PG_COMM_ID_BYTES 128
int pg_comm_unique_id(void *id);
int pg_comm_init(pg_context *ctx, int rank, int world, const void *id);
int pg_comm_destroy(pg_context *ctx);
int pg_comm_rank(const pg_context *ctx);
int pg_comm_world(const pg_context *ctx);
int pg_shard_broadcast_db(pg_context *ctx, const pg_seqdb *db_on_root, int root, pg_seqdb **out_db);
int pg_shard_allgather_db(pg_context *ctx, const pg_seqdb *slice, pg_seqdb **out_db);
int pg_shard_iteration(pg_context *ctx, const pg_seqdb *db, const pg_km_params *kp, const pg_rs_params *rp, const pg_ex_params *ep,
                       pg_seqdb **out_slice, uint32_t *own_lo, uint32_t *own_hi, pg_hit **hits, uint64_t *n_hits, pg_aln **alns, uint64_t *n_alns);
int pg_shard_exchange_stats(const pg_context *ctx, float *ms /* [2] */, uint64_t *bytes /* [2] */);
/* collective: releases the peer-mapped receive buffers of the fused exchanges (re-created by the next pg_shard_iteration) */
int pg_shard_release_buffers(pg_context *ctx);
/* the equal-work cut of the representative key space pg_shard_iteration uses (host arithmetic only) */
int pg_shard_balanced_bounds(const uint64_t *hist, int bins, uint32_t max_key, int world, double per_key_weight, uint32_t *bounds /* world + 1 */);

/* Key range [lo, hi) owned by rank r of `world` for a DB whose largest key is max_key. */
void pg_shard_owner_range(uint32_t max_key, int rank, int world, uint32_t *lo, uint32_t *hi);
uint32_t pg_seqdb_max_key(const pg_seqdb *db);

void pg_free_host(void *p);

#ifdef __cplusplus
}
#endif
#endif
