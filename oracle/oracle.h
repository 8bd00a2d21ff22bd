/* oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a scalar, single-threaded restatement of the reference's
 * assemble-iteration hot path (kmermatcher -> rescorediagonal -> assembleresults /
 * nuclassembleresults) used to check the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it; the product (plass_b200/) never does.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py checks every function here against
 * outputs of the unmodified reference binary (oracle/_ref, built by oracle/ref_build.mk) committed
 * as fixtures under tests/golden/ (generator: tests/golden/make_golden.py).
 *
 * All citations are file:line in /root/reference ("mm/" = lib/mmseqs/src/).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A sequence DB as the reference's DBReader exposes it (mm/commons/DBReader.h:151-236):
 * entry i = data[offsets[i] .. offsets[i]+lens[i]) = residues + '\n' + '\0', so seqLen = lens[i]-2.
 * keys are the DB keys; entries are in index order (key-sorted for sequence DBs). */
typedef struct {
    const char *data;
    const uint64_t *offsets;
    const uint32_t *lens;
    const uint32_t *keys;
    uint64_t n;
    int dbtype; /* 0 = amino acids, 1 = nucleotides (mm/commons/Parameters.h:65-79) */
} or_seqdb;

/* kmermatcher parameters actually read on this path (mm/commons/Parameters.cpp:871-892). */
typedef struct {
    int kmer_size;              /* -k */
    int alph_size;              /* --alph-size: 13 (reduced) or 21 for aa; ignored for nt */
    int kmers_per_seq;          /* --kmer-per-seq */
    float kmers_per_seq_scale;  /* --kmer-per-seq-scale (aa or nt value as applicable) */
    int hash_shift;             /* --hash-shift (XXH64 seed) */
    int include_only_extendable;/* --include-only-extendable */
    int ignore_multi_kmer;      /* --ignore-multi-kmer */
    int cov_mode;               /* --cov-mode */
    float cov_thr;              /* -c */
    uint64_t hash_start;        /* split hash range, 0 / UINT64_MAX when unsplit (kmermatcher.cpp:736-778) */
    uint64_t hash_end;
} or_km_params;

/* One k-mer record = KmerPosition<T> (mm/linclust/kmermatcher.h:49-54). */
typedef struct {
    uint64_t kmer;
    uint32_t id;
    int32_t seq_len;
    int32_t pos;
} or_kmer_rec;

/* One prefilter hit line "target \t score \t diag" inside the block of `rep`
 * (mm/prefiltering/QueryMatcher.h:35-51,114-126).  score < 0 encodes reverse strand (nt). */
typedef struct {
    uint32_t rep;
    uint32_t target;
    int32_t score;
    int32_t diag; /* as printed: (short) of the 16-bit diagonal */
} or_hit;

/* rescorediagonal parameters (mm/commons/Parameters.cpp:422-439). */
typedef struct {
    int rescore_mode;   /* only 3 = END_TO_END is implemented */
    float seq_id_thr;   /* --min-seq-id */
    double eval_thr;    /* -e */
    int cov_mode;
    float cov_thr;
    int aln_len_thr;    /* --min-aln-len */
    int seq_id_mode;    /* --seq-id-mode */
} or_rs_params;

/* One accepted alignment = Matcher::result_t fields that are printed (mm/alignment/Matcher.h:32-91). */
typedef struct {
    uint32_t query;
    uint32_t target;
    int32_t bits;
    float seq_id;
    double evalue;
    int32_t q_start, q_end, q_len;
    int32_t db_start, db_end, db_len;
} or_aln;

typedef struct {
    float seq_id_thr;   /* --min-seq-id */
    int max_seq_len;    /* --max-seq-len */
    int keep_target;    /* --keep-target */
    int rescore_mode;   /* 3 */
} or_ex_params;

/* XXH64 of one little-endian u64 (kmermatcher.cpp:33-38; xxhash.h XXH64 with len 8). */
uint64_t or_hash_u64(uint64_t v, uint64_t seed);

/* fillKmerPositionArray (kmermatcher.cpp:77-385): returns malloc'd records in per-sequence
 * emission order (sequence-hash record first, then selected k-mers). */
int or_extract_kmers(const or_seqdb *db, const or_km_params *p, or_kmer_rec **out, uint64_t *n_out);

/* Full kmermatcher (kmermatcher.cpp:387-924, single split): hits in output order
 * (rep ascending, target ascending).  is_rep[key] = repSequence bitmap (size last key + 1). */
int or_kmermatch(const or_seqdb *db, const or_km_params *p, or_hit **out, uint64_t *n_out);

/* rescorediagonal over "self line + hits" for every key (rescorediagonal.cpp:45-379), query == target DB. */
int or_rescore(const or_seqdb *db, const or_hit *hits, uint64_t n_hits, const or_rs_params *p,
               or_aln **out, uint64_t *n_out);

/* assembleresults / nuclassembleresults (src/assembler/assembleresult.cpp:110-356,
 * nuclassembleresult.cpp:144-398).  Output: a sequence DB in key order (out_n entries; fewer than
 * db->n only when keep_target == 0); extended[i] = 1 if entry i is a new contig.  out_data entries
 * are residues + '\n' + '\0'. */
int or_extend(const or_seqdb *db, const or_aln *alns, uint64_t n_alns, const or_ex_params *p,
              char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
              uint8_t **extended, uint64_t *out_n, uint64_t *out_bytes);

/* Text formatting, byte-identical to the reference writers. Return bytes written (without NUL). */
int or_format_hit(char *buf, uint32_t target, int32_t score, int32_t diag);      /* QueryMatcher.h:114-126 */
int or_format_aln(char *buf, const or_aln *a);                                    /* Matcher.cpp:323-370 */

/* E-value helpers (mm/alignment/EvalueComputation.h:18-40 + ALP area). nt != 0 selects nucleotide.out. */
double or_evalue(int nt, double db_residues, double score, double q_len);
double or_bitscore(int nt, double score);
double or_raw_from_bits(int nt, double bits);

/* ---- components ranked "next" in SURVEY.md section 8(f) (oracle_next.cpp) ---- */

/* findassemblystart (src/assembler/findassemblystart.cpp:35-176): alns = alignment DB of db against itself, ordered by
 * query.  Output DB in key order; add_stop[i] = position the new '*' precedes in sequence i, or -1. */
int or_findstart(const or_seqdb *db, const or_aln *alns, uint64_t n_alns,
                 char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                 uint64_t *out_n, uint64_t *out_bytes, int32_t **add_stop);

/* extractorfs flags that reach the ORF finder (mm/commons/Parameters.cpp, extractorfs parameter list). */
typedef struct {
    int min_length;            /* --min-length (codons) */
    int max_length;            /* --max-length */
    int max_gaps;              /* --max-gaps */
    int contig_start_mode;     /* --contig-start-mode 0 / 1 / 2 */
    int contig_end_mode;       /* --contig-end-mode */
    int orf_start_mode;        /* --orf-start-mode 0 START_TO_STOP / 1 ANY_TO_STOP / 2 LAST_START_TO_STOP */
    unsigned forward_frames;   /* bit mask: 1 | 2 | 4 for frames 1,2,3 (Orf::getFrames) */
    unsigned reverse_frames;
    int translation_table;     /* only 1 */
    int use_all_table_starts;  /* --use-all-table-starts */
} or_orf_params;

/* extractorfs (mm/util/extractorfs.cpp:20-159), translate != 0: fused with translatenucs --add-orf-stop 1
 * (mm/util/translatenucs.cpp:14-128).  Fragments keyed 0..n-1 in (read, emission) order; orf_info = n x {read key,
 * fromPos, toPos, incompleteStart | incompleteEnd << 1} = the ORF header DB. */
int or_extractorfs(const or_seqdb *db, const or_orf_params *p, int translate,
                   char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                   uint64_t *out_n, uint64_t *out_bytes, uint32_t **orf_info);

/* translatenucs (mm/util/translatenucs.cpp:14-128) on any nucleotide DB; flags NULL = --add-orf-stop 0, else per sequence
 * bit 0 = '*' in front, bit 1 = '*' behind.  Keys kept, entries shorter than a codon dropped. */
int or_translatenucs(const or_seqdb *db, const uint8_t *flags, int max_seq_len,
                     char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                     uint64_t *out_n, uint64_t *out_bytes);

/* cyclecheck (src/assembler/cyclecheck.cpp:71-274): split[i] (caller-allocated, db->n) = splitDiagonal or 0. */
int or_cyclecheck(const or_seqdb *db, int max_seq_len, int kmer_size, uint32_t *split);

void or_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
