/* lambda_b200 -- C ABI of the B200-native seed-and-extend engine.
 *
 * The reference (seqan/lambda, lambda3) has no plugin/FFI boundary: it is one template-instantiated
 * binary.  The seam this library replaces is the loop body of realMain()
 * (reference src/search.cpp:428-457):
 *
 *      search(localHolder);            // src/search_algo.hpp:607   seeding + locate + pre-scoring
 *      iterateMatches(localHolder);    // src/search_algo.hpp:1365  widen/merge, DP pass 1, e-value
 *                                      //                           filter, DP pass 2 + traceback, stats
 *      iterativeSearchPre/Post(...);   // src/search_algo.hpp:1391/1409  phase-1 / phase-2 rule
 *      writeRecords(localHolder);      // src/search_algo.hpp:1335  per-query sort/unique/top-N
 *
 * in  = a batch of original-alphabet query sequences + the loaded index + options
 * out = the reference's lH.blastMatches (one lgpu_hit per BlastMatch) + the StatsHolder counters.
 *
 * Everything crossing the boundary is plain C: pointers, sizes, POD structs.  No C++ types, no
 * torch types, no exceptions.  All functions return 0 on success or a negative lgpu_status; the
 * message for the last failure is available from lgpu_last_error().
 *
 * There is NO CPU fallback behind this interface: every compute entry point runs CUDA kernels on
 * the device the index/context was created on and fails with LGPU_ERR_CUDA if that is impossible.
 */
#ifndef LAMBDA_B200_H
#define LAMBDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGPU_VERSION 100 /* 0.1.0 */

typedef enum
{
    LGPU_OK              = 0,
    LGPU_ERR_ARG         = -1, /* bad argument / unsupported option combination                  */
    LGPU_ERR_IO          = -2, /* file missing / truncated / wrong generation                    */
    LGPU_ERR_CUDA        = -3, /* no device, kernel failure, out of device memory                */
    LGPU_ERR_UNSUPPORTED = -4, /* index flavour or option outside the implemented hot-path scope */
    LGPU_ERR_INTERNAL    = -5
} lgpu_status;

/* AlphabetEnum / DbIndexType values exactly as stored in the .lba header
 * (reference src/shared_definitions.hpp:60-64, 127-136). */
enum { LGPU_ALPH_UNDEFINED = 0, LGPU_ALPH_DNA3BS = 1, LGPU_ALPH_DNA4 = 2, LGPU_ALPH_DNA5 = 3,
       LGPU_ALPH_AMINO_ACID = 4, LGPU_ALPH_MURPHY10 = 5, LGPU_ALPH_LI10 = 6 };
enum { LGPU_INDEX_FM = 0, LGPU_INDEX_BIFM = 1 };

/* search domain = which lambda3 sub-command (reference src/shared_options.hpp domain_t) */
enum { LGPU_DOMAIN_PROTEIN = 0, LGPU_DOMAIN_NUCLEOTIDE = 1, LGPU_DOMAIN_BISULFITE = 2 };

/* ------------------------------------------------------------------------------------------------
 * Index.  Replaces index_file<> + GlobalDataHolder::{transSbjSeqs,redSbjSeqs}
 * (reference src/shared_definitions.hpp:346-379, src/search_algo.hpp:245-321).
 * ---------------------------------------------------------------------------------------------- */

/* Flat view of a loaded index in HOST memory; all pointers borrowed.  Layout of the blobs is the
 * reference's on-disk layout, i.e. exactly what cereal wrote:
 *   occ_blocks : n_blocks x block_bytes; { u32 counts[sigma]; pad to 8; u64 planes[sigma_bits] }
 *                (FMC occtable/InterleavedEPRV2.h:74-76)
 *   super_blocks: n_super x sigma u64        (InterleavedEPRV2.h:147)
 *   C          : sigma+1 u64                 (InterleavedEPRV2.h:148)
 *   ssa        : sampled suffix array, (seqId << bits_for_position) | pos   (FMC CSA.h:17-86)
 *   csa_bv     : n_csa_sb x 48 bytes { u64 entry; u8 blocks[4]; pad4; u64 bits[4] }
 *                (FMC BitvectorCompact.h:21-24)
 *   seqs       : original-alphabet ranks, 1 byte per residue, + n_seqs+1 delimiters */
typedef struct
{
    uint32_t         index_type; /* LGPU_INDEX_*  (only FM is in scope)            */
    uint32_t         orig_alph;  /* LGPU_ALPH_*   alphabet of `seqs`               */
    uint32_t         trans_alph; /* LGPU_ALPH_*                                    */
    uint32_t         red_alph;   /* LGPU_ALPH_*   alphabet the FM index is built on */
    uint32_t         sigma;      /* reduced alphabet size + 1 (0 = sentinel)       */
    uint32_t         sigma_bits;
    uint32_t         block_bytes;
    uint32_t         planes_offset;
    void const *     occ_blocks;
    uint64_t         n_blocks;
    uint64_t const * super_blocks;
    uint64_t         n_super;
    uint64_t const * C;
    uint64_t const * ssa;
    uint64_t         n_ssa;
    void const *     csa_bv;
    uint64_t         n_csa_sb;
    uint64_t         sampling_rate;
    uint64_t         bits_for_position;
    uint8_t const *  seqs;
    uint64_t         n_residues;
    uint64_t const * seq_delims; /* n_seqs + 1 */
    uint64_t         n_seqs;
    char const *     ids;        /* concatenated ids (may be NULL: ids are host-only)  */
    uint64_t const * id_delims;  /* n_seqs + 1                                        */
} lgpu_index_desc;

/* Host-side .lba reader (cereal BinaryOutputArchive layout, reference
 * src/shared_definitions.hpp:330-379).  The returned object owns the file mapping; `desc` points
 * into it and stays valid until lgpu_lba_close(). */
typedef struct lgpu_lba lgpu_lba;
/* Taxonomy stored in the index (index_file::sTaxIds / taxonParentIDs / taxonHeights / taxonNames,
 * src/shared_definitions.hpp:352-356).  Not read by the search; the host needs it for the lowest common
 * ancestor of a record (_writeRecord, src/search_algo.hpp:886-908) and the taxonomy columns / tags.
 * Pointers are NULL / counts 0 when the index was built without taxonomy. */
typedef struct
{
    uint32_t const * s_tax_ids;         /* concatenated tax ids of all subjects                    */
    uint64_t         n_s_tax_ids;
    uint64_t const * s_tax_delims;      /* n_seqs + 1 (NULL: no tax ids in the index)              */
    uint32_t const * taxon_parents;     /* parent of every taxon, indexed by tax id; 0 = unassigned */
    uint8_t const *  taxon_heights;     /* depth of every taxon                                    */
    uint64_t         n_taxa;
    char const *     taxon_names;       /* concatenated scientific names                           */
    uint64_t const * taxon_name_delims; /* n_taxa + 1 (NULL: no names in the index)                */
} lgpu_taxonomy;

int                     lgpu_lba_open(lgpu_lba ** out, char const * path);
lgpu_taxonomy const *   lgpu_lba_taxonomy(lgpu_lba const *);
lgpu_index_desc const * lgpu_lba_desc(lgpu_lba const *);
void                    lgpu_lba_close(lgpu_lba *);

typedef struct lgpu_index lgpu_index;
/* Copies the index to HBM of `device`; the host arrays may be released after return. */
int      lgpu_index_create(lgpu_index ** out, lgpu_index_desc const * host_arrays, int device);
/* A replica of an index that already lives on another GPU of the box, copied device to device (NVLink / NVSwitch when
 * the two devices are peers) instead of once more from the host: with N GPUs the host reads the index file once. */
int      lgpu_index_clone(lgpu_index ** out, lgpu_index const * src, int device);
/* Creates the CUDA context of `device`.  Optional: call it from a thread at program start so that context creation
 * (a few hundred milliseconds) overlaps with reading the query file and mapping the index. */
int      lgpu_device_warmup(int device);
void     lgpu_index_destroy(lgpu_index *);
uint64_t lgpu_index_device_bytes(lgpu_index const *);
/* dbTotalLength / dbNumberOfSeqs as the reference computes them (src/search_algo.hpp:317-319) */
uint64_t lgpu_index_db_total_length(lgpu_index const *);
uint64_t lgpu_index_db_num_seqs(lgpu_index const *);

/* ------------------------------------------------------------------------------------------------
 * Parameters.  Replaces LambdaOptions (reference src/search_options.hpp:60-108) for the fields the
 * hot path reads.  lgpu_params_default() fills the per-domain defaults (:263,290-337) and then
 * applies a profile (:631-682): "none", "fast", "sensitive", "pairs-default", "pairs-sensitive".
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    uint32_t seed_length;
    uint32_t max_seed_dist;
    uint32_t seed_offset;
} lgpu_search_opts;

typedef struct
{
    uint32_t         domain;            /* LGPU_DOMAIN_*                                         */
    lgpu_search_opts opts0;             /* phase-1 seeds (searchOpts0)                           */
    lgpu_search_opts opts;              /* phase-2 / non-iterative seeds (searchOpts)            */
    uint32_t         seed_half_exact;   /* seedHalfExact                                         */
    uint32_t         adaptive_seeding;  /* adaptiveSeeding                                       */
    uint32_t         iterative_search;  /* iterativeSearch                                       */
    uint32_t         max_matches;       /* maxMatches (-n)                                       */
    int32_t          pre_scoring;       /* preScoring                                            */
    double           pre_scoring_thresh;/* preScoringThresh                                      */
    int32_t          scoring_method;    /* 45 / 62 / 80 (protein); ignored for nucleotides       */
    int32_t          gap_open;          /* BLAST convention, e.g. -11 (cost of the first gap
                                           character is gap_open + gap_extend)                  */
    int32_t          gap_extend;        /* e.g. -1                                               */
    int32_t          match;             /* nucleotide match score, e.g. 2                        */
    int32_t          mismatch;          /* nucleotide mismatch score, e.g. -3                    */
    int32_t          min_bit_score;     /* minBitScore, -1 = off                                 */
    double           max_evalue;        /* maxEValue, < 0 = off                                  */
    int32_t          id_cutoff;         /* idCutOff (percent identity)                           */
    uint32_t         finalize;          /* 1: apply writeRecords/_writeRecord (sort, unique,
                                           top max_matches) before returning hits; 0: return the
                                           raw lH.blastMatches multiset                          */
    uint32_t         query_alph;        /* alphabet of the query batch: 0 = the domain's default
                                           (amino acids for searchp, dna5 otherwise);
                                           LGPU_ALPH_DNA5 with a protein index = translated query
                                           (BLASTX / TBLASTX, --query-alphabet dna5);
                                           LGPU_ALPH_AMINO_ACID = protein query (BLASTP / TBLASTN) */
    uint32_t         want_cigar;        /* 1: also return the gapped rows of every hit as run-length
                                           operations (needed for SAM / pairwise output)          */
    uint32_t         window_band;       /* 0: the reference's rule, band = floor(sqrt(query length)) + 1
                                           (_bandSize, src/search_misc.hpp:46-50); > 0: that many subject
                                           residues on either side of the seed diagonal's window instead
                                           (src/search_algo.hpp:929-937).  The DP over the window stays
                                           unbanded as in the reference (:1102).  Any value other than 0
                                           leaves parity with the reference binary (it has no such
                                           option); the oracle restates it for the band sweep of
                                           BASELINE configs[3].                                   */
} lgpu_params;

int lgpu_params_default(lgpu_params * out, uint32_t domain, char const * profile);

/* ------------------------------------------------------------------------------------------------
 * Context = the reference's per-thread LocalDataHolder (src/search_datastructures.hpp:399-531):
 * device work buffers, one CUDA stream.  One context per host thread; contexts of one index may be
 * used concurrently.
 * ---------------------------------------------------------------------------------------------- */
typedef struct lgpu_ctx lgpu_ctx;
int          lgpu_ctx_create(lgpu_ctx ** out, lgpu_index const *, lgpu_params const *);
void         lgpu_ctx_destroy(lgpu_ctx *);
/* Number of sub-batches a large lgpu_search_batch() call is cut into; each runs the whole pipeline on
 * its own CUDA stream + host thread so that host-only and latency-bound stretches overlap with the
 * kernels of the others (default: 1 for protein searches, whose host part is under a millisecond now that the
 * records are finalised on the device, 4 for nucleotide / bisulfite searches; env LAMBDA_B200_STREAMS; 1 = strictly
 * serial, used for per-kernel timing; calls with fewer than ~32k queries are not cut).  Results do not depend on it. */
int          lgpu_ctx_set_streams(lgpu_ctx *, uint32_t n);
char const * lgpu_last_error(lgpu_ctx const *); /* ctx may be NULL: error of the last failed
                                                   create/open call on this thread */

/* Query batch: original-alphabet ranks (aa27 ranks for protein queries, dna5 ranks for nucleotide
 * queries -- searchn, searchbs and translated searchp), all
 * sequences concatenated, offsets[n_queries + 1].  Pointers are HOST memory unless
 * `on_device` != 0 (then they are device pointers on the context's device and the H2D copy is
 * skipped -- used to time the resident-input path). */
typedef struct
{
    uint8_t const *  residues;
    uint64_t const * offsets;
    uint64_t         n_queries;
    uint32_t         on_device;
} lgpu_query_batch;

/* One reported alignment = one seqan::BlastMatch of lH.blastMatches
 * (SQ/blast/blast_record.h; fields filled at src/search_algo.hpp:1205-1223,1032-1035,1308-1322).
 * Coordinates are 0-based half-open in translated/strand space, like the reference's members. */
typedef struct
{
    uint32_t q_id;       /* _n_qId: index of the query inside the batch              */
    uint32_t s_id;       /* _n_sId: index of the subject in the index                */
    uint32_t q_start, q_end, s_start, s_end;
    uint32_t q_len, s_len; /* qLength / sLength (original sequences)                 */
    int32_t  score;      /* alignStats.alignmentScore                                 */
    uint32_t n_match, n_mismatch, n_gap_open, n_gap_ext, n_positive, aln_len;
    int8_t   q_frame, s_frame; /* qFrameShift / sFrameShift                           */
    uint8_t  phase;      /* 1 = found with searchOpts0 (phase 1), 2 = phase 2        */
    uint8_t  reserved;
    double   bit_score;  /* computeBitScore, SQ/blast/blast_statistics.h:1027        */
    double   evalue;     /* computeEValueThreadSafe, src/search_misc.hpp:57          */
    uint32_t cigar_off, cigar_len; /* lgpu_params.want_cigar: the gapped rows (alignRow0 / alignRow1 of the
                                      BlastMatch) as run-length operations cigar_ops[cigar_off ..
                                      cigar_off + cigar_len), in traceback order (alignment END first)    */
} lgpu_hit;

/* One run of the alignment: length << 2 | kind, kind 0 = aligned columns (M), 1 = query residues against a
 * gap in the subject row (I), 2 = subject residues against a gap in the query row (D)
 * (src/search_output.hpp:116-196 builds the SAM CIGAR from exactly these runs). */
#define LGPU_CIGAR_M 0u
#define LGPU_CIGAR_I 1u
#define LGPU_CIGAR_D 2u

typedef struct
{
    lgpu_hit const * hits; /* owned by the context until the next search call */
    uint64_t         n;
    uint32_t const * cigar_ops; /* NULL unless lgpu_params.want_cigar */
    uint64_t         n_cigar_ops;
} lgpu_hits;

/* The reference's StatsHolder funnel (src/search_datastructures.hpp:70-215) + device timings. */
typedef struct
{
    uint64_t hits_after_seeding;
    uint64_t hits_failed_pre_extend;
    uint64_t hits_failed_evalue;
    uint64_t hits_failed_bitscore;
    uint64_t hits_failed_identity;
    uint64_t hits_duplicate;  /* merged seeds (_widenAndPreprocessMatches)    */
    uint64_t hits_duplicate2; /* late duplicates (_writeRecord)               */
    uint64_t hits_abundant;
    uint64_t hits_final;
    uint64_t pairs;
    uint64_t qrys_with_hit;
    /* work actually done on the device */
    uint64_t n_extensions_score; /* DP pass-1 alignments                       */
    uint64_t n_extensions_trace; /* DP pass-2 alignments                       */
    uint64_t cells_score;        /* sum qlen x wlen over pass-1 alignments     */
    uint64_t cells_trace;        /* same for pass 2                            */
    uint64_t kernel_launches;    /* kernels launched by this library           */
    /* CUDA-event times on the context's stream, milliseconds */
    float ms_seed, ms_sort_merge, ms_extend_score, ms_extend_trace, ms_h2d, ms_d2h, ms_total;
    float ms_host; /* host-only work inside the call: score thresholds, e-values, record finalisation */
} lgpu_stats;

/* search() + iterateMatches() + iterativeSearchPre/Post [+ writeRecords] for one batch.  Blocking.
 * `stats` is accumulated into (zero it yourself), like the reference's per-thread StatsHolder. */
int lgpu_search_batch(lgpu_ctx *, lgpu_query_batch const * q, lgpu_hits * out, lgpu_stats * stats);

/* The records of the LAST lgpu_search_batch() call also stay on the device (they are sorted, made unique and cut to
 * max_matches there -- writeRecords / _writeRecord, src/search_algo.hpp:821-913,1335-1362).
 * lgpu_ctx_export_hits() copies them into a caller-owned DEVICE buffer of `cap_records` records, query ids rebased by
 * `first_query` -- the send buffer of the multi-GPU hit gather, so that the records of a shard travel GPU to GPU
 * without a detour through the host.  *n_out = number of records of the call; nothing is copied if it exceeds
 * `cap_records`.  The device records carry bit_score = evalue = 0: both are functions of (score, q_len) that the host
 * computes with its own libm -- lgpu_hits_fill_scores() does that for records in HOST memory (e.g. gathered ones). */
int lgpu_ctx_export_hits(lgpu_ctx *, void * dev_dst, uint64_t cap_records, uint64_t first_query, uint64_t * n_out);
int lgpu_hits_fill_scores(lgpu_ctx *, lgpu_hit * host_records, uint64_t n);

/* ------------------------------------------------------------------------------------------------
 * Stage-level entry points (what the reference's unit-less test-suite lacks; used by the parity
 * tests to compare each kernel with the oracle separately).
 * ---------------------------------------------------------------------------------------------- */

/* reference Match (src/search_datastructures.hpp:46-61), positions narrowed to u32 */
typedef struct
{
    uint32_t qry_id;  /* frame-expanded: query * qryNumFrames + frame */
    uint32_t subj_id; /* frame-expanded subject id                    */
    uint32_t qry_start, qry_end, subj_start, subj_end;
} lgpu_match;

/* search() only (src/search_algo.hpp:607-762) with the phase-`phase` (1|2) seed options.
 * Returns the pre-score-passing matches (order unspecified) and adds to the seeding counters. */
int lgpu_seed_batch(lgpu_ctx *, lgpu_query_batch const * q, int phase, lgpu_match const ** matches,
                    uint64_t * n_matches, lgpu_stats * stats);

/* _widenAndPreprocessMatches (src/search_algo.hpp:1137-1175) on caller-supplied matches.
 * Output = the merged, sorted, unique extension windows. */
int lgpu_merge_matches(lgpu_ctx *, lgpu_query_batch const * q, lgpu_match const * matches, uint64_t n,
                       lgpu_match const ** merged, uint64_t * n_merged, lgpu_stats * stats);

/* DP pass 1 (_performAlignment<withTrace=false>, src/search_algo.hpp:1246) on caller-supplied
 * windows: scores[i] = local affine score of whole query (frame qry_id) vs subject window. */
int lgpu_extend_scores(lgpu_ctx *, lgpu_query_batch const * q, lgpu_match const * windows, uint64_t n,
                       int32_t * scores, lgpu_stats * stats);

/* DP pass 2 (+ traceback, _expandAlign, computeAlignmentStats; src/search_algo.hpp:1296-1322) on
 * caller-supplied windows; no e-value filtering.  out[i] corresponds to windows[i]. */
int lgpu_extend_trace(lgpu_ctx *, lgpu_query_batch const * q, lgpu_match const * windows, uint64_t n,
                      lgpu_hit * out, lgpu_stats * stats);

/* FM-index primitives on the device, for known-answer tests:
 * rank(idx, symb)  (InterleavedEPRV2.h:211-216) and locate(row) (ReverseFMIndex.h:62-91). */
int lgpu_fm_rank(lgpu_index const *, uint64_t const * idx, uint8_t const * symb, uint64_t n, uint64_t * out);
int lgpu_fm_locate(lgpu_index const *, uint64_t const * rows, uint64_t n, uint64_t * subj, uint64_t * pos);

/* ------------------------------------------------------------------------------------------------
 * Host-side statistics and output helpers (kept on the host like the reference: a10/a14/App. C).
 * ---------------------------------------------------------------------------------------------- */
int lgpu_bit_score(lgpu_params const *, int32_t raw_score, double * out);
/* Karlin-Altschul parameters of the scoring scheme (SQ/blast/blast_statistics.h tables) and the substitution
 * matrix the alignments use (32 x 32 int8, [query rank * 32 + subject rank]); for report-style output. */
int lgpu_ka_params(lgpu_params const *, double * lambda, double * k, double * h);
int lgpu_score_matrix(lgpu_params const *, int8_t * out_32x32);
int lgpu_evalue(lgpu_params const *, int32_t raw_score, uint64_t query_len, uint64_t db_total_len,
                double * out);
/* smallest raw score that passes both the bit-score and the e-value filter for this query length */
int lgpu_min_raw_score(lgpu_params const *, uint64_t query_len, uint64_t db_total_len, int32_t * out);

/* BLAST tabular (.m8) "std" line for one hit, formatted like SQ/blast/blast_tabular_out.h:248-400.
 * Writes at most `cap` bytes including the trailing '\n' and NUL; returns the line length. */
int lgpu_format_m8(lgpu_params const *, lgpu_hit const *, char const * q_id, char const * s_id, char * buf,
                   size_t cap);

/* BLAST tabular line with a custom column list (`lambda3 --output-columns`, src/search_options.hpp:224-232,
 * 710-760).  A column is identified by its index in the reference's BlastMatchField::Enum
 * (SQ/blast/blast_tabular.h:403-454): lgpu_tabular_column("qseqid") -> 1, "std" -> 0 (expands to the twelve
 * default columns), unknown label -> -1.  lgpu_tabular_column_label() returns the text the "# Fields:" comment
 * line of .m9 uses for that column (NULL if out of range), lgpu_tabular_column_supported() whether a line can be
 * formatted with it: columns the reference itself does not implement print "n/i" like there; the taxonomy
 * columns (staxids, lcaid, lcataxid) are not supported (the taxonomy of the index is not loaded).
 * lgpu_format_tabular() returns the line length, LGPU_ERR_ARG for bad arguments / unsupported columns. */
int          lgpu_tabular_column(char const * option_label);
char const * lgpu_tabular_column_name(uint32_t column); /* the option label of a column, NULL if out of range */
char const * lgpu_tabular_column_label(uint32_t column);
int          lgpu_tabular_column_supported(uint32_t column);
int          lgpu_tabular_column_implemented(uint32_t column); /* BlastMatchField::implemented: 0 = prints "n/i" */
int          lgpu_format_tabular(lgpu_params const *, lgpu_hit const *, char const * q_id, char const * s_id,
                                 uint32_t const * columns, size_t n_columns, char * buf, size_t cap);

int lgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LAMBDA_B200_H */
