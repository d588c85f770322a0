/*
 * mmalign.h -- C ABI of the B200-native alignment-scoring and retrieval path.
 *
 * The reference (guille-gil/Multimodal-Alignment-of-Noisy-Image-Text-Pairs-using-
 * Weak-Supervision) has no FFI of its own: its scoring path is Python that sends
 * SQL to PostgreSQL/pgvector.  The entry points below are what a binding for
 * that path replaces; each cites the reference site (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MMALIGN_E* code and never
 *     throws; mmalign_last_error() gives the text.
 *   - all array arguments may be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes).  Host inputs are uploaded into buffers owned by
 *     the context; device inputs are BORROWED and must stay valid until the next
 *     set_* call or mmalign_destroy.  Host outputs are filled with
 *     device-to-host copies before the call returns.
 *   - set_images / set_chunks are ASYNCHRONOUS: they queue their uploads (a copy
 *     stream of the context) and the operand preparation behind them and return.
 *     Pageable host inputs have been consumed when the call returns; PAGE-LOCKED
 *     host inputs (cudaHostAlloc / cudaHostRegister) are borrowed like device
 *     inputs until the next call that waits for them: mmalign_run and every other
 *     entry point that reads the tables, or mmalign_sync.  Device inputs must be
 *     complete on the legacy default stream when set_* is called.
 *   - mmalign_run overlaps what it can: with host embeddings and/or host outputs it
 *     cuts the query rows into slabs, so that the upload of slab s+1 and the
 *     download of slab s-1's results run beside the kernels of slab s (copy
 *     engines, separate streams); the results are the same bytes either way.
 *   - one context per (process, device); kernels are ordered on the stream passed
 *     to mmalign_run (a cudaStream_t, NULL = default stream); not thread-safe per
 *     context.
 *   - there is NO CPU fallback: a device that is not sm_100 is an error.
 *   - ids never cross the ABI: the caller maps image_id / chunk_id to dense row
 *     indices ("lower index wins a tie") and (manual_id, page) to a 64-bit page
 *     key; MMALIGN_NULL_KEY is SQL NULL and never joins.
 */
#ifndef MMALIGN_H
#define MMALIGN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMALIGN_ABI_VERSION 3
#define MMALIGN_NULL_KEY 0xFFFFFFFFFFFFFFFFull

enum {
    MMALIGN_OK = 0,
    MMALIGN_EINVAL = -1,   /* bad argument */
    MMALIGN_ECUDA = -2,    /* CUDA runtime / driver error */
    MMALIGN_EDEVICE = -3,  /* not an sm_100 device */
    MMALIGN_ESTATE = -4,   /* call order (run before set_*) */
    MMALIGN_ELIMIT = -5    /* a documented capacity limit was exceeded */
};

/* schema bits: src/insert_clip_embeddings.py:446-453 (use_lexical, use_positional) */
enum {
    MMALIGN_VANILLA = 1u,    /* vanilla_clip    (F, F) */
    MMALIGN_LEXICAL = 2u,    /* clip_lexical    (T, F) */
    MMALIGN_POSITIONAL = 4u, /* clip_positional (F, T) */
    MMALIGN_COMBINED = 8u,   /* clip_combined   (T, T) */
    MMALIGN_RAW_SCORES = 0x100u /* mmalign_alignments only: rec = {lexical, positional, 0} before
                                   the 0.05 / 0.1 thresholds of :387, :393, :400 */
};

enum {
    MMALIGN_CAND_SAME_PAGE = 0, /* the reference's join: src/evaluate_alignments.py:128-131 */
    MMALIGN_CAND_ALL = 1        /* full N x M ranking (BASELINE.json north_star) */
};

enum {
    MMALIGN_PATH_AUTO = 0,  /* tcgen05 fused kernel + exact rescoring, exact scan for uncertified rows */
    MMALIGN_PATH_EXACT = 1, /* exact fp32 scan of every row (slow; validation) */
    MMALIGN_PATH_FUSED = 2  /* like AUTO but an uncertified row is an error instead of a rescan */
};

typedef struct mmalign_ctx mmalign_ctx;

typedef struct {
    uint32_t schema_mask;  /* which of the four schemas to rank, in bit order */
    int32_t candidates;    /* MMALIGN_CAND_* */
    int32_t n_k;           /* number of K values, 1..8 */
    int32_t k_list[8];     /* src/evaluate_alignments.py:388 [1,5,10]; :275 [1,5,10,20] */
    int32_t mrr_cutoff;    /* src/evaluate_alignments.py:206: 100 */
    double lam_lex;        /* weight of a 'lexical' alignment record in the ranking score */
    double lam_pos;        /* weight of a 'positional' record */
    double lam_comb;       /* weight of a 'combined' record; all 0 = reference ranking */
    int32_t path;          /* MMALIGN_PATH_* */
    int32_t kprime;        /* depth the fused kernel's candidate lists are complete to; 0 = auto */
    int32_t n_ranks;       /* sharded passes: number of GPUs the chunk table is sharded over; 0/1 = one */
    float eps_scale;       /* multiplies the certificate's error bound eps; 0 = 1.0.  Values above 1 only make the
                              certificate more conservative (more rows take the exact scan), never less exact */
    int64_t shard_col0;    /* mmalign_fused_pass: the contraction runs on chunk rows [shard_col0, shard_col0 + */
    int64_t shard_cols;    /*   shard_cols) of the table given to set_chunks; 0, 0 = the whole table          */
    int64_t slab_row0;     /* mmalign_rescore_slab: image rows [slab_row0, slab_row0 + slab_rows) are ranked;  */
    int64_t slab_rows;     /*   0, 0 = every image                                                             */
    int32_t pipeline_rows; /* mmalign_run: query rows per pipeline slab; 0 = auto (4 waves of 128-row blocks when    */
    int32_t reserved;      /*   inputs or outputs live on the host, else one slab), -1 = never cut; reserved = 0     */
} mmalign_params;

/* Any pointer may be NULL (output not wanted).  S = popcount(schema_mask),
 * Kmax = max(k_list), P = mmalign_num_pairs(). */
typedef struct {
    int64_t *topk_idx;    /* [S][N][Kmax] global chunk index, -1 padded          (:109-143) */
    double *topk_score;   /* [S][N][Kmax] ranking score, -inf padded                        */
    int32_t *pair_rank;   /* [S][P] 1-based rank of each true pair within
                             max(Kmax, mrr_cutoff), else 0                       (:188-190, :208-214) */
    double *pair_sim;     /* [P] exact cosine of each true pair                  (:72-106)  */
    int64_t *hits;        /* [S][n_k] pairs with rank <= k                       (:182-192) */
    double *rr_sum;       /* [S] sum of 1/rank over pairs with rank <= mrr_cutoff (:203-216) */
    double *sim_sum;      /* [1] sum of pair_sim                                 (:226-231) */
    int64_t *num_pairs;   /* [1] P                                               (:393-395) */
    double *pair_score;   /* [S][P] ranking score of each true pair (multi-GPU rank step)   */
    int64_t *deep_idx;    /* [S][N][max(Kmax, mrr_cutoff)] lists to the full exact depth    */
    double *deep_score;   /*        (multi-GPU rank step; see mmalign_count_beating)        */
    int64_t *stats;       /* [16] 0: rows rescanned exactly, 1: candidates rescored,
                             2: fused-kernel launches, 3: kernels launched by the call,
                             4: K' used, 5/6/7: microseconds (CUDA events) of the fused
                             kernel / the rescoring kernel / the exact rescan, 8: rows whose
                             re-scored candidates broke the certificate's error bound (they were
                             rescanned exactly), 9: pipeline slabs of the run, 10: SMs the rescoring of a slab
                             had beside the contraction of the next one (0 = no overlap), 11..15: 0 */
} mmalign_out;

int mmalign_abi_version(void);

/* replaces connect_db(): src/evaluate_alignments.py:37-45 (one context instead of
 * one TCP connection per call) */
int mmalign_create(mmalign_ctx **ctx, int device);
void mmalign_destroy(mmalign_ctx *ctx);
const char *mmalign_last_error(const mmalign_ctx *ctx); /* ctx may be NULL */

/* replaces the `images` / `text_chunks` tables: src/setup_vector_db.py:102-131.
 *   emb       [n][D] fp32 clip_embedding (any norm; cosine re-normalises like `<=>`)
 *   page_key  [n]    (manual_id, page) packed by the caller, MMALIGN_NULL_KEY = NULL
 *   bbox      [n][4] x0,y0,x1,y1 as doubles (JSON floats); all-zero = missing
 *   terms     [n][term_words] bit t = lexical term t occurs in the item
 *             (src/insert_clip_embeddings.py:149-150); NULL = every term (the
 *             reference's images carry no term set)
 * set_chunks takes this rank's shard: rows [col_offset, col_offset + m_local)
 * of the global chunk table. */
int mmalign_set_images(mmalign_ctx *ctx, const float *emb, const uint64_t *page_key,
                       const double *bbox, const uint64_t *terms, int64_t n, int32_t D,
                       int32_t term_words);
int mmalign_set_chunks(mmalign_ctx *ctx, const float *emb, const uint64_t *page_key,
                       const double *bbox, const uint64_t *terms, int64_t m_local, int32_t D,
                       int32_t term_words, int64_t n_terms, int64_t col_offset);

/* tunables of a context (none changes a result):
 *   "piece_bytes"   host embedding rows are uploaded in pieces of about this many bytes (default 64 MiB; each piece
 *                   announces itself, so preparation and the first contraction start behind the first pieces
 *                   instead of behind the whole table)
 *   "cta_pairs"     the fused kernel on clusters of two CTAs: 0 = one CTA per SM on its own (default),
 *                   1 = tcgen05.mma.cta_group::2 (one M = 256 MMA per pair, half of B per CTA),
 *                   2 = cta_group::1 MMAs per CTA over a B ring the pair fills together by TMA multicast
 *   "k2_sms"        SMs left to the exact rescoring of slab s while slab s+1 is contracted (default 0: one after
 *                   the other)
 *   "compact_one"   routine compaction of the candidate lists: 0 = every flagged list at once (default), 1 = one
 *                   list per gap between two tiles, 2 = split over two gaps
 *   "epi_sleep_ns"  pause between the epilogue warps' polls of their accumulator barrier (default 0) */
int mmalign_set_option(mmalign_ctx *ctx, const char *name, int64_t value);

/* waits for every upload and preparation queued by set_images / set_chunks (after it, page-locked host
 * inputs may be reused or freed) */
int mmalign_sync(mmalign_ctx *ctx);

/* Encoder output in half precision (the batched CLIP encode of
 * src/insert_clip_embeddings.py:91-141, :281-353 stays on OpenCLIP; its fp16 / bf16 batch can feed this path
 * without an fp32 hop through the host).  emb [n][D] of emb_dtype, host or device; the rows are widened (exactly)
 * into fp32 master rows owned by the context.  Everything else as mmalign_set_images / mmalign_set_chunks. */
enum { MMALIGN_F32 = 0, MMALIGN_F16 = 1, MMALIGN_BF16 = 2 };
int mmalign_set_images_half(mmalign_ctx *ctx, const void *emb, int32_t emb_dtype, const uint64_t *page_key,
                            const double *bbox, const uint64_t *terms, int64_t n, int32_t D, int32_t term_words);
int mmalign_set_chunks_half(mmalign_ctx *ctx, const void *emb, int32_t emb_dtype, const uint64_t *page_key,
                            const double *bbox, const uint64_t *terms, int64_t m_local, int32_t D, int32_t term_words,
                            int64_t n_terms, int64_t col_offset);

/* pgvector interop: the `images` / `text_chunks` tables of src/setup_vector_db.py:102-131
 * as a PostgreSQL binary COPY stream, e.g.
 *   COPY (SELECT image_id, manual_id, page, bbox, clip_embedding FROM s.images ORDER BY id) TO STDOUT (FORMAT binary)
 * mmalign_copy_scan   host-only walk of the tuples (variable-length fields): field_off [rows][n_cols] = byte offset
 *                     of each field's data, field_len = its length, -1 = NULL.  Returns the number of tuples
 *                     (with field_off == NULL it only counts), or -1 bad signature / WITH OIDS, -2 truncated stream,
 *                     -3 a tuple with another field count, -4 more than `cap` tuples.
 * mmalign_copy_decode GPU decode of the bulk columns into the arrays set_images / set_chunks take: the `vector(D)`
 *                     column (pgvector vector_send: int16 dim, int16 0, float4[dim], big-endian) -> emb [n][D] fp32;
 *                     the REAL[] bbox column -> bbox [n][4] doubles (NULL / wrong-length / NULL-element boxes -> zeros,
 *                     which score 0.0 like src/insert_clip_embeddings.py:161-169); the INTEGER page column -> page [n],
 *                     page_null [n].  data / outputs host or device; field_off / field_len host or device; any output
 *                     may be NULL; bbox_col / page_col < 0 = column absent.  MMALIGN_EINVAL if a vector field is NULL
 *                     or not of dimension D. */
int64_t mmalign_copy_scan(const uint8_t *data, int64_t n_bytes, int32_t n_cols, int64_t *field_off, int32_t *field_len,
                          int64_t cap);
int mmalign_copy_decode(mmalign_ctx *ctx, const uint8_t *data, int64_t n_bytes, const int64_t *field_off,
                        const int32_t *field_len, int64_t n, int32_t n_cols, int32_t vec_col, int32_t bbox_col,
                        int32_t page_col, int32_t D, float *emb, double *bbox, int32_t *page, uint8_t *page_null,
                        void *stream);

/* replaces get_image_text_pairs(): src/evaluate_alignments.py:48-69.  Pairs are
 * ordered by (image index, chunk index).  pair_offsets [N+1], pair_chunk [P]
 * (global chunk index); either may be NULL. */
int mmalign_num_pairs(mmalign_ctx *ctx, int64_t *num_pairs);
int mmalign_get_pairs(mmalign_ctx *ctx, int64_t *pair_offsets, int64_t *pair_chunk);

/* replaces get_top_k_similar_chunks / compute_similarity / compute_top_k_accuracy /
 * compute_mrr / compute_average_similarity: src/evaluate_alignments.py:72-231 */
int mmalign_run(mmalign_ctx *ctx, const mmalign_params *params, mmalign_out *out, void *stream);

/* replaces the alignment loop of insert_embeddings():
 * src/insert_clip_embeddings.py:369-414.  rec [P][3] = weak_score of the
 * 'lexical' / 'positional' / 'combined' record of each true pair in `schema`
 * (one MMALIGN_* bit), 0.0 where the reference inserts none. */
int mmalign_alignments(mmalign_ctx *ctx, uint32_t schema, double *rec, void *stream);

/* Sharded (multi-GPU) run in passes, fully sharded variant (no rank holds another rank's chunks) -- the
 * caller performs the collectives in between (distributed.py::AllGatherScorer):
 *   1. mmalign_fused_pass    K1 on this rank's shard; tau_row [N] = score above which the rank's candidate
 *                            lists hold every local column.       -> all-reduce(max) of tau_row, and of
 *                            mmalign_chunk_err_max
 *   2. mmalign_rescore_pass  exact scores of the rank's entries above the global tau; fills the device
 *                            outputs (lists, pair arrays) and cert_count [S][N] = local entries provably
 *                            above every column left out.         -> all-reduce(sum) of cert_count; rows whose
 *                            sum is below max(Kmax, mrr_cutoff) are not certified
 *   3. mmalign_rescan_rows   exact scan of those rows (same rows on every rank)
 *   then merge_topk / count_beating / reduce_metrics as below.  Output pointers are DEVICE pointers. */
int mmalign_fused_pass(mmalign_ctx *ctx, const mmalign_params *params, float *tau_row, void *stream);
int mmalign_chunk_err_max(mmalign_ctx *ctx, float *err_max);
int mmalign_rescore_pass(mmalign_ctx *ctx, const mmalign_params *params, const float *tau_global,
                         float eps_chunk_global, mmalign_out *out, int32_t *cert_count, void *stream);
int mmalign_rescan_rows(mmalign_ctx *ctx, const mmalign_params *params, const int32_t *rows,
                        int64_t n_rows, mmalign_out *out, void *stream);

/* Sharded run, default exchange (distributed.py::ShardedScorer): the contraction is sharded by CHUNK
 * columns, the exact rescoring by QUERY rows.  Every rank holds the whole corpus (each rank ingests 1/G
 * of it and the rest arrives by an NCCL all-gather over NVLink) and calls set_images / set_chunks on all of it.
 *   1. mmalign_fused_pass    with params.shard_col0 / shard_cols = this rank's column range (tau_row may be NULL)
 *   2. mmalign_export_lists  packs, for each destination rank d, the candidates of the image rows of slab d:
 *                            keys [n_dest][slab_rows][stride] (lo = GLOBAL chunk index, hi = fp32 score bits),
 *                            count [n_dest][slab_rows] (-1 = the row overflowed `stride`: it will be scanned
 *                            exactly), tau [n_dest][slab_rows] (the rank's columns above tau are all present).
 *                            stride >= mmalign_list_stride() of every rank.   -> all-to-all of the three arrays
 *   3. mmalign_rescore_slab  with params.slab_row0 / slab_rows = this rank's query slab and the received lists
 *                            keys [n_src][list_rows][stride], count / tau [n_src][list_rows] (list_rows = the
 *                            slab_rows of the export; the rank's last rows may be padding): exact rescoring, certificate, exact scan of uncertified rows, metric
 *                            sums -- mmalign_run restricted to the slab; outputs are sized by the slab:
 *                            topk [S][slab_rows][Kmax], pair arrays [S][P_slab] (mmalign_num_pairs_range).
 *                            -> all-reduce(sum) of hits / rr_sum / sim_sum / num_pairs
 * Device pointers for keys / count / tau. */
int mmalign_list_stride(mmalign_ctx *ctx, int32_t *stride);

/* Sharded run, query-row layout (distributed.py::ShardedScorer, contraction="rows"): rank g ranks its own
 * query slab against the WHOLE chunk table.  Each rank prepares its own chunk shard and the prepared operands
 * travel, so that the contraction can start after the small exchange while the fp32 master rows (needed by the
 * exact rescoring only) are still in flight:
 *   1. mmalign_prep_rows           K0 (src/insert_clip_embeddings.py:113-115 normalisation, bf16 rounding, sum of
 *                                  squares, rounding-error norm) of this rank's shard into its slot of the gathered
 *                                  buffers; device pointers     -> all-gather of bf16 / norm2 / err / page keys
 *   2. mmalign_set_chunks_prepared the gathered table: fp32 rows, keys, boxes, term sets and the prepared operands
 *                                  (all DEVICE pointers, borrowed)
 *   3. mmalign_rescore_after       the event (cudaEvent_t) after which the fp32 rows, boxes and term sets are
 *                                  complete -- the all-gather of those runs on a side stream; the next mmalign_run
 *                                  makes its exact rescoring (not the contraction) wait for it
 *   4. mmalign_set_images (own slab) + mmalign_run                -> all-reduce(sum) of the metric sums */
int mmalign_prep_rows(mmalign_ctx *ctx, const float *emb, int64_t n, int32_t D, void *bf16_out, float *norm2_out,
                      float *err_out, void *stream);
int mmalign_set_chunks_prepared(mmalign_ctx *ctx, const float *emb, const uint64_t *page_key, const double *bbox,
                                const uint64_t *terms, const void *bf16, const float *norm2, const float *err,
                                int64_t m, int32_t D, int32_t term_words, int64_t n_terms, int64_t col_offset,
                                void *stream);
int mmalign_rescore_after(mmalign_ctx *ctx, void *event);
int mmalign_export_lists(mmalign_ctx *ctx, int32_t n_dest, int64_t slab_rows, int32_t stride,
                         uint64_t *keys, int32_t *count, float *tau, void *stream);
int mmalign_rescore_slab(mmalign_ctx *ctx, const mmalign_params *params, const uint64_t *keys,
                         const int32_t *count, const float *tau, int32_t n_src, int64_t list_rows,
                         int32_t stride, mmalign_out *out, void *stream);
int mmalign_num_pairs_range(mmalign_ctx *ctx, int64_t row0, int64_t rows, int64_t *num_pairs);

/* cross-rank merge after an all-gather of every rank's mmalign_run lists
 * (chunks are sharded over ranks):
 *   in_idx/in_score [G][S*N][K]  ->  out_idx/out_score [S*N][K]
 * ordered by (score desc, index asc). */
int mmalign_merge_topk(mmalign_ctx *ctx, const int64_t *in_idx, const double *in_score,
                       int32_t n_ranks, int64_t n_lists, int32_t K, int64_t *out_idx,
                       double *out_score, void *stream);

/* multi-GPU rank step: how many entries of THIS rank's exact lists (deep_idx /
 * deep_score [S][N][K] from mmalign_run, K = max(Kmax, mrr_cutoff)) beat each
 * query pair (q_image, q_chunk global index, q_score [S][n_q]).  counts [S][n_q];
 * summed over ranks, rank = 1 + sum if sum < cutoff, else 0.  Device pointers. */
int mmalign_count_beating(mmalign_ctx *ctx, const int64_t *deep_idx, const double *deep_score,
                          int64_t N, int32_t S, int32_t K, int64_t n_q, const int64_t *q_image,
                          const int64_t *q_chunk, const double *q_score, int32_t *counts,
                          void *stream);

/* metric sums from per-pair arrays (deterministic order): hits [S][n_k], rr_sum [S], sim_sum [1] */
int mmalign_reduce_metrics(mmalign_ctx *ctx, const int32_t *pair_rank, const double *pair_sim,
                           int32_t S, int64_t P, const int32_t *k_list, int32_t n_k,
                           int32_t mrr_cutoff, int64_t *hits, double *rr_sum, double *sim_sum,
                           void *stream);

/* Ingest: the chunks' lexical term sets, built on the GPU.  Replaces the substring scans of
 * compute_lexical_alignment, src/insert_clip_embeddings.py:149-150
 *     chunk_text_lower = text_chunk["text"].lower();  term in chunk_text_lower
 * for the whole table at once.
 *   text      the chunks' texts, ALREADY LOWER-CASED (str.lower() is Unicode-aware and stays with the caller),
 *             UTF-8, concatenated; host or device pointer
 *   text_off  [m + 1] byte offsets into text (text_off[0] = 0); host or device pointer
 *   terms     the lexical components (src/insert_clip_embeddings.py:237-239), UTF-8, concatenated; HOST pointer
 *   term_off  [n_terms + 1] byte offsets into terms; HOST pointer.  n_terms <= 4096
 *   bits      [m][term_words] out: bit t of row j = term t is a substring of chunk j's text (an empty term is a
 *             substring of every text, as in Python); term_words >= ceil(n_terms / 64); host or device pointer
 * The result is what mmalign_set_chunks takes as `terms`. */
int mmalign_term_bitsets(mmalign_ctx *ctx, const uint8_t *text, const int64_t *text_off, int64_t m,
                         const uint8_t *terms, const int64_t *term_off, int32_t n_terms,
                         int32_t term_words, uint64_t *bits, void *stream);

/* validation hook: the raw bf16 x bf16 -> fp32 score tile matrix of the fused
 * kernel, out [N][m_local] fp32 (small sizes only). */
int mmalign_debug_scores(mmalign_ctx *ctx, float *out, void *stream);
/* validation hook: the bf16 operands K0 prepared (L2-normalised rows, rounded to nearest even), [N][D] and
 * [m_local][D] 16-bit values; either may be NULL; host or device pointers. */
int mmalign_debug_operands(mmalign_ctx *ctx, void *img_bf16, void *chk_bf16, void *stream);
/* validation hook: the checked build of the library (csrc/Makefile `make check`: every kernel tests its own
 * indices and invariants -- list capacities, shared-memory slots, chunk columns, output positions) reports what it
 * saw since the process started.  out[9]: out[0] = 1 for a checked build, 0 for the release build (which compiles
 * the checks out and reports zeros); then {violations, first failing source line} for fused_tc.cu, rescore.cu,
 * prep.cu, ingest.cu.  Synchronises the device. */
int mmalign_check_report(uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* MMALIGN_H */
