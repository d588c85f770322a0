/*
 * mmalign_oracle.c -- CPU restatement of the reference's alignment-scoring and
 * retrieval path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file.  The product path (the CUDA library in
 * multimodal-..._b200/csrc) never links, imports or calls it.
 *
 * Parity status: the weak-supervision functions (H6/H7/H8) and the metric
 * functions (H4/H5/H5b) are PINNED against the unmodified reference code run in
 * this container (tests/golden/, made by tests/golden/make_golden.py).
 * The cosine arithmetic itself (H2/H3) lives in the third-party pgvector
 * PostgreSQL extension, which is absent from /root/reference and is not pinned
 * to a version there: for that one function this file restates pgvector's
 * published algorithm (fp32 accumulation of dot, |a|^2, |b|^2; double division by
 * sqrt(na*nb); clamp; distance = 1 - sim) -> "parity unpinned" for the cosine
 * kernel, pinned for everything built on top of it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NULL_KEY 0xFFFFFFFFFFFFFFFFull /* SQL NULL page: never joins */

/* ------------------------------------------------------------------------ */
/* Cosine (pgvector `<=>`), called at src/evaluate_alignments.py:97 and :128 */
/* ------------------------------------------------------------------------ */

/*
 * fp32 dot product in the CANONICAL summation order shared with the CUDA
 * kernels: 128 interleaved fused-multiply-add chains (chain j takes k == j mod
 * 128, increasing k), folded 4 -> 1 per "lane" and then by an xor butterfly
 * over 32 lanes.  pgvector's own loop is auto-vectorised by whatever compiler
 * built the server (-ftree-vectorize -fassociative-math), so its order is not
 * defined; this is one valid realisation, chosen so that CPU and GPU agree
 * bit for bit.
 */
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("arch=haswell", "default"))) /* fmaf is exact either way */
#endif
float orc_dot(const float *a, const float *b, int D)
{
    float acc[128];
    for (int j = 0; j < 128; ++j) acc[j] = 0.0f;
    int k = 0;
    for (; k + 128 <= D; k += 128)
        for (int j = 0; j < 128; ++j) acc[j] = fmaf(a[k + j], b[k + j], acc[j]);
    for (int j = 0; k + j < D; ++j) acc[j] = fmaf(a[k + j], b[k + j], acc[j]);
    float lane[32], nxt[32];
    for (int l = 0; l < 32; ++l)
        lane[l] = (acc[4 * l] + acc[4 * l + 1]) + (acc[4 * l + 2] + acc[4 * l + 3]);
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
        memcpy(lane, nxt, sizeof lane);
    }
    return lane[0];
}

/* pgvector's literal sequential loop (no FMA), for the tolerance cross-check. */
float orc_dot_sequential(const float *a, const float *b, int D)
{
    volatile float s = 0.0f;
    for (int k = 0; k < D; ++k) {
        volatile float p = a[k] * b[k];
        s = s + p;
    }
    return s;
}

/* similarity from the three fp32 sums, as SQL sees it: 1 - (a <=> b). */
double orc_sim_from_sums(float dot, float na, float nb)
{
    double sim = (double)dot / sqrt((double)na * (double)nb);
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = -1.0;
    double dist = 1.0 - sim; /* pgvector cosine_distance returns this float8 */
    return 1.0 - dist;       /* evaluate_alignments.py:97 / :128: "1 - (...)" */
}

double orc_cosine(const float *a, const float *b, int D)
{
    return orc_sim_from_sums(orc_dot(a, b, D), orc_dot(a, a, D), orc_dot(b, b, D));
}

double orc_cosine_sequential(const float *a, const float *b, int D)
{
    return orc_sim_from_sums(orc_dot_sequential(a, b, D), orc_dot_sequential(a, a, D),
                             orc_dot_sequential(b, b, D));
}

/* ------------------------------------------------------------------------ */
/* Weak-supervision terms                                                    */
/* ------------------------------------------------------------------------ */

/* src/insert_clip_embeddings.py:144-156.  hits = number of lexical terms that
 * occur in the chunk (bit t of the chunk's term set), T = len(lexical_components). */
double orc_lexical(int64_t hits, int64_t T)
{
    if (T <= 0) return 0.0;                 /* :146-147 */
    double denom = (double)T * 0.1;         /* :154 len(...) * 0.1 */
    if (!(denom > 1.0)) denom = 1.0;        /* max(..., 1) */
    double s = (double)hits / denom;
    return s < 1.0 ? s : 1.0;               /* min(1.0, ...) */
}

/* src/insert_clip_embeddings.py:159-210.  A missing / wrong-length bbox is
 * passed as all zeros, which the zero-width test (:172) maps to 0.0 exactly as
 * :161-169 would. */
double orc_positional(const double *ib, const double *cb)
{
    if ((ib[2] - ib[0] == 0.0) || (ib[3] - ib[1] == 0.0)) return 0.0; /* :172 */
    if ((cb[2] - cb[0] == 0.0) || (cb[3] - cb[1] == 0.0)) return 0.0; /* :174 */
    double x1 = ib[0] > cb[0] ? ib[0] : cb[0]; /* :178 max */
    double y1 = ib[1] > cb[1] ? ib[1] : cb[1];
    double x2 = ib[2] < cb[2] ? ib[2] : cb[2]; /* :180 min */
    double y2 = ib[3] < cb[3] ? ib[3] : cb[3];
    if (x2 <= x1 || y2 <= y1) {             /* :183 */
        double icx = (ib[0] + ib[2]) / 2, icy = (ib[1] + ib[3]) / 2;
        double ccx = (cb[0] + cb[2]) / 2, ccy = (cb[1] + cb[3]) / 2;
        double dx = icx - ccx, dy = icy - ccy;
        double dist = sqrt(dx * dx + dy * dy);      /* :191 */
        double s = 1.0 - (dist / 1000.0);           /* :196-197 */
        return s > 0.0 ? s : 0.0;
    }
    double w = x2 - x1, h = y2 - y1;
    double inter = (w > 0 ? w : 0) * (h > 0 ? h : 0); /* :201 */
    double ia = (ib[2] - ib[0]) * (ib[3] - ib[1]);
    double ca = (cb[2] - cb[0]) * (cb[3] - cb[1]);
    double uni = ia + ca - inter;
    if (uni == 0.0) return 0.0;             /* :206 */
    return inter / uni;
}

/*
 * src/insert_clip_embeddings.py:385-414 for one same-page pair.
 * out[0..2] = weak_score of the 'lexical' / 'positional' / 'combined' record,
 * 0.0 where the reference inserts no record.  Returns a bit mask of the
 * records present.
 */
int orc_weak_records(int use_lex, int use_pos, double lex, double pos, double *out)
{
    int have_lex = use_lex && lex > 0.05; /* :387 */
    int have_pos = use_pos && pos > 0.05; /* :393 */
    int mask = 0;
    out[0] = out[1] = out[2] = 0.0;
    if (use_lex && use_pos && have_lex && have_pos) { /* :398 len(scores) == 2 */
        double c = (lex + pos) / 2;
        if (c > 0.1) { out[2] = c; mask |= 4; }       /* :400 */
    } else {                                          /* :409-414 */
        if (have_lex) { out[0] = lex; mask |= 1; }
        if (have_pos) { out[1] = pos; mask |= 2; }
    }
    return mask;
}

static const int SCHEMA_LEX[4] = {0, 1, 0, 1}; /* :446-453 vanilla/lexical/positional/combined */
static const int SCHEMA_POS[4] = {0, 0, 1, 1};

static int64_t popcount_and(const uint64_t *a, const uint64_t *b, int words)
{
    int64_t h = 0;
    for (int w = 0; w < words; ++w) h += __builtin_popcountll(b ? (a[w] & b[w]) : a[w]);
    return h;
}

/* ------------------------------------------------------------------------ */
/* Pair enumeration: src/evaluate_alignments.py:48-69                        */
/* ------------------------------------------------------------------------ */

/* offsets[i]..offsets[i+1] index the same-page chunks of image i in increasing
 * chunk index.  Returns P.  pair_chunk may be NULL (count only). */
int64_t orc_pairs(const uint64_t *img_key, int64_t N, const uint64_t *chk_key, int64_t M,
                  int64_t *offsets, int64_t *pair_chunk)
{
    int64_t p = 0;
    for (int64_t i = 0; i < N; ++i) {
        offsets[i] = p;
        if (img_key[i] == ORC_NULL_KEY) continue;
        for (int64_t j = 0; j < M; ++j)
            if (chk_key[j] == img_key[i]) {
                if (pair_chunk) pair_chunk[p] = j;
                ++p;
            }
    }
    offsets[N] = p;
    return p;
}

/* ------------------------------------------------------------------------ */
/* Scoring + ranking: evaluate_alignments.py:109-143 (top-K), :169-231        */
/* ------------------------------------------------------------------------ */

typedef struct { double s; int64_t j; } cand_t;

static int cand_cmp(const void *x, const void *y)
{
    const cand_t *a = x, *b = y;
    if (a->s > b->s) return -1; /* ORDER BY similarity DESC (:132) */
    if (a->s < b->s) return 1;
    return (a->j > b->j) - (a->j < b->j); /* tie: lower chunk index first */
}

/*
 * One image row against every candidate chunk, for every requested schema.
 *   candidates: 0 = same manual+page only (the reference's SQL join, :128-131)
 *               1 = all chunks (north-star full N x M mode)
 *   score_s(i,j) = cos(i,j) + sum over the alignment records of (i,j) in schema s
 *                  of lam[type] * weak_score        (same-page pairs only)
 * Outputs (any may be NULL):
 *   topk_idx/topk_score [S][N][kmax]  (padding: -1 / -inf)
 *   pair_rank [S][P]    1-based rank of each true pair if <= cutoff, else 0
 *   pair_sim  [P]       exact cosine of each true pair
 * S = popcount(schema_mask), schemas in increasing bit order.
 * Rows are spread over `nthreads` pthreads (0 = all online cores); rows are
 * independent, so the result does not depend on the thread count.
 */
typedef struct {
    const float *img_emb; const uint64_t *img_key; const double *img_bbox; const uint64_t *img_terms;
    int64_t N;
    const float *chk_emb; const uint64_t *chk_key; const double *chk_bbox; const uint64_t *chk_terms;
    int64_t M;
    int D, term_words; int64_t T;
    uint32_t schema_mask; int candidates;
    double lam_lex, lam_pos, lam_comb;
    int kmax, cutoff;
    const int64_t *offsets, *pair_chunk;
    int64_t *topk_idx; double *topk_score; int32_t *pair_rank; double *pair_sim;
    const float *chk_n;
    int64_t next; /* atomic row-block counter */
} eval_job;

static void eval_row(const eval_job *J, int64_t i, cand_t *c, double *cosv)
{
    const int D = J->D;
    const int64_t N = J->N, M = J->M, P = J->offsets[N];
    int sch[4], S = 0;
    for (int s = 0; s < 4; ++s) if (J->schema_mask & (1u << s)) sch[S++] = s;
    const float *a = J->img_emb + i * D;
    float na = orc_dot(a, a, D);
    const int64_t p0 = J->offsets[i], p1 = J->offsets[i + 1];
    const int64_t *pair_chunk = J->pair_chunk;
    int64_t nc;
    if (J->candidates) {
        for (int64_t j = 0; j < M; ++j)
            cosv[j] = orc_sim_from_sums(orc_dot(a, J->chk_emb + j * D, D), na, J->chk_n[j]);
        nc = M;
    } else {
        for (int64_t p = p0; p < p1; ++p) {
            int64_t j = pair_chunk[p];
            cosv[p - p0] = orc_sim_from_sums(orc_dot(a, J->chk_emb + j * D, D), na, J->chk_n[j]);
        }
        nc = p1 - p0;
    }
    if (J->pair_sim)
        for (int64_t p = p0; p < p1; ++p)
            J->pair_sim[p] = J->candidates ? cosv[pair_chunk[p]] : cosv[p - p0];
    for (int si = 0; si < S; ++si) {
        int s = sch[si];
        for (int64_t q = 0; q < nc; ++q) {
            c[q].j = J->candidates ? q : pair_chunk[p0 + q];
            c[q].s = cosv[q];
        }
        if (SCHEMA_LEX[s] || SCHEMA_POS[s]) {
            for (int64_t p = p0; p < p1; ++p) {
                int64_t j = pair_chunk[p];
                double lex = 0.0, pos = 0.0, rec[3];
                if (SCHEMA_LEX[s])
                    lex = orc_lexical(popcount_and(J->chk_terms + j * J->term_words,
                                                   J->img_terms ? J->img_terms + i * J->term_words : 0,
                                                   J->term_words), J->T);
                if (SCHEMA_POS[s]) pos = orc_positional(J->img_bbox + 4 * i, J->chk_bbox + 4 * j);
                orc_weak_records(SCHEMA_LEX[s], SCHEMA_POS[s], lex, pos, rec);
                double w = J->lam_lex * rec[0] + J->lam_pos * rec[1] + J->lam_comb * rec[2];
                int64_t q = J->candidates ? j : p - p0;
                c[q].s = c[q].s + w;
            }
        }
        qsort(c, (size_t)nc, sizeof(cand_t), cand_cmp);
        if (J->topk_idx)
            for (int r = 0; r < J->kmax; ++r) {
                size_t o = ((size_t)si * N + i) * J->kmax + r;
                J->topk_idx[o] = r < nc ? c[r].j : -1;
                J->topk_score[o] = r < nc ? c[r].s : -INFINITY;
            }
        if (J->pair_rank) {
            for (int64_t p = p0; p < p1; ++p) J->pair_rank[(size_t)si * P + p] = 0;
            int64_t lim = nc < J->cutoff ? nc : J->cutoff;
            for (int64_t r = 0; r < lim; ++r) {
                int64_t lo = p0, hi = p1; /* is c[r] a true pair? pair list is sorted by chunk */
                while (lo < hi) {
                    int64_t mid = (lo + hi) / 2;
                    if (pair_chunk[mid] < c[r].j) lo = mid + 1; else hi = mid;
                }
                if (lo < p1 && pair_chunk[lo] == c[r].j)
                    J->pair_rank[(size_t)si * P + lo] = (int32_t)(r + 1);
            }
        }
    }
}

static void *eval_worker(void *arg)
{
    eval_job *J = arg;
    size_t cap = (size_t)(J->M > 0 ? J->M : 1);
    cand_t *c = malloc(sizeof(cand_t) * cap);
    double *cosv = malloc(sizeof(double) * cap);
    for (;;) {
        int64_t i0 = __atomic_fetch_add(&J->next, 4, __ATOMIC_RELAXED);
        if (i0 >= J->N) break;
        for (int64_t i = i0; i < i0 + 4 && i < J->N; ++i) eval_row(J, i, c, cosv);
    }
    free(c);
    free(cosv);
    return 0;
}

int orc_eval(const float *img_emb, const uint64_t *img_key, const double *img_bbox,
             const uint64_t *img_terms, int64_t N,
             const float *chk_emb, const uint64_t *chk_key, const double *chk_bbox,
             const uint64_t *chk_terms, int64_t M,
             int D, int term_words, int64_t T,
             uint32_t schema_mask, int candidates,
             double lam_lex, double lam_pos, double lam_comb,
             int kmax, int cutoff,
             const int64_t *offsets, const int64_t *pair_chunk,
             int64_t *topk_idx, double *topk_score, int32_t *pair_rank, double *pair_sim,
             int nthreads)
{
    float *chk_n = malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
    for (int64_t j = 0; j < M; ++j) chk_n[j] = orc_dot(chk_emb + j * D, chk_emb + j * D, D);
    eval_job J = {img_emb, img_key, img_bbox, img_terms, N, chk_emb, chk_key, chk_bbox, chk_terms, M,
                  D, term_words, T, schema_mask & 15u, candidates, lam_lex, lam_pos, lam_comb,
                  kmax, cutoff, offsets, pair_chunk, topk_idx, topk_score, pair_rank, pair_sim,
                  chk_n, 0};
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > 256) nthreads = 256;
    if (nthreads < 1) nthreads = 1;
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&th[started], 0, eval_worker, &J) == 0) ++started;
    eval_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
    free(chk_n);
    return nthreads;
}

/*
 * Alignment records for one schema (the `alignments` table producer,
 * insert_clip_embeddings.py:369-414), one entry per same-page pair and type:
 * rec[p*3 + t] = weak_score or 0.0 where no record is inserted.
 */
int orc_alignments(const double *img_bbox, const uint64_t *img_terms, int64_t N,
                   const double *chk_bbox, const uint64_t *chk_terms,
                   int term_words, int64_t T, int schema,
                   const int64_t *offsets, const int64_t *pair_chunk, double *rec)
{
    for (int64_t i = 0; i < N; ++i)
        for (int64_t p = offsets[i]; p < offsets[i + 1]; ++p) {
            int64_t j = pair_chunk[p];
            double lex = 0.0, pos = 0.0;
            if (SCHEMA_LEX[schema])
                lex = orc_lexical(popcount_and(chk_terms + j * term_words,
                                               img_terms ? img_terms + i * term_words : 0,
                                               term_words), T);
            if (SCHEMA_POS[schema]) pos = orc_positional(img_bbox + 4 * i, chk_bbox + 4 * j);
            orc_weak_records(SCHEMA_LEX[schema], SCHEMA_POS[schema], lex, pos, rec + 3 * p);
        }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Lexical term sets: src/insert_clip_embeddings.py:149-150                  */
/*     chunk_text_lower = text_chunk["text"].lower()                         */
/*     ... for term in lexical_components if term in chunk_text_lower        */
/* ------------------------------------------------------------------------ */
/* text: the chunks' lower-cased UTF-8 texts, concatenated (text_off [m+1]); terms likewise (term_off [T+1]).
 * bits [m][term_words]: bit t of row j = term t is a substring of text j (Python's `in`: an empty term occurs in
 * every text; on valid UTF-8 a byte match is a code-point match).  Plain nested loops on purpose. */
int orc_term_bitsets(const uint8_t *text, const int64_t *text_off, int64_t m, const uint8_t *terms,
                     const int64_t *term_off, int T, int term_words, uint64_t *bits)
{
    if (term_words * 64 < T) return -1;
    memset(bits, 0, (size_t)m * term_words * sizeof(uint64_t));
    for (int64_t j = 0; j < m; ++j) {
        const uint8_t *s = text + text_off[j];
        const int64_t len = text_off[j + 1] - text_off[j];
        for (int t = 0; t < T; ++t) {
            const uint8_t *pat = terms + term_off[t];
            const int64_t tl = term_off[t + 1] - term_off[t];
            int found = tl == 0;
            for (int64_t p = 0; !found && p + tl <= len; ++p) found = memcmp(s + p, pat, (size_t)tl) == 0;
            if (found) bits[j * term_words + (t >> 6)] |= 1ull << (t & 63);
        }
    }
    return 0;
}
