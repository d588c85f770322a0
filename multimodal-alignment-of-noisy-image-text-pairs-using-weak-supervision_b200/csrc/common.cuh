// common.cuh -- shared device helpers of the alignment-scoring path (sm_100a only).
//
// Build flag contract: the whole library is compiled with -fmad=false, so the
// only fused multiply-adds are the explicit fmaf() calls of the canonical dot
// product; the fp64 weak-supervision terms round exactly like the reference's
// Python float arithmetic (src/insert_clip_embeddings.py:144-210).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <vector>
#include "../../include/mmalign.h"

// ---------------------------------------------------------------------------
// Checked build (make check -> libmmalign_check.so, run with MMALIGN_LIB=...): the kernels test their own indices
// and invariants -- list capacities, shared-memory slots, chunk columns, output positions -- and count violations
// per source file (first failing line kept).  compute-sanitizer is not available on the GPU pool this is developed
// on; this is its stand-in, run over the whole -m gpu suite (profiles/sanitizer/).  Compiled out of the release build.
// ---------------------------------------------------------------------------
#ifdef MMALIGN_CHECKED
#define MMA_CHECK_DECL static __device__ unsigned int g_chk_fail[2];
#define MMA_CHECK(cond) do { if (!(cond)) { if (atomicAdd(&g_chk_fail[0], 1u) == 0u) g_chk_fail[1] = __LINE__; } } while (0)
#define MMA_CHECK_READER(name) int name(unsigned int *out) { return (int)cudaMemcpyFromSymbol(out, g_chk_fail, sizeof(g_chk_fail)); }
#else
#define MMA_CHECK_DECL
#define MMA_CHECK(cond) do { } while (0)
#define MMA_CHECK_READER(name) int name(unsigned int *out) { out[0] = out[1] = 0u; return 0; }
#endif

namespace mma {

int check_read_fused(unsigned int *out);    // {violations, first failing line} of fused_tc.cu
int check_read_rescore(unsigned int *out);  // ... of rescore.cu
int check_read_prep(unsigned int *out);     // ... of prep.cu
int check_read_ingest(unsigned int *out);   // ... of ingest.cu

constexpr int kMaxSchemas = 4;
constexpr int kMaxK = 8;

// ---------------------------------------------------------------------------
// Shared state between the launchers (api.cu owns the buffers).
// ---------------------------------------------------------------------------
struct Side {                 // images or this rank's chunk shard
    int64_t n = 0;
    int D = 0;
    int term_words = 0;
    const float *emb = nullptr;       // [n][D] fp32 master rows
    const uint64_t *key = nullptr;    // [n] page keys
    const double *bbox = nullptr;     // [n][4]
    const uint64_t *terms = nullptr;  // [n][term_words] or null (= all terms)
    __nv_bfloat16 *emb_bf16 = nullptr; // [n][D] L2-normalised rows, bf16 (fused-kernel operand)
    float *norm2 = nullptr;           // [n] sum x^2 in the canonical order
    float *err = nullptr;             // [n] |normalised row - its bf16 rounding|_2
};

struct PairIndex {            // CSR of same-page chunks per image (evaluate_alignments.py:48-69)
    int64_t P = 0;
    int64_t *offsets = nullptr;      // [N+1] pair offsets, pairs ordered (image, chunk)
    int32_t *sorted_chunk = nullptr; // [M] local chunk indices sorted by (page key, index)
    int64_t *sp_start = nullptr;     // [N] first position in sorted_chunk of the image's page
    int64_t c_max = 0;               // largest number of same-page chunks of any image
};

struct RunParams {
    int S;                 // schemas ranked
    int schema[kMaxSchemas];
    int candidates;
    int n_k;
    int k_list[kMaxK];
    int kmax;              // max(k_list): width of the top-K lists
    int mrr_cutoff;
    int kneed;             // max(kmax, mrr_cutoff): depth the ranking must be exact to
    double lam_lex, lam_pos, lam_comb;
    int64_t n_terms;       // T = len(lexical_components)
    int64_t col_offset;    // global index of local chunk 0
    float eps_scale;       // certificate margin multiplier (>= 1 in production)
};

struct Outputs {           // device pointers (api.cu stages host outputs)
    int64_t *topk_idx;
    double *topk_score;
    int32_t *pair_rank;
    double *pair_sim;
    double *pair_score;    // [S][P] ranking score of each true pair (multi-GPU rank step)
    int64_t *deep_idx;     // [S][N][kneed] lists to the full exact depth (multi-GPU rank step)
    double *deep_score;
};

// candidate lists written by the fused kernel (fused_tc.cu), read by rescore.cu
struct CandLists {
    uint64_t *keys = nullptr;  // [n_lists][cap]  lo = column, hi = fp32 score bits
    float *tau = nullptr;      // [n_lists] list is complete above tau
    int32_t *count = nullptr;  // [n_lists]
    int cap = 0;               // entries per list
    int n_splits = 0;          // column splits per row block
    int64_t n_row_blocks = 0;
    int kprime = 0;            // depth the union of a row's lists is complete to
    int kprime_list = 0;       // entries each list keeps
    // imported lists (mmalign_rescore_slab): one list per source rank and slab row, GLOBAL chunk indices
    const uint64_t *imp_keys = nullptr;  // [imp_src][imp_rows][imp_stride]
    const int32_t *imp_count = nullptr;  // [imp_src][imp_rows], -1 = overflowed at the source
    const float *imp_tau = nullptr;      // [imp_src][imp_rows]
    int imp_src = 0, imp_stride = 0;
    int64_t imp_rows = 0;
};

struct RowRange {             // image rows a rescoring launch ranks (0, 0 = all); outputs are indexed relative to it
    int64_t row0 = 0, n_rows = 0;
    int64_t pair0 = 0, P_out = 0;  // window of the per-pair outputs: first pair, number of pairs
    int64_t o_row0 = 0, o_rows = 0; // window of the per-row outputs (0, 0 = the rows themselves)
};

// ---------------------------------------------------------------------------
// Ordered keys
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t k)
{
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}
// candidate-list entry: lo = chunk column, hi = fp32 score bits
__device__ __forceinline__ float cand_score(uint64_t k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t cand_col(uint64_t k) { return (uint32_t)k; }

// ---------------------------------------------------------------------------
// Canonical fp32 dot product (bit-identical to oracle/mmalign_oracle.c: orc_dot):
// lane l owns float4 chunks l, l+32, ... ; four fmaf chains per lane; folded
// (0+1)+(2+3); xor butterfly 16,8,4,2,1.  Rows must be 16-byte aligned, D % 4 == 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_dot(const float4 *__restrict__ a, const float4 *__restrict__ b,
                                          int d4, int lane)
{
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    // four 512-byte warp loads in flight per step; the fmaf chains still run in increasing k
    for (int c = lane; c < d4; c += 128) {
        float4 y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + 32 * u < d4) y[u] = __ldg(b + c + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + 32 * u < d4) {
                const float4 x = a[c + 32 * u];
                c0 = fmaf(x.x, y[u].x, c0);
                c1 = fmaf(x.y, y[u].y, c1);
                c2 = fmaf(x.z, y[u].z, c2);
                c3 = fmaf(x.w, y[u].w, c3);
            }
    }
    float s = (c0 + c1) + (c2 + c3);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xFFFFFFFFu, s, off);
    return s;
}

// Two independent canonical dot products with the loads of both in flight together (rescoring kernel:
// the gathers are latency-bound).  Each result is bit-identical to warp_dot().
__device__ __forceinline__ void warp_dot2(const float4 *__restrict__ a, const float4 *__restrict__ b0,
                                          const float4 *__restrict__ b1, int d4, int lane, float &r0, float &r1)
{
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    for (int c = lane; c < d4; c += 128) {
        float4 y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + 32 * u < d4) { y[u] = __ldg(b0 + c + 32 * u); z[u] = __ldg(b1 + c + 32 * u); }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + 32 * u < d4) {
                const float4 x = a[c + 32 * u];
                p0 = fmaf(x.x, y[u].x, p0); p1 = fmaf(x.y, y[u].y, p1);
                p2 = fmaf(x.z, y[u].z, p2); p3 = fmaf(x.w, y[u].w, p3);
                q0 = fmaf(x.x, z[u].x, q0); q1 = fmaf(x.y, z[u].y, q1);
                q2 = fmaf(x.z, z[u].z, q2); q3 = fmaf(x.w, z[u].w, q3);
            }
    }
    float s = (p0 + p1) + (p2 + p3), t = (q0 + q1) + (q2 + q3);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        s = s + __shfl_xor_sync(0xFFFFFFFFu, s, off);
        t = t + __shfl_xor_sync(0xFFFFFFFFu, t, off);
    }
    r0 = s; r1 = t;
}

// Four independent canonical dot products, the loads of all four in flight together (warp-per-row rescoring:
// 16 x 512 B per warp).  Each result is bit-identical to warp_dot().
__device__ __forceinline__ void warp_dot4(const float4 *__restrict__ a, const float4 *__restrict__ b0,
                                          const float4 *__restrict__ b1, const float4 *__restrict__ b2,
                                          const float4 *__restrict__ b3, int d4, int lane, float (&r)[4])
{
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[q][k] = 0.f;
    for (int c = lane; c < d4; c += 128) {
        float4 y[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + 32 * u < d4) {
                y[0][u] = __ldg(b0 + c + 32 * u); y[1][u] = __ldg(b1 + c + 32 * u);
                y[2][u] = __ldg(b2 + c + 32 * u); y[3][u] = __ldg(b3 + c + 32 * u);
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + 32 * u < d4) {
                const float4 x = a[c + 32 * u];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[q][0] = fmaf(x.x, y[q][u].x, acc[q][0]); acc[q][1] = fmaf(x.y, y[q][u].y, acc[q][1]);
                    acc[q][2] = fmaf(x.z, y[q][u].z, acc[q][2]); acc[q][3] = fmaf(x.w, y[q][u].w, acc[q][3]);
                }
            }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = (acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = r[q] + __shfl_xor_sync(0xFFFFFFFFu, r[q], off);
}

// pgvector cosine as SQL sees it, 1 - (a <=> b): evaluate_alignments.py:97, :128
__device__ __forceinline__ double sim_from_sums(float dot, float na, float nb)
{
    double sim = (double)dot / sqrt((double)na * (double)nb);
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = -1.0;
    const double dist = 1.0 - sim;
    return 1.0 - dist;
}

// ---------------------------------------------------------------------------
// Weak-supervision terms in fp64
// ---------------------------------------------------------------------------
// src/insert_clip_embeddings.py:144-156
__device__ __forceinline__ double lexical_score(long long hits, long long T)
{
    if (T <= 0) return 0.0;
    double denom = (double)T * 0.1;
    if (!(denom > 1.0)) denom = 1.0;
    const double s = (double)hits / denom;
    return s < 1.0 ? s : 1.0;
}

// src/insert_clip_embeddings.py:159-210 (all-zero bbox = missing, caught by :172/:174)
__device__ __forceinline__ double positional_score(const double *ib, const double *cb)
{
    if ((ib[2] - ib[0] == 0.0) || (ib[3] - ib[1] == 0.0)) return 0.0;
    if ((cb[2] - cb[0] == 0.0) || (cb[3] - cb[1] == 0.0)) return 0.0;
    const double x1 = ib[0] > cb[0] ? ib[0] : cb[0];
    const double y1 = ib[1] > cb[1] ? ib[1] : cb[1];
    const double x2 = ib[2] < cb[2] ? ib[2] : cb[2];
    const double y2 = ib[3] < cb[3] ? ib[3] : cb[3];
    if (x2 <= x1 || y2 <= y1) {
        const double icx = (ib[0] + ib[2]) / 2, icy = (ib[1] + ib[3]) / 2;
        const double ccx = (cb[0] + cb[2]) / 2, ccy = (cb[1] + cb[3]) / 2;
        const double dx = icx - ccx, dy = icy - ccy;
        const double dist = sqrt(dx * dx + dy * dy);
        const double s = 1.0 - (dist / 1000.0);
        return s > 0.0 ? s : 0.0;
    }
    const double w = x2 - x1, h = y2 - y1;
    const double inter = (w > 0 ? w : 0) * (h > 0 ? h : 0);
    const double ia = (ib[2] - ib[0]) * (ib[3] - ib[1]);
    const double ca = (cb[2] - cb[0]) * (cb[3] - cb[1]);
    const double uni = ia + ca - inter;
    if (uni == 0.0) return 0.0;
    return inter / uni;
}

// src/insert_clip_embeddings.py:385-414; rec = {lexical, positional, combined}
__device__ __forceinline__ void weak_records(bool use_lex, bool use_pos, double lex, double pos,
                                             double *rec)
{
    const bool have_lex = use_lex && lex > 0.05;
    const bool have_pos = use_pos && pos > 0.05;
    rec[0] = rec[1] = rec[2] = 0.0;
    if (use_lex && use_pos && have_lex && have_pos) {
        const double c = (lex + pos) / 2;
        if (c > 0.1) rec[2] = c;
    } else {
        if (have_lex) rec[0] = lex;
        if (have_pos) rec[1] = pos;
    }
}

__device__ __forceinline__ bool schema_uses_lex(int s) { return s == 1 || s == 3; }
__device__ __forceinline__ bool schema_uses_pos(int s) { return s == 2 || s == 3; }

__device__ __forceinline__ long long term_hits(const uint64_t *chunk_terms, const uint64_t *img_terms,
                                               int words)
{
    long long h = 0;
    for (int w = 0; w < words; ++w)
        h += __popcll(img_terms ? (chunk_terms[w] & img_terms[w]) : chunk_terms[w]);
    return h;
}

// ---------------------------------------------------------------------------
// Launchers (each returns the cudaError_t of its launch)
// ---------------------------------------------------------------------------
// prep.cu
int sm_count();  // SMs of the current device
cudaError_t launch_prep(const Side &side, int64_t row0, int64_t rows, int sm_count, cudaStream_t st);
cudaError_t reduce_max_float(const float *x, int64_t n, float *out, cudaStream_t st);
size_t pair_index_scratch_bytes(int64_t N, int64_t M);
cudaError_t build_pair_index(const Side &img, const Side &chk, PairIndex &px, void *scratch, size_t scratch_bytes,
                             cudaStream_t st);
// rescore.cu
cudaError_t launch_rescore(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                           const CandLists *lists, const float *eps_chunk_max, const Outputs &out,
                           int32_t *fail_rows, int32_t *fail_count, unsigned long long *fail_thr,
                           unsigned long long *cand_counter, int32_t *error_flag, const float *tau_global,
                           int32_t *cert_count, RowRange rows, cudaStream_t st, int64_t grid_limit = 0,
                           int32_t *big_rows = nullptr, int32_t *big_count = nullptr,  // scratch of the warp-per-row kernels: [rows], [1] zeroed,
                           void *k2_scratch = nullptr,                                  // and k2_scratch_bytes(rows)
                           long long *n_launches = nullptr);                            // kernels this call launched
size_t k2_scratch_bytes(int64_t rows);
// scratch of the two-stage exact scan: per failed row (slot) its threshold, and what stage 1 kept for it
constexpr int kScanSlots = 2048, kScanCap = 1024;
struct ScanScratch {
    unsigned long long *thr = nullptr;  // [N] written by the rescoring kernel next to fail_rows
    void *buf = nullptr;                // [kScanSlots][kScanCap] 16-byte entries
    int32_t *cnt = nullptr;             // [kScanSlots]
};
size_t scan_scratch_bytes();
constexpr int kListSlack = 8;  // a list compacted to K' entries may keep up to K' + kListSlack (fused_tc.cu)
cudaError_t launch_export_lists(const CandLists &L, int64_t N, int n_dest, int64_t slab_rows, int stride,
                                int64_t col_offset, uint64_t *keys, int32_t *count, float *tau, cudaStream_t st);
cudaError_t launch_row_tau(const CandLists &L, int64_t N, float *tau_row, cudaStream_t st);
cudaError_t launch_exact_scan(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                              const int32_t *rows, const int32_t *n_rows_dev, int64_t n_rows_host,
                              const Outputs &out, int32_t *error_flag, RowRange range, const ScanScratch *pre,
                              cudaStream_t st);
cudaError_t launch_alignments(const Side &img, const Side &chk, const PairIndex &px, int schema,
                              int64_t n_terms, bool raw, double *rec, cudaStream_t st);
cudaError_t launch_pair_chunk(const PairIndex &px, int64_t N, int64_t col_offset, int64_t *pair_chunk,
                              cudaStream_t st);
size_t metrics_scratch_bytes(int S, int n_k);
cudaError_t launch_reduce_metrics(const int32_t *pair_rank, const double *pair_sim, int S, int64_t P,
                                  const int32_t *k_list_dev, int n_k, int mrr_cutoff, int64_t *hits,
                                  double *rr_sum, double *sim_sum, void *scratch, cudaStream_t st);
cudaError_t launch_merge_topk(const int64_t *in_idx, const double *in_score, int G, int64_t n_lists,
                              int K, int64_t *out_idx, double *out_score, cudaStream_t st);
cudaError_t launch_count_beating(const int64_t *deep_idx, const double *deep_score, int64_t N, int S, int K,
                                 int64_t n_q, const int64_t *q_image, const int64_t *q_chunk,
                                 const double *q_score, int32_t *counts, cudaStream_t st);
// ingest.cu
struct TermTableHost {  // the lexical terms, indexed on the host for term_bitsets_kernel (T terms: tiny)
    std::vector<int32_t> off, bucket_start, bucket_term, group_term;
    std::vector<uint32_t> always, hash;  // hash: 2 words per slot
    int hash_bits = 6;
};
int build_term_table(const uint8_t *terms, const int64_t *term_off, int n_terms, int term_words, TermTableHost *out);
cudaError_t launch_term_bitsets(const uint8_t *text, const int64_t *text_off, int64_t m, const uint8_t *term_bytes,
                                const int32_t *term_off, const int32_t *bucket_start, const int32_t *bucket_term,
                                const uint32_t *hash, int hash_bits, const int32_t *group_term,
                                const uint32_t *always, int term_words, uint64_t *bits, cudaStream_t st);
int64_t copy_scan(const uint8_t *data, int64_t n_bytes, int n_cols, int64_t *field_off, int32_t *field_len, int64_t cap);
cudaError_t launch_copy_decode(const uint8_t *data, const int64_t *field_off, const int32_t *field_len, int64_t n, int n_cols,
                               int vec_col, int bbox_col, int page_col, int D, float *emb, double *bbox, int32_t *page,
                               uint8_t *page_null, int32_t *err, cudaStream_t st);
cudaError_t launch_widen_rows(const void *src, int dtype, int64_t count, float *dst, cudaStream_t st);
// fused_tc.cu
struct FusedPlan {
    int64_t n_row_blocks;
    int n_splits;
    int tiles_per_split;
    int64_t n_lists;
    int cap;               // list capacity (entries)
    int kprime;            // depth the union of a row's lists is complete to
    int kprime_list;       // entries each list keeps
    int grid;
    size_t smem_bytes;
    int stages;
    bool a_resident;
    int epi_sleep_ns;      // pause between the epilogue warps' polls of their accumulator barrier
    int compact_one;       // routine list compaction split over two tile gaps (2), one list per gap (1), all flagged lists at once (0)
    bool pairs;            // clusters of two CTAs (tensor map B carries 128-row boxes): tcgen05.mma.cta_group::2, or
    bool mc;               // ... cta_group::1 MMAs per CTA over a B ring the two CTAs fill together by multicast
};
// cluster: 0 = one CTA per SM on its own, 1 = CTA pairs (cta_group::2), 2 = CTA pairs sharing B by multicast
int fused_plan(int64_t N, int64_t M, int D, int kneed, int kprime_req, int sm_count, int n_ranks, FusedPlan *plan,
               int cluster = 0);
cudaError_t launch_fused(const Side &img, const Side &chk, const FusedPlan &plan, const void *tmap_a,
                         const void *tmap_b, CandLists &lists, float *dump, cudaStream_t st, int64_t col_base = 0);
int encode_tensor_map(void *tmap_out, const void *base, int64_t rows, int D, int box_rows,
                      char *err, size_t errlen);

} // namespace mma
