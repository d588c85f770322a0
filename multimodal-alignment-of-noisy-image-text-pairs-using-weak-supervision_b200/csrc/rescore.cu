// rescore.cu -- exact scoring, ranking and metric kernels (SIMT, HBM/L2-bound).
//
//  rescore_kernel     K2: one CTA per image row.  Takes the row's candidate lists
//                     from the fused tcgen05 kernel (complete above a threshold tau),
//                     adds every same-page chunk (the true pairs), re-scores all of
//                     them exactly (canonical fp32 dot -> pgvector cosine in fp64,
//                     fp64 weak-supervision terms), orders them per schema by
//                     (score desc, index asc), writes top-K lists, true-pair ranks and
//                     similarities, and CERTIFIES the row: the kneed-th exact score
//                     must exceed tau + eps, eps a rigorous bound on |bf16 score -
//                     exact cosine|.  Uncertified rows go to the exact scan.
//  exact_scan_kernel  the exact scan: every column of the row in fp32, streaming
//                     top-kneed in shared memory.  Slow, exact, needs no certificate.
//  In MMALIGN_CAND_SAME_PAGE mode (the reference's join) only the same-page chunks
//  are candidates and rescore_kernel alone is the whole computation.
//
// Reference sites: src/evaluate_alignments.py:72-143 (scores, top-K), :169-231
// (metrics); src/insert_clip_embeddings.py:144-210, :369-414 (weak terms).
#include "common.cuh"
#include <math_constants.h>

namespace mma {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kEntCap = 1024;  // candidates + same-page entries handled per row
constexpr int kSpCap = 512;    // same-page chunks per image
constexpr int kScanRound = kWarps * 16;

struct SortEnt {
    double s;
    int32_t j;  // local chunk index
    int32_t e;  // entry slot
};

__device__ __forceinline__ bool ent_before(const SortEnt &a, const SortEnt &b)
{
    return a.s > b.s || (a.s == b.s && a.j < b.j);  // ORDER BY similarity DESC, lower index first
}

struct RowSmem {
    float *a;        // [D]
    int32_t *cols;   // [kEntCap]
    double *cosv;    // [kEntCap]
    double *weak;    // [S][kSpCap]
    SortEnt *buf;    // [kEntCap]
};

__host__ __device__ inline size_t row_smem_bytes(int D)
{
    return (size_t)kEntCap * sizeof(SortEnt) + (size_t)kEntCap * sizeof(double) +
           (size_t)kMaxSchemas * kSpCap * sizeof(double) + (size_t)kEntCap * sizeof(int32_t) +
           (size_t)D * sizeof(float);
}

__device__ __forceinline__ RowSmem carve(unsigned char *base, int D)
{
    RowSmem r;
    r.buf = reinterpret_cast<SortEnt *>(base);
    base += (size_t)kEntCap * sizeof(SortEnt);
    r.cosv = reinterpret_cast<double *>(base);
    base += (size_t)kEntCap * sizeof(double);
    r.weak = reinterpret_cast<double *>(base);
    base += (size_t)kMaxSchemas * kSpCap * sizeof(double);
    r.cols = reinterpret_cast<int32_t *>(base);
    base += (size_t)kEntCap * sizeof(int32_t);
    r.a = reinterpret_cast<float *>(base);
    (void)D;
    return r;
}

struct RowArgs {
    const float *img_emb; const uint64_t *img_key; const double *img_bbox; const uint64_t *img_terms;
    const float *img_n2; const float *img_err;
    const float *chk_emb; const uint64_t *chk_key; const double *chk_bbox; const uint64_t *chk_terms;
    const float *chk_n2;
    int64_t N, M;
    int D, term_words;
    const int64_t *offsets; const int32_t *sorted_chunk; const int64_t *sp_start;
    int64_t P;
    RunParams rp;
    Outputs out;
    int32_t *error_flag;  // set to 1 when a capacity limit is hit
};

__device__ void block_bitonic(SortEnt *buf, int n2)
{
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2; t += kThreads) {
                const int x = t ^ j;
                if (x > t) {
                    const SortEnt a = buf[t], b = buf[x];
                    const bool first_half = (t & k) == 0;
                    if (first_half ? ent_before(b, a) : ent_before(a, b)) { buf[t] = b; buf[x] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n)
{
    int p = 32;
    while (p < n) p <<= 1;
    return p;
}

// Row i: entries [0, n_ca) are candidate columns (cols[]), never same-page; the
// same-page chunks are appended here.  Returns false when the row is not certified.
__device__ bool finish_row(const RowArgs &A, const RowSmem &sm, int64_t i, int n_ca, bool certify,
                           float tau, float eps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p0 = A.offsets[i];
    const int c = (int)(A.offsets[i + 1] - p0);
    const int n = n_ca + c;
    const int d4 = A.D >> 2;
    const RunParams &rp = A.rp;
    // same-page chunks, in increasing chunk index
    for (int p = threadIdx.x; p < c; p += kThreads) sm.cols[n_ca + p] = A.sorted_chunk[A.sp_start[i] + p];
    __syncthreads();
    // exact cosine of every entry
    const float na = A.img_n2[i];
    for (int e = warp; e < n; e += kWarps) {
        const int j = sm.cols[e];
        const float dot = warp_dot(reinterpret_cast<const float4 *>(sm.a),
                                   reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)j * A.D), d4, lane);
        if (lane == 0) sm.cosv[e] = sim_from_sums(dot, na, A.chk_n2[j]);
    }
    // weak-supervision bonus of the same-page entries, per schema
    for (int t = threadIdx.x; t < c * rp.S; t += kThreads) {
        const int p = t % c, si = t / c, s = rp.schema[si];
        const int j = sm.cols[n_ca + p];
        double w = 0.0;
        if (s != 0) {
            double lex = 0.0, pos = 0.0, rec[3];
            if (schema_uses_lex(s))
                lex = lexical_score(term_hits(A.chk_terms + (int64_t)j * A.term_words,
                                              A.img_terms ? A.img_terms + i * A.term_words : nullptr,
                                              A.term_words), rp.n_terms);
            if (schema_uses_pos(s)) pos = positional_score(A.img_bbox + 4 * i, A.chk_bbox + 4 * (int64_t)j);
            weak_records(schema_uses_lex(s), schema_uses_pos(s), lex, pos, rec);
            w = rp.lam_lex * rec[0] + rp.lam_pos * rec[1] + rp.lam_comb * rec[2];
        }
        sm.weak[si * kSpCap + p] = w;
    }
    __syncthreads();
    if (A.out.pair_sim)
        for (int p = threadIdx.x; p < c; p += kThreads) A.out.pair_sim[p0 + p] = sm.cosv[n_ca + p];
    const int n2 = next_pow2(n);
    bool ok = true;
    for (int si = 0; si < rp.S; ++si) {
        for (int e = threadIdx.x; e < n2; e += kThreads) {
            SortEnt x;
            if (e < n) {
                x.s = sm.cosv[e];
                if (e >= n_ca) x.s = x.s + sm.weak[si * kSpCap + (e - n_ca)];
                x.j = sm.cols[e];
                x.e = e;
            } else {
                x.s = -CUDART_INF;
                x.j = 0x7FFFFFFF;
                x.e = -1;
            }
            sm.buf[e] = x;
        }
        __syncthreads();
        if (A.out.pair_score)
            for (int p = threadIdx.x; p < c; p += kThreads)
                A.out.pair_score[(int64_t)si * A.P + p0 + p] = sm.buf[n_ca + p].s;
        __syncthreads();
        block_bitonic(sm.buf, n2);
        if (A.out.topk_idx)
            for (int r = threadIdx.x; r < rp.kmax; r += kThreads) {
                const int64_t o = ((int64_t)si * A.N + i) * rp.kmax + r;
                A.out.topk_idx[o] = r < n ? (int64_t)sm.buf[r].j + rp.col_offset : -1;
                A.out.topk_score[o] = r < n ? sm.buf[r].s : -CUDART_INF;
            }
        if (A.out.deep_idx)
            for (int r = threadIdx.x; r < rp.kneed; r += kThreads) {
                const int64_t o = ((int64_t)si * A.N + i) * rp.kneed + r;
                A.out.deep_idx[o] = r < n ? (int64_t)sm.buf[r].j + rp.col_offset : -1;
                A.out.deep_score[o] = r < n ? sm.buf[r].s : -CUDART_INF;
            }
        if (A.out.pair_rank) {
            const int lim = n < rp.kneed ? n : rp.kneed;
            for (int r = threadIdx.x; r < lim; r += kThreads) {
                const int e = sm.buf[r].e;
                if (e >= n_ca) A.out.pair_rank[(int64_t)si * A.P + p0 + (e - n_ca)] = r + 1;
            }
        }
        if (certify && tau > -CUDART_INF_F)
            ok = ok && (n >= rp.kneed) && (sm.buf[rp.kneed - 1].s > (double)tau + (double)eps);
        __syncthreads();
    }
    return ok;
}

__device__ __forceinline__ void stage_row(const RowArgs &A, const RowSmem &sm, int64_t i)
{
    const float4 *src = reinterpret_cast<const float4 *>(A.img_emb + i * A.D);
    float4 *dst = reinterpret_cast<float4 *>(sm.a);
    for (int c = threadIdx.x; c < (A.D >> 2); c += kThreads) dst[c] = src[c];
}

// ---------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
rescore_kernel(RowArgs A, CandLists L, bool use_lists, const float *eps_chunk_max,
               int32_t *fail_rows, int32_t *fail_count, unsigned long long *cand_counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RowSmem sm = carve(smem_raw, A.D);
    __shared__ int s_nca;
    __shared__ float s_tau;
    for (int64_t i = blockIdx.x; i < A.N; i += gridDim.x) {
        const int c = (int)(A.offsets[i + 1] - A.offsets[i]);
        if (threadIdx.x == 0) { s_nca = 0; s_tau = -CUDART_INF_F; }
        stage_row(A, sm, i);
        __syncthreads();
        bool overflow = c > kSpCap;
        bool ok = !overflow;
        if (!use_lists) {
            if (ok) {
                if (threadIdx.x == 0 && cand_counter) atomicAdd(cand_counter, (unsigned long long)c);
                finish_row(A, sm, i, 0, false, 0.f, 0.f);
            }
        } else if (ok) {
            // lists of this row: one per (column split, accumulator half)
            const int64_t rb = i >> 7;
            const int r = (int)(i & 127);
            const int n_l = L.n_splits * 2;
            if (threadIdx.x < 32) {
                float t = -CUDART_INF_F;
                for (int l = threadIdx.x; l < n_l; l += 32) {
                    const int64_t id = (((int64_t)(l >> 1) * L.n_row_blocks + rb) * 2 + (l & 1)) * 128 + r;
                    t = fmaxf(t, L.tau[id]);
                }
                for (int off = 16; off >= 1; off >>= 1) t = fmaxf(t, __shfl_xor_sync(0xFFFFFFFFu, t, off));
                if (threadIdx.x == 0) s_tau = t;
            }
            __syncthreads();
            const float tau_union = s_tau;  // the union of the lists is complete above this
            const uint64_t ik = A.img_key[i];
            const float eps = A.img_err[i] * 1.001f + eps_chunk_max[0] * 1.001f + (float)A.D * 2.4e-7f + 2e-6f;
            // attempt 0: the best K' of the union by approximate score; attempt 1: the whole union
            for (int attempt = 0; attempt < 2; ++attempt) {
                if (threadIdx.x == 0) { s_nca = 0; s_tau = tau_union; }
                __syncthreads();
                for (int l = 0; l < n_l; ++l) {
                    const int64_t id = (((int64_t)(l >> 1) * L.n_row_blocks + rb) * 2 + (l & 1)) * 128 + r;
                    const int cnt = L.count[id];
                    const uint64_t *keys = L.keys + id * L.cap;
                    for (int e = threadIdx.x; e < cnt; e += kThreads) {
                        const uint64_t k = keys[e];
                        const uint32_t col = cand_col(k);
                        const float sa = cand_score(k);
                        if (sa > tau_union && (ik == MMALIGN_NULL_KEY || A.chk_key[col] != ik)) {
                            const int pos = atomicAdd(&s_nca, 1);
                            if (pos < kEntCap) { SortEnt x; x.s = (double)sa; x.j = (int32_t)col; x.e = 0; sm.buf[pos] = x; }
                        }
                    }
                }
                __syncthreads();
                const int n_all = s_nca;
                if (n_all + c > kEntCap) { ok = false; break; }
                const bool truncate = attempt == 0 && n_all > L.kprime;
                if (truncate) {
                    const int n2 = next_pow2(n_all);
                    for (int e = n_all + threadIdx.x; e < n2; e += kThreads) {
                        SortEnt x; x.s = -CUDART_INF; x.j = 0x7FFFFFFF; x.e = -1;
                        sm.buf[e] = x;
                    }
                    __syncthreads();
                    block_bitonic(sm.buf, n2);
                    // the union stays complete above the last kept approximate score
                    if (threadIdx.x == 0) { s_nca = L.kprime; s_tau = (float)sm.buf[L.kprime - 1].s; }
                    __syncthreads();
                }
                const int n_ca = s_nca;
                for (int e = threadIdx.x; e < n_ca; e += kThreads) sm.cols[e] = sm.buf[e].j;
                __syncthreads();
                if (threadIdx.x == 0 && cand_counter) atomicAdd(cand_counter, (unsigned long long)(n_ca + c));
                ok = finish_row(A, sm, i, n_ca, true, s_tau, eps);
                if (ok || !truncate) break;
                // not certified at depth K': clear this row's ranks and retry with everything the lists hold
                if (A.out.pair_rank)
                    for (int t = threadIdx.x; t < c * A.rp.S; t += kThreads)
                        A.out.pair_rank[(int64_t)(t / c) * A.P + A.offsets[i] + (t % c)] = 0;
                __syncthreads();
            }
        }
        if (!ok && threadIdx.x == 0) {
            if (use_lists) fail_rows[atomicAdd(fail_count, 1)] = (int32_t)i;
            else atomicExch(A.error_flag, 1);  // same-page mode: page larger than kSpCap
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Exact scan
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
exact_scan_kernel(RowArgs A, const int32_t *rows, const int32_t *n_rows_dev, int64_t n_rows_host)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RowSmem sm = carve(smem_raw, A.D);
    __shared__ int s_cnt;
    __shared__ double s_thr;
    __shared__ int s_thr_j;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_rows = n_rows_dev ? (int64_t)*n_rows_dev : n_rows_host;
    const int d4 = A.D >> 2;
    const int kneed = A.rp.kneed;
    for (int64_t b = blockIdx.x; b < n_rows; b += gridDim.x) {
        const int64_t i = rows ? rows[b] : b;
        const int64_t p0 = A.offsets[i];
        const int c = (int)(A.offsets[i + 1] - p0);
        if (threadIdx.x == 0) { s_cnt = 0; s_thr = -CUDART_INF; s_thr_j = 0x7FFFFFFF; }
        stage_row(A, sm, i);
        // a previous (uncertified) pass may have written ranks for this row
        if (A.out.pair_rank)
            for (int t = threadIdx.x; t < c * A.rp.S; t += kThreads)
                A.out.pair_rank[(int64_t)(t / c) * A.P + p0 + (t % c)] = 0;
        __syncthreads();
        if (c > kSpCap) {
            if (threadIdx.x == 0) atomicExch(A.error_flag, 1);
            continue;
        }
        const float na = A.img_n2[i];
        const uint64_t ik = A.img_key[i];
        for (int64_t base = 0; base < A.M; base += kScanRound) {
            const double thr = s_thr;
            const int thr_j = s_thr_j;
            for (int q = 0; q < 16; ++q) {
                const int64_t j = base + warp * 16 + q;
                if (j >= A.M) break;
                if (ik != MMALIGN_NULL_KEY && A.chk_key[j] == ik) continue;  // enters as a same-page entry
                const float dot = warp_dot(reinterpret_cast<const float4 *>(sm.a),
                                           reinterpret_cast<const float4 *>(A.chk_emb + j * A.D), d4, lane);
                if (lane == 0) {
                    const double s = sim_from_sums(dot, na, A.chk_n2[j]);
                    if (s > thr || (s == thr && (int)j < thr_j)) {
                        const int pos = atomicAdd(&s_cnt, 1);
                        SortEnt x; x.s = s; x.j = (int32_t)j; x.e = 0;
                        sm.buf[pos] = x;
                    }
                }
            }
            __syncthreads();
            if (s_cnt > kEntCap - kScanRound) {  // uniform: shrink to the best kneed
                const int cnt = s_cnt;
                for (int e = cnt + threadIdx.x; e < kEntCap; e += kThreads) {
                    SortEnt x; x.s = -CUDART_INF; x.j = 0x7FFFFFFF; x.e = -1;
                    sm.buf[e] = x;
                }
                __syncthreads();
                block_bitonic(sm.buf, kEntCap);
                if (threadIdx.x == 0) {
                    s_cnt = cnt < kneed ? cnt : kneed;
                    if (cnt >= kneed) { s_thr = sm.buf[kneed - 1].s; s_thr_j = sm.buf[kneed - 1].j; }
                }
                __syncthreads();
            }
        }
        const int cnt = s_cnt;
        const int n2 = next_pow2(cnt);
        for (int e = cnt + threadIdx.x; e < n2; e += kThreads) {
            SortEnt x; x.s = -CUDART_INF; x.j = 0x7FFFFFFF; x.e = -1;
            sm.buf[e] = x;
        }
        __syncthreads();
        block_bitonic(sm.buf, n2);
        const int n_ca = cnt < kneed ? cnt : kneed;
        for (int e = threadIdx.x; e < n_ca; e += kThreads) sm.cols[e] = sm.buf[e].j;
        __syncthreads();
        finish_row(A, sm, i, n_ca, false, 0.f, 0.f);
        __syncthreads();
    }
}

static RowArgs make_args(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                         const Outputs &out, int32_t *error_flag)
{
    RowArgs A;
    A.img_emb = img.emb; A.img_key = img.key; A.img_bbox = img.bbox; A.img_terms = img.terms;
    A.img_n2 = img.norm2; A.img_err = img.err;
    A.chk_emb = chk.emb; A.chk_key = chk.key; A.chk_bbox = chk.bbox; A.chk_terms = chk.terms;
    A.chk_n2 = chk.norm2;
    A.N = img.n; A.M = chk.n; A.D = img.D; A.term_words = chk.term_words;
    A.offsets = px.offsets; A.sorted_chunk = px.sorted_chunk; A.sp_start = px.sp_start; A.P = px.P;
    A.rp = rp; A.out = out; A.error_flag = error_flag;
    return A;
}

cudaError_t launch_rescore(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                           const CandLists *lists, const float *eps_chunk_max, const Outputs &out,
                           int32_t *fail_rows, int32_t *fail_count, unsigned long long *cand_counter,
                           int32_t *error_flag, cudaStream_t st)
{
    if (img.n == 0) return cudaSuccess;
    const size_t smem = row_smem_bytes(img.D);
    cudaError_t e = cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const RowArgs A = make_args(img, chk, px, rp, out, error_flag);
    CandLists L = lists ? *lists : CandLists();
    int64_t grid = img.n < 148 * 16 ? img.n : 148 * 16;
    rescore_kernel<<<(unsigned)grid, kThreads, smem, st>>>(A, L, lists != nullptr, eps_chunk_max, fail_rows,
                                                           fail_count, cand_counter);
    return cudaGetLastError();
}

cudaError_t launch_exact_scan(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                              const int32_t *rows, const int32_t *n_rows_dev, int64_t n_rows_host,
                              const Outputs &out, int32_t *error_flag, cudaStream_t st)
{
    if (img.n == 0 || (!n_rows_dev && n_rows_host == 0)) return cudaSuccess;
    const size_t smem = row_smem_bytes(img.D);
    cudaError_t e = cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const RowArgs A = make_args(img, chk, px, rp, out, error_flag);
    int64_t grid = 148 * 4;
    if (!n_rows_dev && n_rows_host < grid) grid = n_rows_host;
    exact_scan_kernel<<<(unsigned)grid, kThreads, smem, st>>>(A, rows, n_rows_dev, n_rows_host);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// `alignments` records: src/insert_clip_embeddings.py:369-414
// ---------------------------------------------------------------------------
__global__ void alignments_kernel(RowArgs A, int schema, int64_t n_terms, bool raw, double *rec)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < A.P;
         p += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = A.N;  // image of pair p: last i with offsets[i] <= p
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (A.offsets[mid] <= p) lo = mid; else hi = mid;
        }
        const int64_t i = lo;
        const int64_t j = A.sorted_chunk[A.sp_start[i] + (p - A.offsets[i])];
        double lex = 0.0, pos = 0.0;
        if (schema_uses_lex(schema))
            lex = lexical_score(term_hits(A.chk_terms + j * A.term_words,
                                          A.img_terms ? A.img_terms + i * A.term_words : nullptr,
                                          A.term_words), n_terms);
        if (schema_uses_pos(schema)) pos = positional_score(A.img_bbox + 4 * i, A.chk_bbox + 4 * j);
        if (raw) { rec[3 * p] = lex; rec[3 * p + 1] = pos; rec[3 * p + 2] = 0.0; }
        else weak_records(schema_uses_lex(schema), schema_uses_pos(schema), lex, pos, rec + 3 * p);
    }
}

cudaError_t launch_alignments(const Side &img, const Side &chk, const PairIndex &px, int schema,
                              int64_t n_terms, bool raw, double *rec, cudaStream_t st)
{
    if (px.P == 0) return cudaSuccess;
    RunParams rp = {};
    Outputs out = {};
    const RowArgs A = make_args(img, chk, px, rp, out, nullptr);
    int64_t grid = (px.P + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    alignments_kernel<<<(unsigned)grid, 256, 0, st>>>(A, schema, n_terms, raw, rec);
    return cudaGetLastError();
}

__global__ void pair_chunk_kernel(const int64_t *offsets, const int32_t *sorted_chunk,
                                  const int64_t *sp_start, int64_t N, int64_t col_offset, int64_t *pair_chunk)
{
    // one warp per image
    const int lane = threadIdx.x & 31;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < N;
         i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t p0 = offsets[i], c = offsets[i + 1] - p0, s0 = sp_start[i];
        for (int64_t p = lane; p < c; p += 32) pair_chunk[p0 + p] = (int64_t)sorted_chunk[s0 + p] + col_offset;
    }
}

cudaError_t launch_pair_chunk(const PairIndex &px, int64_t N, int64_t col_offset, int64_t *pair_chunk,
                              cudaStream_t st)
{
    if (N == 0 || px.P == 0) return cudaSuccess;
    int64_t grid = (N * 32 + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    pair_chunk_kernel<<<(unsigned)grid, 256, 0, st>>>(px.offsets, px.sorted_chunk, px.sp_start, N, col_offset,
                                                      pair_chunk);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K4: metric sums, fixed summation order (deterministic run to run)
//   hits: src/evaluate_alignments.py:182-192; rr: :203-216; sim: :226-231
// ---------------------------------------------------------------------------
constexpr int kRedBlocks = 256;

__device__ __forceinline__ double block_sum(double v, double *sm)
{
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    return t;  // valid in thread 0
}

// partial layout per block: [S][n_k] hits (as double-exact int64 stored separately), [S] rr, [1] sim
__global__ void __launch_bounds__(256)
metrics_partial_kernel(const int32_t *__restrict__ pair_rank, const double *__restrict__ pair_sim, int S,
                       int64_t P, const int32_t *__restrict__ k_list, int n_k, int mrr_cutoff,
                       long long *part_hits, double *part_rr, double *part_sim)
{
    __shared__ double sm[8];
    __shared__ long long smi[8];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int s = 0; s < S; ++s) {
        long long h[kMaxK];
        for (int q = 0; q < kMaxK; ++q) h[q] = 0;
        double rr = 0.0;
        for (int64_t p = t0; p < P; p += stride) {
            const int r = pair_rank[(int64_t)s * P + p];
            if (r >= 1) {
                for (int q = 0; q < n_k; ++q) h[q] += (r <= k_list[q]);
                if (r <= mrr_cutoff) rr += 1.0 / (double)r;
            }
        }
        const double t = block_sum(rr, sm);
        if (threadIdx.x == 0) part_rr[blockIdx.x * S + s] = t;
        for (int q = 0; q < n_k; ++q) {
            long long v = h[q];
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) smi[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                long long tt = 0;
                for (int w = 0; w < 8; ++w) tt += smi[w];
                part_hits[((int64_t)blockIdx.x * S + s) * n_k + q] = tt;
            }
        }
    }
    double sim = 0.0;
    if (pair_sim)
        for (int64_t p = t0; p < P; p += stride) sim += pair_sim[p];
    const double t = block_sum(sim, sm);
    if (threadIdx.x == 0) part_sim[blockIdx.x] = t;
}

__global__ void metrics_final_kernel(int S, int n_k, const long long *part_hits, const double *part_rr,
                                     const double *part_sim, int64_t *hits, double *rr_sum, double *sim_sum)
{
    const int t = threadIdx.x;
    if (t < S * n_k && hits) {
        long long v = 0;
        for (int b = 0; b < kRedBlocks; ++b) v += part_hits[(int64_t)b * S * n_k + t];
        hits[t] = v;
    }
    if (t < S && rr_sum) {
        double v = 0.0;
        for (int b = 0; b < kRedBlocks; ++b) v += part_rr[b * S + t];
        rr_sum[t] = v;
    }
    if (t == 0 && sim_sum) {
        double v = 0.0;
        for (int b = 0; b < kRedBlocks; ++b) v += part_sim[b];
        sim_sum[0] = v;
    }
}

size_t metrics_scratch_bytes(int S, int n_k)
{
    return (size_t)kRedBlocks * ((size_t)S * n_k * sizeof(long long) + (size_t)S * sizeof(double) + sizeof(double));
}

cudaError_t launch_reduce_metrics(const int32_t *pair_rank, const double *pair_sim, int S, int64_t P,
                                  const int32_t *k_list_dev, int n_k, int mrr_cutoff, int64_t *hits,
                                  double *rr_sum, double *sim_sum, void *scratch, cudaStream_t st)
{
    long long *ph = reinterpret_cast<long long *>(scratch);
    double *prr = reinterpret_cast<double *>(ph + (size_t)kRedBlocks * S * n_k);
    double *psim = prr + (size_t)kRedBlocks * S;
    metrics_partial_kernel<<<kRedBlocks, 256, 0, st>>>(pair_rank, pair_sim, S, P, k_list_dev, n_k, mrr_cutoff,
                                                       ph, prr, psim);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    metrics_final_kernel<<<1, 64, 0, st>>>(S, n_k, ph, prr, psim, hits, rr_sum, sim_sum);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K3: cross-rank merge of sorted top-K lists (after the NCCL all-gather)
// ---------------------------------------------------------------------------
__global__ void merge_topk_kernel(const int64_t *__restrict__ in_idx, const double *__restrict__ in_score,
                                  int G, int64_t n_lists, int K, int64_t *__restrict__ out_idx,
                                  double *__restrict__ out_score)
{
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < n_lists;
         l += (int64_t)gridDim.x * blockDim.x) {
        int head[16];
        for (int g = 0; g < G; ++g) head[g] = 0;
        for (int r = 0; r < K; ++r) {
            int best = -1;
            double bs = 0.0;
            int64_t bj = 0;
            for (int g = 0; g < G; ++g) {
                if (head[g] >= K) continue;
                const int64_t o = ((int64_t)g * n_lists + l) * K + head[g];
                const int64_t j = in_idx[o];
                if (j < 0) { head[g] = K; continue; }
                const double s = in_score[o];
                if (best < 0 || s > bs || (s == bs && j < bj)) { best = g; bs = s; bj = j; }
            }
            if (best >= 0) {
                out_idx[l * K + r] = bj;
                out_score[l * K + r] = bs;
                head[best]++;
            } else {
                out_idx[l * K + r] = -1;
                out_score[l * K + r] = -CUDART_INF;
            }
        }
    }
}

cudaError_t launch_merge_topk(const int64_t *in_idx, const double *in_score, int G, int64_t n_lists, int K,
                              int64_t *out_idx, double *out_score, cudaStream_t st)
{
    if (n_lists == 0) return cudaSuccess;
    int64_t grid = (n_lists + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
    merge_topk_kernel<<<(unsigned)grid, 128, 0, st>>>(in_idx, in_score, G, n_lists, K, out_idx, out_score);
    return cudaGetLastError();
}

// number of entries of this rank's deep lists that beat each (image, chunk, score) query
__global__ void count_beating_kernel(const int64_t *__restrict__ deep_idx, const double *__restrict__ deep_score,
                                     int64_t N, int S, int K, int64_t n_q, const int64_t *__restrict__ q_image,
                                     const int64_t *__restrict__ q_chunk, const double *__restrict__ q_score,
                                     int32_t *__restrict__ counts)
{
    const int64_t total = n_q * S;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(t / n_q);
        const int64_t q = t % n_q;
        const int64_t i = q_image[q], j = q_chunk[q];
        const double sc = q_score[(int64_t)s * n_q + q];
        const int64_t base = ((int64_t)s * N + i) * K;
        int lo = 0, hi = K;  // first position that does NOT beat the query (lists are sorted)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int64_t dj = deep_idx[base + mid];
            const double ds = deep_score[base + mid];
            const bool beats = dj >= 0 && (ds > sc || (ds == sc && dj < j));
            if (beats) lo = mid + 1; else hi = mid;
        }
        counts[(int64_t)s * n_q + q] = lo;
    }
}

cudaError_t launch_count_beating(const int64_t *deep_idx, const double *deep_score, int64_t N, int S, int K,
                                 int64_t n_q, const int64_t *q_image, const int64_t *q_chunk,
                                 const double *q_score, int32_t *counts, cudaStream_t st)
{
    if (n_q == 0) return cudaSuccess;
    int64_t grid = (n_q * S + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    count_beating_kernel<<<(unsigned)grid, 256, 0, st>>>(deep_idx, deep_score, N, S, K, n_q, q_image, q_chunk,
                                                         q_score, counts);
    return cudaGetLastError();
}

} // namespace mma
