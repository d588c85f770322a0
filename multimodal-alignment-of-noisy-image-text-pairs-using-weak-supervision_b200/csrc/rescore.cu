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
MMA_CHECK_DECL

#ifdef MMALIGN_PROFILE_EPI   // the profiling build (make prof): where the rescoring kernel's CTAs spend their cycles
__device__ unsigned long long g_k2_prof[16];
#define K2_T(var) const long long var = clock64()
#define K2_ADD(slot, expr) do { if (threadIdx.x == 0) atomicAdd(&g_k2_prof[slot], (unsigned long long)(expr)); } while (0)
extern "C" int mmalign_profile_k2(unsigned long long *out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, g_k2_prof, sizeof(g_k2_prof));
    if (e == cudaSuccess && reset) {
        unsigned long long z[16] = {};
        e = cudaMemcpyToSymbol(g_k2_prof, z, sizeof z);
    }
    return (int)e;
}
#else
#define K2_T(var)
#define K2_ADD(slot, expr)
#endif

constexpr int kThreads = 128;  // small CTAs: 8 rows in flight per SM hide the per-row chain of dependent loads
constexpr int kWarps = kThreads / 32;
constexpr int kEntCapMax = 1024;  // candidates + same-page entries handled per row (shared-memory sized per launch)
constexpr int kSpCapMax = 512;    // same-page chunks per image
constexpr int kScanRound = kWarps * 16;

// Sort key: score as an order-preserving 64-bit integer, then the chunk index (lower wins).
struct Key {
    unsigned long long k;
    int32_t j;  // local chunk index
    int32_t e;  // entry slot
};

__device__ __forceinline__ unsigned long long ord64(double s)
{
    const long long u = __double_as_longlong(s);
    return u < 0 ? ~(unsigned long long)u : ((unsigned long long)u | 0x8000000000000000ull);
}
__device__ __forceinline__ bool key_before(unsigned long long ka, int ja, unsigned long long kb, int jb)
{
    return ka > kb || (ka == kb && ja < jb);  // ORDER BY similarity DESC, lower index first
}
__device__ __forceinline__ Key key_pad() { Key x; x.k = 0ull; x.j = 0x7FFFFFFF; x.e = -1; return x; }

struct RowSmem {
    Key *buf;                   // [A.ent_cap]
    double *cosv;               // [A.ent_cap] exact cosine of every entry
    double *sp_s;               // [A.sp_cap] ranking score of the same-page entries in the current schema
    unsigned long long *sp_k;   // [A.sp_cap] its order-preserving key
    double *sp_lex, *sp_pos;    // [A.sp_cap] raw lexical / positional score of the same-page entries
    int32_t *sp_cols;           // [A.sp_cap] the row's same-page chunks (local index, increasing)
    int32_t *cols;              // [A.ent_cap]
    float *approx;              // [A.ent_cap] approximate scores of the row's union of lists, sorted descending
    float *a;                   // [D]
};

__host__ __device__ inline size_t row_smem_bytes(int D, int ent_cap, int sp_cap)
{
    return (size_t)ent_cap * (sizeof(Key) + sizeof(double) + sizeof(int32_t) + sizeof(float)) +
           (size_t)sp_cap * (3 * sizeof(double) + sizeof(unsigned long long) + sizeof(int32_t)) + (size_t)D * sizeof(float);
}

__device__ __forceinline__ RowSmem carve(unsigned char *base, int ent_cap, int sp_cap)
{
    RowSmem r;
    r.buf = reinterpret_cast<Key *>(base);
    base += (size_t)ent_cap * sizeof(Key);
    r.cosv = reinterpret_cast<double *>(base);
    base += (size_t)ent_cap * sizeof(double);
    r.sp_s = reinterpret_cast<double *>(base);
    base += (size_t)sp_cap * sizeof(double);
    r.sp_k = reinterpret_cast<unsigned long long *>(base);
    base += (size_t)sp_cap * sizeof(unsigned long long);
    r.sp_lex = reinterpret_cast<double *>(base);
    base += (size_t)sp_cap * sizeof(double);
    r.sp_pos = reinterpret_cast<double *>(base);
    base += (size_t)sp_cap * sizeof(double);
    r.cols = reinterpret_cast<int32_t *>(base);
    base += (size_t)ent_cap * sizeof(int32_t);
    r.sp_cols = reinterpret_cast<int32_t *>(base);
    base += (size_t)sp_cap * sizeof(int32_t);
    r.approx = reinterpret_cast<float *>(base);
    base += (size_t)ent_cap * sizeof(float);
    r.a = reinterpret_cast<float *>(base);
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
    unsigned long long *viol_counter;  // rows whose re-scored candidates broke |exact - approximate| <= eps (may be null)
    int ent_cap, sp_cap;  // shared-memory capacities of this launch (>= 256 / >= 8)
    bool need_lex, need_pos;  // some requested schema uses the lexical / positional term
    // rows [row0, row0 + n_rows) are ranked; per-row outputs are indexed by (i - row0), per-pair outputs by
    // (pair - offsets[row0]) and sized P_out = offsets[row0 + n_rows] - offsets[row0]
    int64_t row0, n_rows, pair0, P_out;
    // window of the per-row outputs: indexed by (i - o_row0) with o_rows rows per schema.  Equal to (row0, n_rows)
    // unless a run is cut into slabs that share one set of output arrays (api.cu: the slab pipeline).
    int64_t o_row0, o_rows;
};

__device__ __forceinline__ int next_pow2(int n)
{
    int p = 32;
    while (p < n) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------
// Block-wide sorts.  Bitonic networks with the keys in registers: strides below 32 are warp shuffles, wider ones
// go through shared memory.  KPT = keys per thread: one (up to 128 keys; the network stops at the next power of
// two >= n) or two (up to 256 keys).  Every thread of the block must call them.
// ---------------------------------------------------------------------------
struct KeyOps {        // exact ranking keys: (score desc, chunk index asc)
    typedef Key T;
    static __device__ __forceinline__ Key pad() { return key_pad(); }
    static __device__ __forceinline__ bool before(const Key &a, const Key &b) { return key_before(a.k, a.j, b.k, b.j); }
    static __device__ __forceinline__ Key shfl(const Key &x, int j)
    {
        Key o;
        o.k = __shfl_xor_sync(0xFFFFFFFFu, x.k, j);
        o.j = __shfl_xor_sync(0xFFFFFFFFu, x.j, j);
        o.e = __shfl_xor_sync(0xFFFFFFFFu, x.e, j);
        return o;
    }
};
struct PackedOps {     // approximate scores: one 64-bit key, hi = ordered fp32 score, lo = ~column; descending
    typedef unsigned long long T;
    static __device__ __forceinline__ T pad() { return 0ull; }
    static __device__ __forceinline__ bool before(T a, T b) { return a > b; }
    static __device__ __forceinline__ T shfl(T x, int j) { return __shfl_xor_sync(0xFFFFFFFFu, x, j); }
};
__device__ __forceinline__ unsigned long long pack_approx(float score, uint32_t col)
{
    return ((unsigned long long)f32_ordered(score) << 32) | (unsigned long long)(0xFFFFFFFFu - col);
}
__device__ __forceinline__ float packed_score(unsigned long long k) { return f32_unordered((uint32_t)(k >> 32)); }
__device__ __forceinline__ int32_t packed_col(unsigned long long k) { return (int32_t)(0xFFFFFFFFu - (uint32_t)k); }

template <typename Ops>
__device__ void sort_regs1(typename Ops::T *buf, int n)  // n <= kThreads
{
    typedef typename Ops::T T;
    const int tid = threadIdx.x;
    int m = 32;
    while (m < n) m <<= 1;
    T me = tid < n ? buf[tid] : Ops::pad();
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            T o;
            if (j >= 32) {
                buf[tid] = me;
                __syncthreads();
                o = buf[tid ^ j];
                __syncthreads();
            } else {
                o = Ops::shfl(me, j);
            }
            const bool lower = (tid & j) == 0, up = (tid & k) == 0;
            const bool o_first = Ops::before(o, me);
            if ((lower == up) ? o_first : !o_first) me = o;
        }
    }
    buf[tid] = me;
    __syncthreads();
}

template <typename Ops>
__device__ void sort_regs2(typename Ops::T *buf, int n)  // n <= 2 * kThreads; keys tid and tid + kThreads live in registers
{
    typedef typename Ops::T T;
    const int tid = threadIdx.x;
    T me[2];
    me[0] = tid < n ? buf[tid] : Ops::pad();
    me[1] = tid + kThreads < n ? buf[tid + kThreads] : Ops::pad();
    __syncthreads();
    for (int k = 2; k <= 2 * kThreads; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            T o[2];
            if (j == kThreads) {
                o[0] = me[1]; o[1] = me[0];
            } else if (j >= 32) {
                buf[tid] = me[0]; buf[tid + kThreads] = me[1];
                __syncthreads();
                o[0] = buf[tid ^ j]; o[1] = buf[(tid ^ j) + kThreads];
                __syncthreads();
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r) o[r] = Ops::shfl(me[r], j);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = tid + r * kThreads;
                const bool lower = (i & j) == 0, up = (i & k) == 0;
                const bool o_first = Ops::before(o[r], me[r]);
                if ((lower == up) ? o_first : !o_first) me[r] = o[r];
            }
        }
    }
    buf[tid] = me[0]; buf[tid + kThreads] = me[1];
    __syncthreads();
}

// Sorts buf[0, n) best-first.  Up to 256 keys in registers; larger inputs (the exact scan's streaming buffer, very
// long unions) use the classic shared-memory network over the next power of two (buf must have room for it).
template <typename Ops>
__device__ void sort_block(typename Ops::T *buf, int n)
{
    typedef typename Ops::T T;
    const int tid = threadIdx.x;
    if (n <= kThreads) { sort_regs1<Ops>(buf, n); return; }
    if (n <= 2 * kThreads) { sort_regs2<Ops>(buf, n); return; }
    const int n2 = next_pow2(n);
    for (int e = n + tid; e < n2; e += kThreads) buf[e] = Ops::pad();
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < n2; t += kThreads) {
                const int x = t ^ j;
                if (x > t) {
                    const T a = buf[t], b = buf[x];
                    const bool first_half = (t & k) == 0;
                    const bool b_first = Ops::before(b, a);
                    if (first_half ? b_first : !b_first) { buf[t] = b; buf[x] = a; }
                }
            }
            __syncthreads();
        }
    }
}
__device__ void sort_keys(Key *buf, int n) { sort_block<KeyOps>(buf, n); }

// Entries of image row i in shared memory: [0, c) its same-page chunks (the true pairs, sm.sp_cols), then
// [c, c + n_ca) candidate columns, never same-page.
//
// score_same_page: raw weak-supervision terms and exact cosine of the same-page entries (schema-independent).
__device__ void score_same_page(const RowArgs &A, const RowSmem &sm, int64_t i, int c)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d4 = A.D >> 2;
    const RunParams &rp = A.rp;
    for (int p = threadIdx.x; p < c; p += kThreads) {
        const int j = sm.sp_cols[p];
        sm.cols[p] = j;
        double lex = 0.0, pos = 0.0;
        if (A.need_lex)
            lex = lexical_score(term_hits(A.chk_terms + (int64_t)j * A.term_words,
                                          A.img_terms ? A.img_terms + i * A.term_words : nullptr, A.term_words), rp.n_terms);
        if (A.need_pos) pos = positional_score(A.img_bbox + 4 * i, A.chk_bbox + 4 * (int64_t)j);
        sm.sp_lex[p] = lex;
        sm.sp_pos[p] = pos;
    }
    __syncthreads();
    for (int e = warp; e < c; e += 2 * kWarps) {
        const int e1 = e + kWarps;
        const int j0 = sm.cols[e], j1 = e1 < c ? sm.cols[e1] : j0;
        float d0, d1;
        warp_dot2(reinterpret_cast<const float4 *>(sm.a), reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)j0 * A.D),
                  reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)j1 * A.D), d4, lane, d0, d1);
        if (lane == 0) sm.cosv[e] = (double)d0;
        if (lane == 1 && e1 < c) sm.cosv[e1] = (double)d1;
    }
    __syncthreads();
    const float na = A.img_n2[i];
    const int64_t p0 = A.offsets[i] - A.pair0;
    for (int e = threadIdx.x; e < c; e += kThreads) {
        const double x = sim_from_sums((float)sm.cosv[e], na, A.chk_n2[sm.cols[e]]);
        sm.cosv[e] = x;
        if (A.out.pair_sim) A.out.pair_sim[p0 + e] = x;
    }
    __syncthreads();
}

// ranking score of a same-page entry in schema s: cosine + weighted alignment records
__device__ __forceinline__ double ranking_score(const RunParams &rp, int s, double cosine, double lex, double pos)
{
    double w = 0.0;
    if (s != 0) {
        double rec[3];
        weak_records(schema_uses_lex(s), schema_uses_pos(s), schema_uses_lex(s) ? lex : 0.0,
                     schema_uses_pos(s) ? pos : 0.0, rec);
        w = rp.lam_lex * rec[0] + rp.lam_pos * rec[1] + rp.lam_comb * rec[2];
    }
    return cosine + w;
}
__device__ __forceinline__ double same_page_score(const RunParams &rp, const RowSmem &sm, int s, int p)
{
    return ranking_score(rp, s, sm.cosv[p], sm.sp_lex[p], sm.sp_pos[p]);
}

// Number of values > x in approx[0, n), sorted descending.
__device__ __forceinline__ int count_above(const float *approx, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((double)approx[mid] > x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// finish_row: exact cosine of the n_ca candidates, one sort of the candidates by exact cosine (every schema ranks
// them alike), then per schema the same-page entries are merged by counting; top-K lists, true-pair ranks.
// score_same_page() must have run.
//
// Certificate (`bound` > -inf): every column that was NOT re-scored has exact cosine <= bound.  The row is
// certified when, in every schema, the kmax-th best score exceeds bound (the top-K list is final) and every true
// pair either scores above bound (its rank among the re-scored entries is its rank), or already has kneed
// entries ahead of it -- counting the un-re-scored candidates whose approximate score exceeds the pair's by more
// than eps (approx[n_ca, n_approx), sorted descending): its rank is beyond the cutoff either way.
// Returns false when the row is not certified.
__device__ bool finish_row(const RowArgs &A, const RowSmem &sm, int64_t i, int n_ca, double bound, double eps,
                           const float *approx, int n_approx, int32_t *cert_count = nullptr)
{
    __shared__ double s_kth;
    __shared__ int s_cert, s_bad, s_viol;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_viol = 0;  // (published by the barrier below)
    const int c = (int)(A.offsets[i + 1] - A.offsets[i]);
    const int64_t p0 = A.offsets[i] - A.pair0;  // position of the row's first pair in the per-pair outputs
    const int64_t io = i - A.o_row0;             // position of the row in the per-row outputs
    const int n = n_ca + c;
    const int d4 = A.D >> 2;
    const RunParams &rp = A.rp;
    const bool certify = bound > -CUDART_INF;
    // depth the lists must be final to: the top-K lists, or the full exact depth when the caller takes deep_idx
    const int kcert = A.out.deep_idx ? rp.kneed : rp.kmax;
    // exact cosine of the candidates
    K2_T(f0_);
    const float na = A.img_n2[i];
    for (int e = c + warp; e < n; e += 2 * kWarps) {
        const int e1 = e + kWarps;
        const int j0 = sm.cols[e], j1 = e1 < n ? sm.cols[e1] : j0;
        float d0, d1;
        warp_dot2(reinterpret_cast<const float4 *>(sm.a), reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)j0 * A.D),
                  reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)j1 * A.D), d4, lane, d0, d1);
        if (lane == 0) sm.cosv[e] = (double)d0;  // fp32 dot for now; the fp64 division runs one thread per entry below
        if (lane == 1 && e1 < n) sm.cosv[e1] = (double)d1;
    }
    __syncthreads();
    for (int e = c + threadIdx.x; e < n; e += kThreads) {
        const double x = sim_from_sums((float)sm.cosv[e], na, A.chk_n2[sm.cols[e]]);
        sm.cosv[e] = x;
        Key k; k.k = ord64(x); k.j = sm.cols[e]; k.e = e;
        sm.buf[e - c] = k;
        // The certificate rests on |exact - approximate| <= eps.  Every re-scored candidate tests that bound: a
        // violation (an input the error model does not cover) withdraws the row's certificate, and the row is
        // ranked by the exact scan instead.
        if (approx && certify && !(fabs(x - (double)approx[e - c]) <= eps)) s_viol = 1;
    }
    __syncthreads();
    K2_T(f1_);
    sort_keys(sm.buf, n_ca);
    K2_T(f2_);
    K2_ADD(5, f1_ - f0_);   // candidate gathers + fp64 cosine
    K2_ADD(6, f2_ - f1_);   // sort by exact score
    bool ok = !s_viol;
    if (s_viol && threadIdx.x == 0 && A.viol_counter) atomicAdd(A.viol_counter, 1ull);
    for (int si = 0; si < rp.S; ++si) {
        const int s = rp.schema[si];
        if (threadIdx.x == 0) { s_kth = -CUDART_INF; s_cert = 0; s_bad = 0; }
        // ranking score of the same-page entries in this schema
        for (int p = threadIdx.x; p < c; p += kThreads) {
            const double sc = same_page_score(rp, sm, s, p);
            sm.sp_s[p] = sc;
            sm.sp_k[p] = ord64(sc);
            if (A.out.pair_score) A.out.pair_score[(int64_t)si * A.P_out + p0 + p] = sc;
        }
        __syncthreads();
        const int64_t o_top = ((int64_t)si * A.o_rows + io) * rp.kmax, o_deep = ((int64_t)si * A.o_rows + io) * rp.kneed;
        // final position of an element = its position in its own sorted list + elements of the other list before it
        for (int t = threadIdx.x; t < c + min(n_ca, rp.kneed); t += kThreads) {
            unsigned long long k;
            int j, pos;
            double sc;
            if (t < c) {  // a same-page entry: binary search in the sorted candidates, count the other same-page entries
                k = sm.sp_k[t]; j = sm.cols[t]; sc = sm.sp_s[t];
                int lo = 0, hi = n_ca;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (key_before(sm.buf[mid].k, sm.buf[mid].j, k, j)) lo = mid + 1; else hi = mid;
                }
                pos = lo;
                for (int q = 0; q < c; ++q) pos += key_before(sm.sp_k[q], sm.cols[q], k, j);
                if (pos < rp.kneed) {
                    bool known = true;  // pos is the pair's rank ...
                    if (certify && !(sc > bound)) {  // ... unless a column that was not re-scored may beat it
                        const int ahead = approx ? count_above(approx, n_approx, sc + eps) - n_ca : 0;
                        known = false;
                        if (pos + (ahead > 0 ? ahead : 0) < rp.kneed) s_bad = 1;  // else: beyond the cutoff either way
                    }
                    if (known && A.out.pair_rank) A.out.pair_rank[(int64_t)si * A.P_out + p0 + t] = pos + 1;
                }
            } else {      // a candidate: its sorted position + the same-page entries that beat it
                const Key x = sm.buf[t - c];
                k = x.k; j = x.j; sc = sm.cosv[x.e];
                pos = t - c;
                for (int q = 0; q < c; ++q) pos += key_before(sm.sp_k[q], sm.cols[q], k, j);
            }
            MMA_CHECK(pos >= 0 && pos < n && n <= A.ent_cap && j >= 0 && j < A.M && io >= 0 && io < A.o_rows);
            if (pos < rp.kmax && A.out.topk_idx) {
                A.out.topk_idx[o_top + pos] = (int64_t)j + rp.col_offset;
                A.out.topk_score[o_top + pos] = sc;
            }
            if (pos < rp.kneed && A.out.deep_idx) {
                A.out.deep_idx[o_deep + pos] = (int64_t)j + rp.col_offset;
                A.out.deep_score[o_deep + pos] = sc;
            }
            if (pos == kcert - 1) s_kth = sc;
            if (cert_count && sc > bound) atomicAdd(&s_cert, 1);
        }
        for (int r = n + threadIdx.x; r < rp.kneed; r += kThreads) {  // fewer entries than the lists are wide
            if (r < rp.kmax && A.out.topk_idx) { A.out.topk_idx[o_top + r] = -1; A.out.topk_score[o_top + r] = -CUDART_INF; }
            if (A.out.deep_idx) { A.out.deep_idx[o_deep + r] = -1; A.out.deep_score[o_deep + r] = -CUDART_INF; }
        }
        __syncthreads();
        if (certify && !cert_count) ok = ok && !s_bad && (s_kth > bound);
        // fully sharded runs certify globally: this rank's entries that are provably above every column it left out
        // (counted among its best kneed + same-page entries, which is all the global test needs)
        if (cert_count && threadIdx.x == 0) cert_count[(int64_t)si * A.o_rows + io] = certify ? s_cert : rp.kneed;
        __syncthreads();
    }
    K2_T(f3_);
    K2_ADD(7, f3_ - f2_);   // per-schema merge, outputs
    return ok;
}

__device__ __forceinline__ void stage_row(const RowArgs &A, const RowSmem &sm, int64_t i)
{
    const float4 *src = reinterpret_cast<const float4 *>(A.img_emb + i * A.D);
    float4 *dst = reinterpret_cast<float4 *>(sm.a);
    for (int c = threadIdx.x; c < (A.D >> 2); c += kThreads) dst[c] = src[c];
    const int cnt = (int)(A.offsets[i + 1] - A.offsets[i]);
    if (cnt <= A.sp_cap) {
        const int64_t s0 = A.sp_start[i];
        for (int p = threadIdx.x; p < cnt; p += kThreads) sm.sp_cols[p] = A.sorted_chunk[s0 + p];
    }
}

// ---------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------
// List l of image row i.  Native layout (written by the fused kernel of this GPU): one list per (column split,
// accumulator half), local chunk indices.  Imported layout (mmalign_rescore_slab): one list per source rank,
// global chunk indices, cnt < 0 = the source could not fit the row into the exchange stride.
constexpr int kMaxListsPerRow = 128;
struct ListView { const uint64_t *keys; int cnt; float tau; };
__host__ __device__ __forceinline__ int lists_per_row(const CandLists &L) { return L.imp_keys ? L.imp_src : L.n_splits * 2; }
__device__ __forceinline__ ListView list_view(const CandLists &L, int64_t i, int64_t row0, int l)
{
    ListView v;
    if (L.imp_keys) {
        const int64_t id = (int64_t)l * L.imp_rows + (i - row0);
        v.keys = L.imp_keys + id * L.imp_stride; v.cnt = L.imp_count[id]; v.tau = L.imp_tau[id];
    } else {  // the fused kernel numbers its rows from the first row of its range
        const int64_t li = i - row0;
        const int64_t id = (((int64_t)(l >> 1) * L.n_row_blocks + (li >> 7)) * 2 + (l & 1)) * 128 + (li & 127);
        v.keys = L.keys + id * L.cap; v.cnt = L.count[id]; v.tau = L.tau[id];
    }
    return v;
}

// 64 registers -> 8 CTAs of 128 threads per SM (measured at config 5 with 256-thread CTAs: 110 ms at 64 registers against 128 / 161 ms at 80 / 124)
//
// Depth of the exact rescoring (lists mode).  The row's union of lists, sorted by approximate score a_(1) >= a_(2) ..,
// is complete above tau_union, and |exact - approximate| <= eps.  Exact scores are needed only for candidates that
// can change an output:
//   * the top-Kmax lists: everything with a >= a_(Kmax) - 2 eps (the Kmax best by approximate score have exact
//     cosine >= a_(Kmax) - eps, which whatever lies below that line cannot reach);
//   * the rank of a true pair with ranking score y: everything with a >= y - eps -- unless kneed candidates have
//     a > y + eps, in which case the pair is beyond the cutoff whatever their exact scores are.
// theta = the lowest of these lines (never below tau_union): ~40 candidates per row instead of K' ~ 190 when no
// true pair is near the cutoff.  finish_row() then proves the result with the row's certificate.
__global__ void __launch_bounds__(kThreads, 8)
rescore_kernel(RowArgs A, CandLists L, bool use_lists, const float *eps_chunk_max,
               int32_t *fail_rows, int32_t *fail_count, unsigned long long *fail_thr, unsigned long long *cand_counter,
               const float *tau_global, int32_t *cert_count, const int32_t *row_list, const int32_t *row_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RowSmem sm = carve(smem_raw, A.ent_cap, A.sp_cap);
    __shared__ int s_nca;
    __shared__ float s_tau;
    __shared__ unsigned long long s_theta;  // ord64 of the lowest line, shared minimum
    __shared__ int s_lcnt[kMaxListsPerRow];
    __shared__ const uint64_t *s_lkeys[kMaxListsPerRow];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const RunParams &rp = A.rp;
    unsigned long long *packed = reinterpret_cast<unsigned long long *>(sm.buf);  // the union, before it is re-scored
    // row_list: the rows the warp-per-row kernel handed over (wider than its registers), instead of a range
    const int64_t n_todo = row_list ? (int64_t)*row_count : A.n_rows;
    for (int64_t b = blockIdx.x; b < n_todo; b += gridDim.x) {
        const int64_t i = row_list ? (int64_t)row_list[b] : A.row0 + b;
        K2_T(r0_);
        const int c = (int)(A.offsets[i + 1] - A.offsets[i]);
        if (threadIdx.x == 0) { s_nca = 0; s_tau = -CUDART_INF_F; s_theta = ~0ull; }
        stage_row(A, sm, i);
        __syncthreads();
        K2_T(r1_);
        K2_ADD(0, r1_ - r0_);   // stage the query row
        bool ok = c <= A.sp_cap;
        unsigned long long thr = 0ull;  // lower bound of the row's kneed-th best exact cosine, for the exact scan
        if (!use_lists) {
            if (ok) {
                if (threadIdx.x == 0 && cand_counter) atomicAdd(cand_counter, (unsigned long long)c);
                score_same_page(A, sm, i, c);
                finish_row(A, sm, i, 0, -CUDART_INF, 0.0, nullptr, 0);
            }
        } else if (ok) {
            const int n_l = lists_per_row(L);
            if (threadIdx.x < 32) {
                float t = -CUDART_INF_F;
                for (int l = threadIdx.x; l < n_l; l += 32) {
                    const ListView v = list_view(L, i, A.row0, l);
                    t = fmaxf(t, v.cnt < 0 ? CUDART_INF_F : v.tau);  // an overflowed list certifies nothing
                    s_lcnt[l] = v.cnt;
                    s_lkeys[l] = v.keys;
                }
                for (int off = 16; off >= 1; off >>= 1) t = fmaxf(t, __shfl_xor_sync(0xFFFFFFFFu, t, off));
                if (threadIdx.x == 0) s_tau = t;
            }
            score_same_page(A, sm, i, c);  // (its barriers also publish s_tau / s_lcnt / s_lkeys)
            K2_T(r2_);
            K2_ADD(1, r2_ - r1_);   // list headers + same-page entries (weak terms, gathers, fp64 cosine)
            // the union of the lists is complete above this (fully sharded runs: the maximum over all ranks)
            const float tau_union = tau_global ? fmaxf(tau_global[i], s_tau) : s_tau;
            if (tau_union == CUDART_INF_F) ok = false;
            const uint64_t ik = A.img_key[i];
            const double eps = (double)rp.eps_scale *
                               (double)(A.img_err[i] * 1.001f + eps_chunk_max[0] * 1.001f + (float)A.D * 2.4e-7f + 2e-6f);
            if (ok) {
                // few long lists (this GPU's fused kernel): the block sweeps one list at a time; many short ones
                // (one per source rank): one list per warp, so their loads overlap
                const bool per_warp = n_l >= kWarps;
                const int l_step = per_warp ? kWarps : 1, e0 = per_warp ? lane : threadIdx.x, e_step = per_warp ? 32 : kThreads;
                for (int l = per_warp ? warp : 0; l < n_l; l += l_step) {
                    const int cnt = s_lcnt[l];
                    const uint64_t *keys = s_lkeys[l];
                    for (int e = e0; e < cnt; e += e_step) {
                        const uint64_t k = keys[e];
                        const uint32_t col = cand_col(k);
                        const float sa = cand_score(k);
                        MMA_CHECK((int64_t)col < A.M);
                        bool same_page = false;  // same-page chunks enter through the pair index, not through the lists
                        if (ik != MMALIGN_NULL_KEY) {
                            if (c <= 32) { for (int q = 0; q < c; ++q) same_page = same_page || sm.sp_cols[q] == (int32_t)col; }
                            else same_page = A.chk_key[col] == ik;
                        }
                        if (sa > tau_union && !same_page) {
                            const int pos = atomicAdd(&s_nca, 1);
                            if (pos < A.ent_cap) packed[pos] = pack_approx(sa, col);
                        }
                    }
                }
                __syncthreads();
                K2_T(r3_);
                K2_ADD(2, r3_ - r2_);   // sweep of the candidate lists
                const int n_all = s_nca;
                if (n_all + c > A.ent_cap) ok = false;
                if (ok) {
                    sort_block<PackedOps>(packed, n_all);  // by approximate score, lower column first
                    K2_T(r4_);
                    K2_ADD(3, r4_ - r3_);   // sort by approximate score
                    for (int e = threadIdx.x; e < n_all; e += kThreads) sm.approx[e] = packed_score(packed[e]);
                    __syncthreads();
                    if (n_all >= rp.kneed) thr = ord64((double)sm.approx[rp.kneed - 1] - eps);
                    int n_ca = n_all;
                    if (!tau_global && !A.out.deep_idx) {  // (deep lists are final to kneed entries: no pruning)
                        // lowest line that needs exact scores
                        if (threadIdx.x == 0 && n_all >= rp.kmax)
                            atomicMin(&s_theta, ord64((double)sm.approx[rp.kmax - 1] - 2.0 * eps));
                        for (int t = threadIdx.x; t < c * rp.S; t += kThreads) {
                            const double y = same_page_score(rp, sm, rp.schema[t / c], t % c);
                            if (count_above(sm.approx, n_all, y + eps) < rp.kneed) atomicMin(&s_theta, ord64(y - eps));
                        }
                        __syncthreads();
                        const unsigned long long theta = n_all >= rp.kmax ? s_theta : 0ull;
                        // candidates at or above theta
                        int lo = 0, hi = n_all;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (ord64((double)sm.approx[mid]) >= theta) lo = mid + 1; else hi = mid;
                        }
                        n_ca = lo;
                    }
                    // nothing that is not re-scored can have an exact cosine above this
                    const double bound = (double)(n_ca < n_all ? fmaxf(sm.approx[n_ca], tau_union) : tau_union) + eps;
                    for (int e = threadIdx.x; e < n_ca; e += kThreads) sm.cols[c + e] = packed_col(packed[e]);
                    __syncthreads();
                    K2_T(r5_);
                    K2_ADD(4, r5_ - r4_);   // depth of the exact rescoring (theta)
                    K2_ADD(8, 1);
                    if (threadIdx.x == 0 && cand_counter) atomicAdd(cand_counter, (unsigned long long)(n_ca + c));
                    if (tau_global) finish_row(A, sm, i, n_ca, bound, eps, nullptr, 0, cert_count);  // certified after the cross-rank count
                    else ok = finish_row(A, sm, i, n_ca, bound, eps, sm.approx, n_all);
                }
            }
        }
        if (!ok && threadIdx.x == 0) {
            if (use_lists && fail_rows) {
                const int slot = atomicAdd(fail_count, 1);
                fail_rows[slot] = (int32_t)i;
                if (fail_thr) fail_thr[slot] = thr;  // the exact scan only looks at columns that reach it (exact_prefilter_kernel)
            } else atomicExch(A.error_flag, 1);  // page larger than A.sp_cap (or, fully sharded, lists beyond capacity)
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// K2 in three kernels, one WARP per image row
// ---------------------------------------------------------------------------
// The same computation as rescore_kernel's lists mode -- same formulas, same order of floating-point operations,
// same outputs bit for bit -- cut where its character changes, for the common shape of a row (at most 32 lists, at
// most 32 same-page chunks, a union of at most 256 candidates; other rows go to rescore_kernel in row-list mode):
//   select_kernel  K2a  list headers, same-page entries (weak terms, exact cosine), sweep of the lists, sort by
//                       approximate score, depth of the exact rescoring.  Latency/issue-bound: many light warps.
//   gather_kernel  K2b  the row gathers and canonical dot products of the candidates, nothing else: two 2 KB rows
//                       in flight per warp, which is what tools/gather_probe.cu measures at 7.3 TB/s.
//   rank_kernel    K2c  fp64 cosine, sort by exact score, per-schema merge, outputs, certificate.
// In one kernel (block per row, and a first warp-per-row version) the gathers ran 40 % of each row's time and the
// row's serial phases the rest, with too few rows resident to cover either; apart they are each near their own limit.
// State between them travels through a per-row scratch record (K2Scratch, 3.9 KB per row).
// No block barriers; sorts in registers (8 keys per lane, blocked layout: strides 1-4 never leave the lane) with real
// loops over the network's stages -- a fully unrolled network (750 KB of code, 16 warps at 16 different places of it)
// was measured 6x slower than the instruction cache allows.
constexpr int kW2Warps = 4;     // warps (= rows in flight) per CTA
constexpr int kW2Cap = 256;     // union entries per row
constexpr int kW2Sp = 32;       // same-page chunks per row: one per lane
constexpr int kW2Keep = 208;    // what a union that outgrows kW2Cap is cut back to (>= the default depth K' = 192)

struct K2Row {                  // scratch header of a row
    int32_t n_all, n_ca;        // union entries above tau_union; re-scored candidates
    int32_t state;              // 0 = ranked by these kernels, 1 = left to rescore_kernel, 2 = uncertifiable (already on the fail list)
    float tau_union;
    unsigned long long thr;     // lower bound of the row's kneed-th best exact cosine, for the exact scan
    unsigned long long pad;
};
struct K2Scratch {
    int claim;                  // rows a warp claims at a time from the launch's row counters (1 for short launches: no stragglers)
    unsigned long long *next;   // [3] next unclaimed row of the select / gather / rank kernel (zeroed per launch)
    K2Row *hdr;                 // [rows]
    unsigned long long *pk;     // [rows][256] union, packed, sorted by approximate score (descending)
    float *dotv;                // [rows][256] fp32 dot products of the re-scored candidates, in pk order
    double *sp;                 // [rows][3][32] exact cosine, lexical, positional term of the same-page entries
};
size_t k2_scratch_bytes(int64_t rows)
{
    return (size_t)rows * (sizeof(K2Row) + kW2Cap * 8 + kW2Cap * 4 + 3 * kW2Sp * 8) + 1024;
}
static K2Scratch carve_scratch(void *base, int64_t rows)
{
    K2Scratch k;
    unsigned char *p = reinterpret_cast<unsigned char *>(base);
    k.pk = reinterpret_cast<unsigned long long *>(p); p += (size_t)rows * kW2Cap * 8;
    k.sp = reinterpret_cast<double *>(p); p += (size_t)rows * 3 * kW2Sp * 8;
    k.hdr = reinterpret_cast<K2Row *>(p); p += (size_t)rows * sizeof(K2Row);
    k.dotv = reinterpret_cast<float *>(p); p += (size_t)rows * kW2Cap * 4;
    k.next = reinterpret_cast<unsigned long long *>(p);
    return k;
}

// Bitonic sort of 32 * KPT keys held KPT per lane, element index = lane * KPT + r, best first.  Strides below KPT
// are compare-exchanges inside the lane, the others one shuffle per key.  `n` keys matter (the rest are pads that
// sort last): the network stops at the next power of two >= n.
struct WKey { unsigned long long k; int32_t j; };
__device__ __forceinline__ bool wbefore(unsigned long long a, unsigned long long b) { return a > b; }
__device__ __forceinline__ bool wbefore(const WKey &a, const WKey &b) { return key_before(a.k, a.j, b.k, b.j); }
__device__ __forceinline__ unsigned long long wshfl(unsigned long long x, int m) { return __shfl_xor_sync(0xFFFFFFFFu, x, m); }
__device__ __forceinline__ WKey wshfl(const WKey &x, int m)
{
    WKey o;
    o.k = __shfl_xor_sync(0xFFFFFFFFu, x.k, m);
    o.j = __shfl_xor_sync(0xFFFFFFFFu, x.j, m);
    return o;
}
template <int KPT, typename T>
__device__ __forceinline__ void warp_sort(T (&me)[KPT], int lane, int n)
{
    const int i0 = lane * KPT;
#pragma unroll 1
    for (int k = 2; (k >> 1) < n && k <= 32 * KPT; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j >= KPT; j >>= 1) {
            const int lj = j / KPT;
            const bool lower = (lane & lj) == 0;
            const bool want_first = lower == ((i0 & k) == 0);  // (k >= 2 * KPT here: the lane's keys share the direction)
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const T o = wshfl(me[r], lj);
                const bool o_first = wbefore(o, me[r]);
                if (want_first ? o_first : !o_first) me[r] = o;
            }
        }
#pragma unroll
        for (int j = KPT / 2; j >= 1; j >>= 1) {
            if (j < k) {
#pragma unroll
                for (int r = 0; r < KPT; ++r) {
                    const int x = r ^ j;
                    if (x > r) {
                        const bool up = ((i0 + r) & k) == 0;
                        const T a = me[r], b = me[x];
                        const bool b_first = wbefore(b, a);
                        if (up ? b_first : !b_first) { me[r] = b; me[x] = a; }
                    }
                }
            }
        }
    }
}

// number of values > x among the n approximate scores of pk (sorted descending)
__device__ __forceinline__ int count_above_packed(const unsigned long long *pk, int n, double x)
{
    if (n == 0 || (double)packed_score(pk[n - 1]) > x) return n;  // (the usual answer for a pair far below the cutoff)
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((double)packed_score(pk[mid]) > x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ double unord64(unsigned long long k)
{
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
}
__device__ __forceinline__ double row_eps(const RowArgs &A, int64_t i, float eps_chunk)
{
    return (double)A.rp.eps_scale * (double)(A.img_err[i] * 1.001f + eps_chunk * 1.001f + (float)A.D * 2.4e-7f + 2e-6f);
}

// ---- K2a
struct SelectSmem { float4 *a; unsigned long long *pk; double *sp; int32_t *spcol; };
__host__ __device__ inline size_t select_smem_bytes(int D) { return (size_t)D * 4 + kW2Cap * 8 + 3 * kW2Sp * 8 + kW2Sp * 4; }

__global__ void __launch_bounds__(kW2Warps * 32, 8)
select_kernel(RowArgs A, CandLists L, K2Scratch K, const float *__restrict__ eps_chunk_max, int32_t *fail_rows,
              int32_t *fail_count, unsigned long long *fail_thr, unsigned long long *cand_counter, int32_t *big_rows,
              int32_t *big_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    SelectSmem w;
    {
        unsigned char *base = smem_raw + (size_t)warp * select_smem_bytes(A.D);
        w.a = reinterpret_cast<float4 *>(base); base += (size_t)A.D * 4;
        w.pk = reinterpret_cast<unsigned long long *>(base); base += kW2Cap * 8;
        w.sp = reinterpret_cast<double *>(base); base += 3 * kW2Sp * 8;
        w.spcol = reinterpret_cast<int32_t *>(base);
    }
    const RunParams &rp = A.rp;
    const int n_l = lists_per_row(L);
    const int d4 = A.D >> 2;
    const float eps_chunk = eps_chunk_max[0];
    unsigned long long cand_total = 0ull;
    for (;;) {  // rows are claimed a few at a time: their cost varies (40 to 250 gathers), a static split leaves stragglers
        long long b0 = 0;
        if (lane == 0) b0 = (long long)atomicAdd(K.next + 0, (unsigned long long)K.claim);
        b0 = __shfl_sync(0xFFFFFFFFu, b0, 0);
        if (b0 >= A.n_rows) break;
        const int64_t b_end = b0 + K.claim < A.n_rows ? b0 + K.claim : A.n_rows;
    for (int64_t b = b0; b < b_end; ++b) {
        const int64_t i = A.row0 + b;
        K2_T(r0_);
        const int64_t off0 = A.offsets[i];
        const int c = (int)(A.offsets[i + 1] - off0);
        bool big = c > kW2Sp;
        // ---- the query row, its same-page chunks, its list headers: independent loads, issued together
        __syncwarp();
        if (!big) {
            const float4 *src = reinterpret_cast<const float4 *>(A.img_emb + i * A.D);
            for (int q = lane; q < d4; q += 32) w.a[q] = src[q];
        }
        int spj = -1;
        if (!big && lane < c) spj = A.sorted_chunk[A.sp_start[i] + lane];
        w.spcol[lane] = spj;
        ListView v;
        v.keys = nullptr; v.cnt = 0; v.tau = -CUDART_INF_F;
        if (lane < n_l) v = list_view(L, i, A.row0, lane);
        float tau_union = lane < n_l ? (v.cnt < 0 ? CUDART_INF_F : v.tau) : -CUDART_INF_F;  // an overflowed list certifies nothing
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) tau_union = fmaxf(tau_union, __shfl_xor_sync(FULL, tau_union, off));
        const float na = A.img_n2[i];
        const double eps = row_eps(A, i, eps_chunk);
        __syncwarp();
        K2_T(r1_);
        K2_ADD(0, r1_ - r0_);
        // ---- same-page entries: weak terms (lane p), exact cosine (two gathers per batch)
        if (!big) {
            double lex = 0.0, pos = 0.0, cosp = 0.0;
            float nb = 1.f;
            if (lane < c) {
                nb = A.chk_n2[spj];
                if (A.need_lex)
                    lex = lexical_score(term_hits(A.chk_terms + (int64_t)spj * A.term_words,
                                                  A.img_terms ? A.img_terms + i * A.term_words : nullptr, A.term_words), rp.n_terms);
                if (A.need_pos) pos = positional_score(A.img_bbox + 4 * i, A.chk_bbox + 4 * (int64_t)spj);
            }
            float spdot = 0.f;
            for (int e = 0; e < c; e += 2) {
                const int c0 = __shfl_sync(FULL, spj, e), c1 = __shfl_sync(FULL, spj, e + 1 < c ? e + 1 : e);
                float d0, d1;
                warp_dot2(w.a, reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)c0 * A.D),
                          reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)c1 * A.D), d4, lane, d0, d1);
                if (lane == e) spdot = d0;
                if (lane == e + 1) spdot = d1;
            }
            if (lane < c) {
                cosp = sim_from_sums(spdot, na, nb);
                if (A.out.pair_sim) A.out.pair_sim[off0 - A.pair0 + lane] = cosp;
            }
            w.sp[lane] = cosp; w.sp[kW2Sp + lane] = lex; w.sp[2 * kW2Sp + lane] = pos;
        }
        K2_T(r2_);
        K2_ADD(1, r2_ - r1_);
        const bool ok = tau_union != CUDART_INF_F;
        // ---- sweep of the lists: entries above tau_union that are not same-page (those enter through the pair index).
        // A union that outgrows the 256 slots (many lists per row: column groups, column splits of a remainder launch,
        // lists imported from other ranks) is cut back to its best kW2Keep entries and the completeness threshold
        // rises to the last kept score -- the union stays complete above it, which is all the certificate needs, as
        // long as the depth the lists were planned for (L.kprime) fits; deeper plans go to the block-per-row kernel.
        int n_all = 0;
        if (!big && ok) {
            const bool may_cut = L.kprime > 0 && L.kprime <= kW2Keep;
            for (int l = 0; l < n_l; ++l) {
                const int cnt = __shfl_sync(FULL, v.cnt, l);
                const uint64_t *keys = reinterpret_cast<const uint64_t *>(__shfl_sync(FULL, (unsigned long long)v.keys, l));
                // four batches of 32 entries in flight, one copy of the loop body
                uint64_t k0 = lane < cnt ? __ldg(keys + lane) : 0ull, k1 = lane + 32 < cnt ? __ldg(keys + lane + 32) : 0ull;
                uint64_t k2 = lane + 64 < cnt ? __ldg(keys + lane + 64) : 0ull, k3 = lane + 96 < cnt ? __ldg(keys + lane + 96) : 0ull;
#pragma unroll 1
                for (int e0 = 0; e0 < cnt; e0 += 32) {
                    const int e = e0 + lane;
                    const uint64_t k4 = e + 128 < cnt ? __ldg(keys + e + 128) : 0ull;
                    if (may_cut && n_all > kW2Cap - 32) {  // (uniform)
                        __syncwarp();
                        unsigned long long me[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) me[r] = lane * 8 + r < n_all ? w.pk[lane * 8 + r] : 0ull;
                        __syncwarp();
                        warp_sort<8>(me, lane, n_all);
#pragma unroll
                        for (int r = 0; r < 8; ++r) w.pk[lane * 8 + r] = me[r];
                        __syncwarp();
                        tau_union = fmaxf(tau_union, packed_score(w.pk[kW2Keep - 1]));
                        n_all = kW2Keep;
                    }
                    const uint32_t col = cand_col(k0);
                    const float sa = cand_score(k0);
                    MMA_CHECK(e >= cnt || (int64_t)col < A.M);  // a list entry names a chunk of the table
                    bool keep = e < cnt && sa > tau_union;
                    if (keep) for (int q = 0; q < c; ++q) keep = keep && w.spcol[q] != (int32_t)col;
                    const unsigned m = __ballot_sync(FULL, keep);
                    const int at = n_all + __popc(m & lt_mask);
                    if (keep && at < kW2Cap) w.pk[at] = pack_approx(sa, col);
                    n_all += __popc(m);
                    k0 = k1; k1 = k2; k2 = k3; k3 = k4;
                }
            }
            big = n_all > kW2Cap;
        }
        K2Row h;
        h.n_all = n_all; h.n_ca = 0; h.state = big ? 1 : (ok ? 0 : 2); h.tau_union = tau_union; h.thr = 0ull; h.pad = 0ull;
        if (big) {  // wider than these kernels' registers: the block-per-row kernel takes the row
            if (lane == 0) { big_rows[atomicAdd(big_count, 1)] = (int32_t)i; K.hdr[b] = h; }
            K2_ADD(9, 1);
            continue;
        }
        if (!ok) {  // nothing certifies this row: straight to the exact scan
            if (lane == 0) {
                const int slot = atomicAdd(fail_count, 1);
                fail_rows[slot] = (int32_t)i;
                if (fail_thr) fail_thr[slot] = 0ull;
                K.hdr[b] = h;
            }
            continue;
        }
        __syncwarp();
        K2_T(r3_);
        K2_ADD(2, r3_ - r2_);
        // ---- sort by approximate score (lower column first)
        {
            unsigned long long me[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) me[r] = lane * 8 + r < n_all ? w.pk[lane * 8 + r] : 0ull;
            __syncwarp();
            warp_sort<8>(me, lane, n_all);
            unsigned long long *dst = K.pk + b * kW2Cap + lane * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r) { w.pk[lane * 8 + r] = me[r]; }
#pragma unroll
            for (int r = 0; r < 8; r += 2) if (lane * 8 + r < n_all) *reinterpret_cast<ulonglong2 *>(dst + r) = make_ulonglong2(me[r], me[r + 1]);
            __syncwarp();
        }
        K2_T(r4_);
        K2_ADD(3, r4_ - r3_);
        if (n_all >= rp.kneed) h.thr = ord64((double)packed_score(w.pk[rp.kneed - 1]) - eps);
        // ---- depth of the exact rescoring (see rescore_kernel): the lowest line that needs exact scores
        unsigned long long theta = ~0ull;
        if (n_all >= rp.kmax) theta = ord64((double)packed_score(w.pk[rp.kmax - 1]) - 2.0 * eps);
        for (int t = lane; t < c * rp.S; t += 32) {
            const int p = t % c;
            const double y = ranking_score(rp, rp.schema[t / c], w.sp[p], w.sp[kW2Sp + p], w.sp[2 * kW2Sp + p]);
            if (count_above_packed(w.pk, n_all, y + eps) < rp.kneed) {
                const unsigned long long o = ord64(y - eps);
                theta = o < theta ? o : theta;
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, theta, off);
            theta = o < theta ? o : theta;
        }
        if (n_all < rp.kmax) theta = 0ull;
        {
            int lo = 0, hi = n_all;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ord64((double)packed_score(w.pk[mid])) >= theta) lo = mid + 1; else hi = mid;
            }
            h.n_ca = lo;
        }
        MMA_CHECK(h.n_ca >= 0 && h.n_ca <= n_all && n_all <= kW2Cap && c <= kW2Sp && b >= 0 && b < A.n_rows);
        if (lane == 0) K.hdr[b] = h;
        if (lane < c) {
            double *sp = K.sp + b * (3 * kW2Sp);
            sp[lane] = w.sp[lane]; sp[kW2Sp + lane] = w.sp[kW2Sp + lane]; sp[2 * kW2Sp + lane] = w.sp[2 * kW2Sp + lane];
        }
        cand_total += (unsigned long long)(h.n_ca + c);
        K2_T(r5_);
        K2_ADD(4, r5_ - r4_);
        K2_ADD(8, 1);
        K2_ADD(10, n_all);
        K2_ADD(11, h.n_ca);
    }
    }
    if (lane == 0 && cand_counter && cand_total) atomicAdd(cand_counter, cand_total);
}

// ---- K2b
__host__ __device__ inline size_t gather_smem_bytes(int D) { return (size_t)D * 4 + kW2Cap * 4; }

__global__ void __launch_bounds__(kW2Warps * 32, 8)
gather_kernel(RowArgs A, K2Scratch K)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw + (size_t)warp * gather_smem_bytes(A.D);
    float4 *a = reinterpret_cast<float4 *>(base);
    int32_t *cols = reinterpret_cast<int32_t *>(base + (size_t)A.D * 4);
    const int d4 = A.D >> 2;
    for (;;) {  // rows are claimed a few at a time: their cost varies (40 to 250 gathers), a static split leaves stragglers
        long long b0 = 0;
        if (lane == 0) b0 = (long long)atomicAdd(K.next + 1, (unsigned long long)K.claim);
        b0 = __shfl_sync(0xFFFFFFFFu, b0, 0);
        if (b0 >= A.n_rows) break;
        const int64_t b_end = b0 + K.claim < A.n_rows ? b0 + K.claim : A.n_rows;
    for (int64_t b = b0; b < b_end; ++b) {
        const K2Row h = K.hdr[b];
        if (h.state != 0 || h.n_ca == 0) continue;
        const int64_t i = A.row0 + b;
        __syncwarp();
        const float4 *src = reinterpret_cast<const float4 *>(A.img_emb + i * A.D);
        for (int q = lane; q < d4; q += 32) a[q] = src[q];
        const unsigned long long *pk = K.pk + b * kW2Cap;
        for (int e = lane; e < h.n_ca; e += 32) cols[e] = packed_col(pk[e]);
        __syncwarp();
        float *dst = K.dotv + b * kW2Cap;
        float keep = 0.f;  // lane l keeps the dot products of entries l, l + 32, ...: one coalesced store per 32
        for (int e = 0; e < h.n_ca; e += 2) {
            const int c0 = cols[e], c1 = cols[e + 1 < h.n_ca ? e + 1 : e];
            MMA_CHECK(h.n_ca <= kW2Cap && c0 >= 0 && c0 < A.M && c1 >= 0 && c1 < A.M);
            float d0, d1;
            warp_dot2(a, reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)c0 * A.D),
                      reinterpret_cast<const float4 *>(A.chk_emb + (int64_t)c1 * A.D), d4, lane, d0, d1);
            if (lane == (e & 31)) keep = d0;
            if (lane == ((e + 1) & 31)) keep = d1;
            if ((e & 31) == 30 || e + 2 >= h.n_ca) {
                const int e_w = (e & ~31) + lane;
                if (e_w < h.n_ca) dst[e_w] = keep;
            }
        }
    }
    }
}

// ---- K2c
struct RankSmem { unsigned long long *pk, *xk, *spk; int32_t *xj, *spcol; };
__host__ __device__ inline size_t rank_smem_bytes() { return kW2Cap * (8 + 8 + 4) + kW2Sp * (8 + 4); }

__global__ void __launch_bounds__(kW2Warps * 32, 6)
rank_kernel(RowArgs A, K2Scratch K, const float *__restrict__ eps_chunk_max, int32_t *fail_rows, int32_t *fail_count,
            unsigned long long *fail_thr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int KPT = kW2Cap / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RankSmem w;
    {
        unsigned char *base = smem_raw + (size_t)warp * rank_smem_bytes();
        w.pk = reinterpret_cast<unsigned long long *>(base); base += kW2Cap * 8;
        w.xk = reinterpret_cast<unsigned long long *>(base); base += kW2Cap * 8;
        w.spk = reinterpret_cast<unsigned long long *>(base); base += kW2Sp * 8;
        w.xj = reinterpret_cast<int32_t *>(base); base += kW2Cap * 4;
        w.spcol = reinterpret_cast<int32_t *>(base);
    }
    const RunParams &rp = A.rp;
    const float eps_chunk = eps_chunk_max[0];
    for (;;) {  // rows are claimed a few at a time: their cost varies (40 to 250 gathers), a static split leaves stragglers
        long long b0 = 0;
        if (lane == 0) b0 = (long long)atomicAdd(K.next + 2, (unsigned long long)K.claim);
        b0 = __shfl_sync(0xFFFFFFFFu, b0, 0);
        if (b0 >= A.n_rows) break;
        const int64_t b_end = b0 + K.claim < A.n_rows ? b0 + K.claim : A.n_rows;
    for (int64_t b = b0; b < b_end; ++b) {
        const K2Row h = K.hdr[b];
        if (h.state != 0) continue;
        MMA_CHECK(h.n_ca >= 0 && h.n_ca <= h.n_all && h.n_all <= kW2Cap);
        K2_T(r0_);
        const int64_t i = A.row0 + b;
        const int n_all = h.n_all, n_ca = h.n_ca;
        const int64_t off0 = A.offsets[i];
        const int c = (int)(A.offsets[i + 1] - off0);
        const float na = A.img_n2[i];
        const double eps = row_eps(A, i, eps_chunk);
        __syncwarp();
        // ---- the row's record: union (approximate scores, columns), dot products, same-page entries
        const unsigned long long *gpk = K.pk + b * kW2Cap;
        const float *gdot = K.dotv + b * kW2Cap;
        double cosp = 0.0, lex = 0.0, pos = 0.0;
        int spj = -1;
        if (lane < c) {
            const double *sp = K.sp + b * (3 * kW2Sp);
            spj = A.sorted_chunk[A.sp_start[i] + lane];
            cosp = sp[lane]; lex = sp[kW2Sp + lane]; pos = sp[2 * kW2Sp + lane];
        }
        w.spcol[lane] = spj;
        bool viol = false;
        WKey me[KPT];
        {
            unsigned long long p[KPT];
            float dv[KPT], nb[KPT];
#pragma unroll
            for (int r = 0; r < KPT; r += 2) {
                const int e = lane * KPT + r;
                ulonglong2 t = make_ulonglong2(0ull, 0ull);
                if (e < n_all) t = *reinterpret_cast<const ulonglong2 *>(gpk + e);
                p[r] = t.x; p[r + 1] = t.y;
            }
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const int e = lane * KPT + r;
                w.pk[e] = p[r];
                dv[r] = e < n_ca ? gdot[e] : 0.f;
                nb[r] = e < n_ca ? A.chk_n2[packed_col(p[r])] : 1.f;
            }
            // exact cosine of the re-scored candidates.  The certificate rests on |exact - approximate| <= eps: every
            // candidate tests that bound; a violation (an input the error model does not cover) withdraws the row's
            // certificate, and the row is ranked by the exact scan instead.
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const int e = lane * KPT + r;
                me[r].k = 0ull; me[r].j = 0x7FFFFFFF;
                if (e < n_ca) {
                    const double x = sim_from_sums(dv[r], na, nb[r]);
                    if (!(fabs(x - (double)packed_score(p[r])) <= eps)) viol = true;
                    me[r].k = ord64(x); me[r].j = packed_col(p[r]);
                }
            }
        }
        K2_T(r1_);
        K2_ADD(5, r1_ - r0_);
        // ---- one sort of the candidates by exact cosine (every schema ranks them alike)
        warp_sort<KPT>(me, lane, n_ca);
#pragma unroll
        for (int r = 0; r < KPT; ++r) { w.xk[lane * KPT + r] = me[r].k; w.xj[lane * KPT + r] = me[r].j; }
        bool ok = true;
        if (__any_sync(FULL, viol)) {
            ok = false;
            if (lane == 0 && A.viol_counter) atomicAdd(A.viol_counter, 1ull);
        }
        __syncwarp();
        K2_T(r2_);
        K2_ADD(6, r2_ - r1_);
        // nothing that was not re-scored can have an exact cosine above this
        const double bound = (double)(n_ca < n_all ? fmaxf(packed_score(w.pk[n_ca]), h.tau_union) : h.tau_union) + eps;
        // ---- per schema: the same-page entries are merged by counting (finish_row)
        const int64_t io = i - A.o_row0;
        const int64_t p0 = off0 - A.pair0;
        const int n = n_ca + c;
        const int n_top = n_ca < rp.kmax ? n_ca : rp.kmax;  // candidates that can reach the top-K lists
        for (int si = 0; si < rp.S; ++si) {
            const int s = rp.schema[si];
            double sc_p = 0.0;
            if (lane < c) {
                sc_p = ranking_score(rp, s, cosp, lex, pos);
                if (A.out.pair_score) A.out.pair_score[(int64_t)si * A.P_out + p0 + lane] = sc_p;
            }
            w.spk[lane] = ord64(sc_p);
            __syncwarp();
            const int64_t o_top = ((int64_t)si * A.o_rows + io) * rp.kmax;
            bool bad = false, have_kth = false;
            double kth = -CUDART_INF;
            for (int t = lane; t < c + n_top; t += 32) {
                unsigned long long k;
                int j, at;
                double sc;
                if (t < c) {  // a same-page entry: binary search in the sorted candidates, count the other same-page entries
                    k = w.spk[t]; j = w.spcol[t]; sc = sc_p;
                    int lo = 0, hi = n_ca;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (key_before(w.xk[mid], w.xj[mid], k, j)) lo = mid + 1; else hi = mid;
                    }
                    at = lo;
                    for (int q = 0; q < c; ++q) at += key_before(w.spk[q], w.spcol[q], k, j);
                    if (at < rp.kneed) {
                        bool known = true;  // `at` is the pair's rank ...
                        if (!(sc > bound)) {  // ... unless a column that was not re-scored may beat it
                            const int ahead = count_above_packed(w.pk, n_all, sc + eps) - n_ca;
                            known = false;
                            if (at + (ahead > 0 ? ahead : 0) < rp.kneed) bad = true;  // else: beyond the cutoff either way
                        }
                        if (known && A.out.pair_rank) A.out.pair_rank[(int64_t)si * A.P_out + p0 + t] = at + 1;
                    }
                } else {      // a candidate: its sorted position + the same-page entries that beat it
                    k = w.xk[t - c]; j = w.xj[t - c]; sc = unord64(k);
                    at = t - c;
                    for (int q = 0; q < c; ++q) at += key_before(w.spk[q], w.spcol[q], k, j);
                }
                MMA_CHECK(at >= 0 && at < n && j >= 0 && j < A.M && io >= 0 && io < A.o_rows &&
                          (t >= c || (p0 + t >= 0 && p0 + t < A.P_out)));
                if (at < rp.kmax && A.out.topk_idx) {
                    A.out.topk_idx[o_top + at] = (int64_t)j + rp.col_offset;
                    A.out.topk_score[o_top + at] = sc;
                }
                if (at == rp.kmax - 1) { kth = sc; have_kth = true; }
            }
            if (A.out.topk_idx)
                for (int r = n + lane; r < rp.kmax; r += 32) {  // fewer entries than the lists are wide
                    A.out.topk_idx[o_top + r] = -1; A.out.topk_score[o_top + r] = -CUDART_INF;
                }
            const unsigned hk = __ballot_sync(FULL, have_kth);
            if (hk) kth = __longlong_as_double(__shfl_sync(FULL, __double_as_longlong(kth), __ffs(hk) - 1));
            bad = __any_sync(FULL, bad);
            ok = ok && !bad && (kth > bound);
            __syncwarp();
        }
        K2_T(r3_);
        K2_ADD(7, r3_ - r2_);
        K2_ADD(12, 1);
        if (!ok && lane == 0) {
            const int slot = atomicAdd(fail_count, 1);
            fail_rows[slot] = (int32_t)i;
            if (fail_thr) fail_thr[slot] = h.thr;
        }
    }
    }
}

// ---------------------------------------------------------------------------
// Exact scan
// ---------------------------------------------------------------------------
// Stage 1 for rows that failed the certificate: the whole GPU scans the row's columns (work item = row x column
// split) in the canonical fp32 order and keeps what reaches the row's threshold thr (a lower bound of its kneed-th
// best exact cosine, from the failed attempt).  Almost nothing does -- the candidate set missed the certificate
// by a margin, not by much -- so stage 2 (exact_scan_kernel) ranks a handful of entries instead of streaming
// M columns through one CTA.  A row whose survivors overflow `cap` is streamed by stage 2 as before.
constexpr int kPreThreads = 256;
constexpr int kPreMinCols = 2048;  // columns per work item, at least

__global__ void __launch_bounds__(kPreThreads)
exact_prefilter_kernel(RowArgs A, const int32_t *__restrict__ rows, const int32_t *__restrict__ n_rows_dev,
                       const unsigned long long *__restrict__ thr, Key *scan_buf, int32_t *scan_cnt, int cap,
                       int max_slots)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *a = reinterpret_cast<float *>(smem_raw);
    const int n_fail = min(*n_rows_dev, max_slots);
    if (n_fail == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d4 = A.D >> 2;
    int64_t n_sp = (int64_t)gridDim.x / n_fail;
    const int64_t sp_max = (A.M + kPreMinCols - 1) / kPreMinCols;
    if (n_sp > sp_max) n_sp = sp_max;
    if (n_sp < 1) n_sp = 1;
    const int64_t cols_per = (A.M + n_sp - 1) / n_sp;
    for (int64_t w = blockIdx.x; w < (int64_t)n_fail * n_sp; w += gridDim.x) {
        const int slot = (int)(w / n_sp);
        const int64_t sp = w % n_sp;
        const int64_t i = rows[slot];
        __syncthreads();
        for (int c = threadIdx.x; c < d4; c += kPreThreads)
            reinterpret_cast<float4 *>(a)[c] = reinterpret_cast<const float4 *>(A.img_emb + i * A.D)[c];
        __syncthreads();
        const unsigned long long t = thr[slot];
        const float na = A.img_n2[i];
        const uint64_t ik = A.img_key[i];
        const int64_t j0 = sp * cols_per, j1 = min(A.M, j0 + cols_per);
        for (int64_t j = j0 + 2 * warp; j < j1; j += 2 * (kPreThreads / 32)) {
            const int64_t jb = j + 1 < j1 ? j + 1 : j;
            float d0, d1;
            warp_dot2(reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(A.chk_emb + j * A.D),
                      reinterpret_cast<const float4 *>(A.chk_emb + jb * A.D), d4, lane, d0, d1);
            if (lane < 2) {
                const int64_t jj = lane == 0 ? j : j + 1;
                if (jj < j1 && !(ik != MMALIGN_NULL_KEY && A.chk_key[jj] == ik)) {  // same-page chunks enter through the pair index
                    const unsigned long long k = ord64(sim_from_sums(lane == 0 ? d0 : d1, na, A.chk_n2[jj]));
                    if (k >= t) {
                        const int pos = atomicAdd(scan_cnt + slot, 1);
                        if (pos < cap) { Key x; x.k = k; x.j = (int32_t)jj; x.e = 0; scan_buf[(int64_t)slot * cap + pos] = x; }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
exact_scan_kernel(RowArgs A, const int32_t *rows, const int32_t *n_rows_dev, int64_t n_rows_host,
                  const Key *scan_buf, const int32_t *scan_cnt, int scan_cap, int scan_slots)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RowSmem sm = carve(smem_raw, A.ent_cap, A.sp_cap);
    __shared__ int s_cnt;
    __shared__ unsigned long long s_thr;  // key of the kneed-th best candidate so far (0 = none yet)
    __shared__ int s_thr_j;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_rows = n_rows_dev ? (int64_t)*n_rows_dev : n_rows_host;
    const int d4 = A.D >> 2;
    const int kneed = A.rp.kneed;
    for (int64_t b = blockIdx.x; b < n_rows; b += gridDim.x) {
        const int64_t i = rows ? rows[b] : A.row0 + b;
        const int64_t p0 = A.offsets[i];
        const int c = (int)(A.offsets[i + 1] - p0);
        if (threadIdx.x == 0) { s_cnt = 0; s_thr = 0ull; s_thr_j = 0x7FFFFFFF; }
        stage_row(A, sm, i);
        // a previous (uncertified) pass may have written ranks for this row
        if (A.out.pair_rank)
            for (int t = threadIdx.x; t < c * A.rp.S; t += kThreads)
                A.out.pair_rank[(int64_t)(t / c) * A.P_out + (p0 - A.pair0) + (t % c)] = 0;
        __syncthreads();
        if (c > A.sp_cap) {
            if (threadIdx.x == 0) atomicExch(A.error_flag, 1);
            continue;
        }
        const float na = A.img_n2[i];
        const uint64_t ik = A.img_key[i];
        // stage 1 already found every column that can matter?
        const int pre = (scan_buf && b < scan_slots) ? scan_cnt[b] : -1;
        const bool prefiltered = pre >= 0 && pre <= scan_cap && pre <= A.ent_cap;
        if (prefiltered) {
            for (int e = threadIdx.x; e < pre; e += kThreads) sm.buf[e] = scan_buf[b * scan_cap + e];
            if (threadIdx.x == 0) s_cnt = pre;
            __syncthreads();
        }
        for (int64_t base = prefiltered ? A.M : 0; base < A.M; base += kScanRound) {
            const unsigned long long thr = s_thr;
            const int thr_j = s_thr_j;
            float my_dot = 0.f;  // lane q keeps the dot product of column base + warp * 16 + q
            for (int q = 0; q < 16; ++q) {
                const int64_t j = base + warp * 16 + q;
                if (j >= A.M) break;
                if (ik != MMALIGN_NULL_KEY && A.chk_key[j] == ik) continue;  // enters as a same-page entry
                const float dot = warp_dot(reinterpret_cast<const float4 *>(sm.a),
                                           reinterpret_cast<const float4 *>(A.chk_emb + j * A.D), d4, lane);
                if (lane == q) my_dot = dot;
            }
            {
                const int64_t j = base + warp * 16 + lane;
                if (lane < 16 && j < A.M && !(ik != MMALIGN_NULL_KEY && A.chk_key[j] == ik)) {
                    const unsigned long long k = ord64(sim_from_sums(my_dot, na, A.chk_n2[j]));
                    if (key_before(k, (int)j, thr, thr_j)) {
                        const int pos = atomicAdd(&s_cnt, 1);
                        Key x; x.k = k; x.j = (int32_t)j; x.e = 0;
                        MMA_CHECK(pos >= 0 && pos < A.ent_cap);  // (a round adds at most kScanRound entries behind the shrink)
                        sm.buf[pos] = x;
                    }
                }
            }
            __syncthreads();
            if (s_cnt > A.ent_cap - kScanRound) {  // uniform: shrink to the best kneed
                const int cnt = s_cnt;
                sort_keys(sm.buf, cnt);
                if (threadIdx.x == 0) {
                    s_cnt = cnt < kneed ? cnt : kneed;
                    if (cnt >= kneed) { s_thr = sm.buf[kneed - 1].k; s_thr_j = sm.buf[kneed - 1].j; }
                }
                __syncthreads();
            }
        }
        const int cnt = s_cnt;
        sort_keys(sm.buf, cnt);
        const int n_ca = cnt < kneed ? cnt : kneed;
        for (int e = threadIdx.x; e < n_ca; e += kThreads) sm.cols[c + e] = sm.buf[e].j;
        __syncthreads();
        score_same_page(A, sm, i, c);
        finish_row(A, sm, i, n_ca, -CUDART_INF, 0.0, nullptr, 0);
        __syncthreads();
    }
}

static void size_caps(RowArgs &A, const PairIndex &px, int64_t union_entries)
{
    int64_t sp = 8;
    while (sp < px.c_max && sp < kSpCapMax) sp <<= 1;
    int64_t ent = 256;  // sort_keys needs room for one key per thread
    while (ent < union_entries + sp && ent < kEntCapMax) ent <<= 1;
    A.sp_cap = (int)sp;
    A.ent_cap = (int)ent;
}

static RowArgs make_args(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                         const Outputs &out, int32_t *error_flag, const RowRange *range = nullptr)
{
    RowArgs A;
    A.img_emb = img.emb; A.img_key = img.key; A.img_bbox = img.bbox; A.img_terms = img.terms;
    A.img_n2 = img.norm2; A.img_err = img.err;
    A.chk_emb = chk.emb; A.chk_key = chk.key; A.chk_bbox = chk.bbox; A.chk_terms = chk.terms;
    A.chk_n2 = chk.norm2;
    A.N = img.n; A.M = chk.n; A.D = img.D; A.term_words = chk.term_words;
    A.offsets = px.offsets; A.sorted_chunk = px.sorted_chunk; A.sp_start = px.sp_start; A.P = px.P;
    A.rp = rp; A.out = out; A.error_flag = error_flag;
    A.viol_counter = error_flag ? reinterpret_cast<unsigned long long *>(error_flag + 2) : nullptr;  // api.cu: `small` layout
    A.ent_cap = kEntCapMax; A.sp_cap = kSpCapMax;
    A.row0 = 0; A.n_rows = img.n; A.pair0 = 0; A.P_out = px.P;
    if (range && range->n_rows >= 0 && !(range->row0 == 0 && range->n_rows == 0)) {
        A.row0 = range->row0; A.n_rows = range->n_rows; A.pair0 = range->pair0; A.P_out = range->P_out;
    }
    A.o_row0 = A.row0; A.o_rows = A.n_rows;
    if (range && range->o_rows > 0) { A.o_row0 = range->o_row0; A.o_rows = range->o_rows; }
    A.need_lex = A.need_pos = false;
    for (int q = 0; q < rp.S; ++q) {
        A.need_lex = A.need_lex || rp.schema[q] == 1 || rp.schema[q] == 3;
        A.need_pos = A.need_pos || rp.schema[q] == 2 || rp.schema[q] == 3;
    }
    return A;
}

cudaError_t launch_rescore(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                           const CandLists *lists, const float *eps_chunk_max, const Outputs &out,
                           int32_t *fail_rows, int32_t *fail_count, unsigned long long *fail_thr,
                           unsigned long long *cand_counter, int32_t *error_flag, const float *tau_global,
                           int32_t *cert_count, RowRange range, cudaStream_t st, int64_t grid_limit,
                           int32_t *big_rows, int32_t *big_count, void *k2_scratch, long long *n_launches)
{
    if (n_launches) *n_launches = 0;
    if (img.n == 0) return cudaSuccess;
    RowArgs A = make_args(img, chk, px, rp, out, error_flag, &range);
    if (A.n_rows == 0) return cudaSuccess;
    CandLists L = lists ? *lists : CandLists();
    // shared memory sized for this launch: the union of a row's lists after the final compaction, plus its page
    size_caps(A, px, !lists ? 0 : L.imp_keys ? (int64_t)L.imp_src * L.imp_stride
                                             : (int64_t)2 * L.n_splits * (L.kprime_list + 16));
    const size_t smem = row_smem_bytes(img.D, A.ent_cap, A.sp_cap);
    cudaError_t e = cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t grid = A.n_rows < (int64_t)sm_count() * 16 ? A.n_rows : (int64_t)sm_count() * 16;
    if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;  // (the rows are handed out with a grid stride)
    // The warp-per-row kernels (select, gather, rank) take every row of the common shape; what they hand over
    // (big_count rows) goes through the block-per-row kernel in row-list mode.  Runs that certify across ranks or
    // return the deep lists, and same-page mode, are the block kernel's alone.
    const size_t sm_a = select_smem_bytes(img.D) * kW2Warps, sm_b = gather_smem_bytes(img.D) * kW2Warps,
                 sm_c = rank_smem_bytes() * kW2Warps;
    const bool by_warp = lists && k2_scratch && big_rows && big_count && fail_rows && fail_count && !tau_global &&
                         !cert_count && !out.deep_idx && lists_per_row(L) <= 32 && sm_a * 2 <= 200 * 1024;
    if (by_warp) {
        K2Scratch K = carve_scratch(k2_scratch, A.n_rows);
        K.claim = A.n_rows >= (int64_t)sm_count() * 32 * 64 ? 4 : 1;
        if ((e = cudaMemsetAsync(K.next, 0, 3 * sizeof(unsigned long long), st)) != cudaSuccess) return e;
        auto grid_for = [&](size_t smem_cta, int max_per_sm) {
            int per_sm = (int)(200 * 1024 / smem_cta);
            if (per_sm > max_per_sm) per_sm = max_per_sm;
            int64_t g = (A.n_rows + kW2Warps - 1) / kW2Warps;
            if (g > (int64_t)sm_count() * per_sm) g = (int64_t)sm_count() * per_sm;
            if (grid_limit > 0) { const int64_t lim = grid_limit / 8 * per_sm; if (g > lim) g = lim > 0 ? lim : 1; }
            return (unsigned)g;
        };
        if ((e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b)) != cudaSuccess) return e;
        select_kernel<<<grid_for(sm_a, 8), kW2Warps * 32, sm_a, st>>>(A, L, K, eps_chunk_max, fail_rows, fail_count, fail_thr,
                                                                   cand_counter, big_rows, big_count);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        gather_kernel<<<grid_for(sm_b, 8), kW2Warps * 32, sm_b, st>>>(A, K);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        rank_kernel<<<grid_for(sm_c, 6), kW2Warps * 32, sm_c, st>>>(A, K, eps_chunk_max, fail_rows, fail_count, fail_thr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (grid > (int64_t)sm_count() * 2) grid = (int64_t)sm_count() * 2;  // (few rows, if any)
        if (n_launches) *n_launches += 3;
    }
    if (n_launches) *n_launches += 1;
    rescore_kernel<<<(unsigned)grid, kThreads, smem, st>>>(A, L, lists != nullptr, eps_chunk_max, fail_rows,
                                                           fail_count, fail_thr, cand_counter, tau_global, cert_count,
                                                           by_warp ? big_rows : nullptr, by_warp ? big_count : nullptr);
    return cudaGetLastError();
}

cudaError_t launch_exact_scan(const Side &img, const Side &chk, const PairIndex &px, const RunParams &rp,
                              const int32_t *rows, const int32_t *n_rows_dev, int64_t n_rows_host,
                              const Outputs &out, int32_t *error_flag, RowRange range, const ScanScratch *pre,
                              cudaStream_t st)
{
    if (img.n == 0 || (!n_rows_dev && n_rows_host == 0)) return cudaSuccess;
    RowArgs A = make_args(img, chk, px, rp, out, error_flag, &range);
    size_caps(A, px, kEntCapMax);  // the scan's streaming buffer wants the full capacity
    const size_t smem = row_smem_bytes(img.D, A.ent_cap, A.sp_cap);
    cudaError_t e = cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const bool two_stage = pre && pre->buf && pre->thr && n_rows_dev && rows && chk.n > 0;
    if (two_stage) {
        e = cudaMemsetAsync(pre->cnt, 0, sizeof(int32_t) * kScanSlots, st);
        if (e != cudaSuccess) return e;
        exact_prefilter_kernel<<<(unsigned)(sm_count() * 4), kPreThreads, (size_t)img.D * sizeof(float), st>>>(
            A, rows, n_rows_dev, pre->thr, reinterpret_cast<Key *>(pre->buf), pre->cnt, kScanCap, kScanSlots);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    int64_t grid = (int64_t)sm_count() * 4;
    if (!n_rows_dev && n_rows_host < grid) grid = n_rows_host;
    exact_scan_kernel<<<(unsigned)grid, kThreads, smem, st>>>(A, rows, n_rows_dev, n_rows_host,
                                                              two_stage ? reinterpret_cast<const Key *>(pre->buf) : nullptr,
                                                              two_stage ? pre->cnt : nullptr, kScanCap, kScanSlots);
    return cudaGetLastError();
}

size_t scan_scratch_bytes() { return (size_t)kScanSlots * kScanCap * sizeof(Key); }

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
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
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

// per-row completeness threshold of this rank's lists (sharded runs exchange it with an all-reduce max)
__global__ void row_tau_kernel(CandLists L, int64_t N, float *tau_row)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rb = i >> 7;
        const int r = (int)(i & 127);
        float t = -CUDART_INF_F;
        for (int l = 0; l < 2 * L.n_splits; ++l)
            t = fmaxf(t, L.tau[(((int64_t)(l >> 1) * L.n_row_blocks + rb) * 2 + (l & 1)) * 128 + r]);
        tau_row[i] = t;
    }
}

// Packs the rows' candidates for the all-to-all of the sharded run: destination d owns image rows
// [d * slab_rows, (d + 1) * slab_rows), so output row w IS global image row w.  One warp per row: the row's
// lists are concatenated above the row's completeness threshold, columns become global chunk indices.
__global__ void export_lists_kernel(CandLists L, int64_t N, int64_t total_rows, int stride, int64_t col_offset,
                                    uint64_t *__restrict__ keys, int32_t *__restrict__ count, float *__restrict__ tau)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_l = L.keys ? 2 * L.n_splits : 0;
    for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; w < total_rows;
         w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        float t = -CUDART_INF_F;
        int n = 0;
        if (w < N) {
            for (int l = 0; l < n_l; ++l) t = fmaxf(t, list_view(L, w, 0, l).tau);
            uint64_t *dst = keys + w * stride;
            for (int l = 0; l < n_l; ++l) {
                const ListView v = list_view(L, w, 0, l);
                for (int e0 = 0; e0 < v.cnt; e0 += 32) {
                    const int e = e0 + lane;
                    uint64_t k = 0;
                    bool keep = false;
                    if (e < v.cnt) { k = v.keys[e]; keep = cand_score(k) > t; }
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
                    const int pos = n + __popc(m & lt_mask);
                    MMA_CHECK(e >= v.cnt || (int64_t)cand_col(k) + col_offset >= 0);
                    if (keep && pos < stride)
                        dst[pos] = (k & 0xFFFFFFFF00000000ull) | (uint64_t)((int64_t)cand_col(k) + col_offset);
                    n += __popc(m);
                }
            }
        }
        if (lane == 0) { count[w] = n > stride ? -1 : n; tau[w] = t; }
    }
}

cudaError_t launch_export_lists(const CandLists &L, int64_t N, int n_dest, int64_t slab_rows, int stride,
                                int64_t col_offset, uint64_t *keys, int32_t *count, float *tau, cudaStream_t st)
{
    const int64_t total = (int64_t)n_dest * slab_rows;
    int64_t grid = (total * 32 + 255) / 256;
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    export_lists_kernel<<<(unsigned)grid, 256, 0, st>>>(L, N, total, stride, col_offset, keys, count, tau);
    return cudaGetLastError();
}

cudaError_t launch_row_tau(const CandLists &L, int64_t N, float *tau_row, cudaStream_t st)
{
    if (N == 0) return cudaSuccess;
    int64_t grid = (N + 255) / 256;
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
    row_tau_kernel<<<(unsigned)grid, 256, 0, st>>>(L, N, tau_row);
    return cudaGetLastError();
}

cudaError_t launch_pair_chunk(const PairIndex &px, int64_t N, int64_t col_offset, int64_t *pair_chunk,
                              cudaStream_t st)
{
    if (N == 0 || px.P == 0) return cudaSuccess;
    int64_t grid = (N * 32 + 255) / 256;
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
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
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
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
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    count_beating_kernel<<<(unsigned)grid, 256, 0, st>>>(deep_idx, deep_score, N, S, K, n_q, q_image, q_chunk,
                                                         q_score, counts);
    return cudaGetLastError();
}

MMA_CHECK_READER(check_read_rescore)

} // namespace mma
