// api.cu -- the C ABI of include/mmalign.h: context, uploads, and the
// K0 -> K1 -> K2 -> exact scan -> K4 pipeline.  No CPU fallback anywhere: every
// result is produced by the kernels in prep.cu / fused_tc.cu / rescore.cu.
#include "common.cuh"
#include <cuda.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <chrono>
#include <math.h>
#include <stdlib.h>

using namespace mma;

static thread_local char g_err[512] = "";

// MMALIGN_TRACE=1: host wall-clock of the phases of a call, to stderr
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char *what;
    explicit Trace(const char *w) : on(getenv("MMALIGN_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
    void mark(const char *phase)
    {
        if (!on) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[mmalign %s] %-22s %9.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t - t0).count());
        t0 = t;
    }
};

struct DevBuf {  // growable device scratch; frees itself (locals on error paths, context members on destroy)
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct SideStore {
    Side s;
    DevBuf up[4];               // uploads of host inputs: emb, key, bbox, terms (reused across set_* calls)
    DevBuf bf16, norm2, err, errmax;
    float *err_max = nullptr;   // [1] max rounding-error norm over the rows
    alignas(64) CUtensorMap tmap;
    bool ready = false;
    void release()
    {
        for (DevBuf &b : up) b.release();
        bf16.release(); norm2.release(); err.release(); errmax.release();
        s = Side();
        err_max = nullptr;
        ready = false;
    }
};

struct mmalign_ctx {
    int device = 0;
    int sm_count = 148;
    char err[512] = "";
    SideStore img, chk;
    int64_t n_terms = 0, col_offset = 0;
    PairIndex px;
    bool px_ready = false;
    DevBuf px_offsets, px_sorted, px_start, px_scratch;
    DevBuf list_keys, list_tau, list_count;
    DevBuf fail_rows, fail_thr, scan_buf, scan_cnt, small;  // small: fail_count, cand_counter, error_flag, k_list, stats
    DevBuf metrics_scratch, stage; // stage: device copies of host outputs
    DevBuf term_table, text_off, text_bytes;  // mmalign_term_bitsets: term table, uploads of host texts
    CandLists lists;               // written by the last fused pass
    bool lists_valid = false;
    int64_t lists_col0 = 0;        // first chunk row of the column range the lists were built on
    int64_t last_fused_us = 0;     // CUDA-event time of the last mmalign_fused_pass
    cudaEvent_t ev[5] = {};        // run start, after fused, after rescore, after exact scan, after metrics
};

static int fail(mmalign_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    snprintf(g_err, sizeof g_err, "%s", buf);
    if (c) snprintf(c->err, sizeof c->err, "%s", buf);
    return code;
}

#define CU(c, x)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess)                                                                     \
            return fail((c), MMALIGN_ECUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                       \
    } while (0)

static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int mmalign_abi_version(void) { return MMALIGN_ABI_VERSION; }

extern "C" const char *mmalign_last_error(const mmalign_ctx *ctx) { return ctx ? ctx->err : g_err; }

extern "C" int mmalign_create(mmalign_ctx **out, int device)
{
    if (!out) return fail(nullptr, MMALIGN_EINVAL, "mmalign_create: ctx is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, MMALIGN_EDEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(nullptr, MMALIGN_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, MMALIGN_EDEVICE, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)",
                    device, prop.major, prop.minor);
    CU(nullptr, cudaSetDevice(device));
    mmalign_ctx *c = new mmalign_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (c->small.reserve(4096) != cudaSuccess) { delete c; return fail(nullptr, MMALIGN_ECUDA, "cudaMalloc failed"); }
    for (cudaEvent_t &e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete c; return fail(nullptr, MMALIGN_ECUDA, "cudaEventCreate failed"); }
    *out = c;
    return MMALIGN_OK;
}

extern "C" void mmalign_destroy(mmalign_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->img.release();
    c->chk.release();
    DevBuf *bufs[] = {&c->px_offsets, &c->px_sorted, &c->px_start, &c->px_scratch, &c->list_keys, &c->list_tau, &c->list_count,
                      &c->fail_rows, &c->fail_thr, &c->scan_buf, &c->scan_cnt, &c->small, &c->metrics_scratch, &c->stage,
                      &c->term_table, &c->text_off, &c->text_bytes};
    for (DevBuf *b : bufs) b->release();
    for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
    delete c;
}

template <typename T>
static int adopt(mmalign_ctx *c, DevBuf &buf, const T *src, size_t count, const T **dst, cudaStream_t st)
{
    *dst = nullptr;
    if (!src || count == 0) return MMALIGN_OK;
    if (is_device_ptr(src)) { *dst = src; return MMALIGN_OK; }  // borrowed
    CU(c, buf.reserve(count * sizeof(T)));
    CU(c, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *dst = static_cast<const T *>(buf.p);
    return MMALIGN_OK;
}

static int set_side(mmalign_ctx *c, SideStore &ss, const float *emb, const uint64_t *key, const double *bbox,
                    const uint64_t *terms, int64_t n, int D, int term_words, int box_rows, const char *what)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "%s: ctx is NULL", what);
    CU(c, cudaSetDevice(c->device));
    if (n < 0 || n > 0x7FFFFFF0ll) return fail(c, MMALIGN_EINVAL, "%s: row count %lld out of range", what, (long long)n);
    if (D <= 0 || D % 4 != 0 || D > 4096) return fail(c, MMALIGN_EINVAL, "%s: D=%d must be a multiple of 4 in 4..4096", what, D);
    if (n > 0 && (!emb || !key)) return fail(c, MMALIGN_EINVAL, "%s: emb and page_key are required", what);
    if (term_words < 0 || (terms && term_words == 0)) return fail(c, MMALIGN_EINVAL, "%s: bad term_words", what);
    cudaStream_t st = 0;
    Trace tr(what);
    CU(c, cudaStreamSynchronize(st));
    ss.ready = false;
    c->px_ready = false;
    c->lists_valid = false;
    ss.s = Side();
    Side &s = ss.s;
    s.n = n; s.D = D; s.term_words = term_words;
    int rc;
    if ((rc = adopt(c, ss.up[0], emb, (size_t)n * D, &s.emb, st))) return rc;
    if ((rc = adopt(c, ss.up[1], key, (size_t)n, &s.key, st))) return rc;
    if ((rc = adopt(c, ss.up[3], terms, (size_t)n * term_words, &s.terms, st))) return rc;
    if (bbox) {
        if ((rc = adopt(c, ss.up[2], bbox, (size_t)n * 4, &s.bbox, st))) return rc;
    } else if (n > 0) {  // missing boxes: all zero -> positional score 0 (insert_clip_embeddings.py:161)
        CU(c, ss.up[2].reserve((size_t)n * 4 * sizeof(double)));
        CU(c, cudaMemsetAsync(ss.up[2].p, 0, (size_t)n * 4 * sizeof(double), st));
        s.bbox = static_cast<const double *>(ss.up[2].p);
    }
    const size_t nn = n > 0 ? (size_t)n : 1;
    CU(c, ss.bf16.reserve(nn * D * sizeof(__nv_bfloat16))); s.emb_bf16 = (__nv_bfloat16 *)ss.bf16.p;
    CU(c, ss.norm2.reserve(nn * sizeof(float))); s.norm2 = (float *)ss.norm2.p;
    CU(c, ss.err.reserve(nn * sizeof(float))); s.err = (float *)ss.err.p;
    CU(c, ss.errmax.reserve(sizeof(float))); ss.err_max = (float *)ss.errmax.p;
    CU(c, launch_prep(s, st));
    CU(c, reduce_max_float(s.err, n, ss.err_max, st));
    if (n > 0 && D % 64 == 0) {
        char msg[256];
        if (encode_tensor_map(&ss.tmap, s.emb_bf16, n, D, box_rows, msg, sizeof msg))
            return fail(c, MMALIGN_ECUDA, "%s: %s", what, msg);
    }
    CU(c, cudaStreamSynchronize(st));
    tr.mark("upload + prep");
    ss.ready = true;
    return MMALIGN_OK;
}

extern "C" int mmalign_set_images(mmalign_ctx *c, const float *emb, const uint64_t *key, const double *bbox,
                                  const uint64_t *terms, int64_t n, int32_t D, int32_t term_words)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_set_images: ctx is NULL");
    return set_side(c, c->img, emb, key, bbox, terms, n, D, term_words, 128, "mmalign_set_images");
}

extern "C" int mmalign_set_chunks(mmalign_ctx *c, const float *emb, const uint64_t *key, const double *bbox,
                                  const uint64_t *terms, int64_t m, int32_t D, int32_t term_words,
                                  int64_t n_terms, int64_t col_offset)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_set_chunks: ctx is NULL");
    if (n_terms < 0 || n_terms > (int64_t)term_words * 64)
        return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks: n_terms=%lld does not fit %d term words", (long long)n_terms, term_words);
    if (col_offset < 0) return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks: negative col_offset");
    int rc = set_side(c, c->chk, emb, key, bbox, terms, m, D, term_words, 256, "mmalign_set_chunks");
    if (rc) return rc;
    c->n_terms = n_terms;
    c->col_offset = col_offset;
    return MMALIGN_OK;
}

static int ensure_index(mmalign_ctx *c)
{
    if (!c->img.ready || !c->chk.ready) return fail(c, MMALIGN_ESTATE, "set_images and set_chunks must be called first");
    if (c->img.s.D != c->chk.s.D) return fail(c, MMALIGN_EINVAL, "image D=%d differs from chunk D=%d", c->img.s.D, c->chk.s.D);
    if (c->img.s.terms && c->img.s.term_words != c->chk.s.term_words)
        return fail(c, MMALIGN_EINVAL, "image term_words=%d differs from chunk term_words=%d", c->img.s.term_words, c->chk.s.term_words);
    if (c->px_ready) return MMALIGN_OK;
    CU(c, cudaSetDevice(c->device));
    const int64_t N = c->img.s.n, M = c->chk.s.n;
    CU(c, c->px_offsets.reserve(sizeof(int64_t) * (N + 1)));
    CU(c, c->px_sorted.reserve(sizeof(int32_t) * (M > 0 ? M : 1)));
    CU(c, c->px_start.reserve(sizeof(int64_t) * (N > 0 ? N : 1)));
    c->px.offsets = (int64_t *)c->px_offsets.p;
    c->px.sorted_chunk = (int32_t *)c->px_sorted.p;
    c->px.sp_start = (int64_t *)c->px_start.p;
    const size_t sb = pair_index_scratch_bytes(N, M);
    CU(c, c->px_scratch.reserve(sb));
    CU(c, build_pair_index(c->img.s, c->chk.s, c->px, c->px_scratch.p, sb, 0));
    c->px_ready = true;
    return MMALIGN_OK;
}

extern "C" int mmalign_num_pairs(mmalign_ctx *c, int64_t *num_pairs)
{
    if (!c || !num_pairs) return fail(c, MMALIGN_EINVAL, "mmalign_num_pairs: NULL argument");
    int rc = ensure_index(c);
    if (rc) return rc;
    *num_pairs = c->px.P;
    return MMALIGN_OK;
}

// ---- output staging: host pointers get a device twin that is copied back --------------------
struct Stager {
    mmalign_ctx *c;
    cudaStream_t st;
    struct Item { void *host; void *dev; size_t bytes; };
    std::vector<Item> items;
    size_t used = 0;
    std::vector<std::pair<void **, size_t>> pending;  // (slot to patch, offset)
    template <typename T> int map(T *user, size_t count, T **dev)
    {
        *dev = nullptr;
        if (!user || count == 0) return 0;
        if (is_device_ptr(user)) { *dev = user; return 0; }
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        items.push_back({user, nullptr, count * sizeof(T)});
        pending.push_back({(void **)dev, used});
        used += bytes;
        return 0;
    }
    int commit()
    {
        if (!used) return 0;
        if (c->stage.reserve(used) != cudaSuccess) return fail(c, MMALIGN_ECUDA, "cudaMalloc of %zu staging bytes failed", used);
        for (size_t i = 0; i < pending.size(); ++i) {
            void *d = (char *)c->stage.p + pending[i].second;
            *pending[i].first = d;
            items[i].dev = d;
        }
        return 0;
    }
    int copy_back()
    {
        for (auto &it : items) CU(c, cudaMemcpyAsync(it.host, it.dev, it.bytes, cudaMemcpyDeviceToHost, st));
        return 0;
    }
};

extern "C" int mmalign_get_pairs(mmalign_ctx *c, int64_t *pair_offsets, int64_t *pair_chunk)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_get_pairs: ctx is NULL");
    int rc = ensure_index(c);
    if (rc) return rc;
    cudaStream_t st = 0;
    const int64_t N = c->img.s.n;
    if (pair_offsets)
        CU(c, cudaMemcpyAsync(pair_offsets, c->px.offsets, sizeof(int64_t) * (N + 1), cudaMemcpyDefault, st));
    if (pair_chunk && c->px.P > 0) {
        Stager sg{c, st};
        int64_t *d = nullptr;
        sg.map(pair_chunk, (size_t)c->px.P, &d);
        if ((rc = sg.commit())) return rc;
        CU(c, launch_pair_chunk(c->px, N, c->col_offset, d, st));
        if ((rc = sg.copy_back())) return rc;
    }
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}

static int schema_index(uint32_t bit)
{
    switch (bit) { case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3; }
    return -1;
}

static int parse_params(mmalign_ctx *c, const mmalign_params *prm, RunParams *out)
{
    const Side &chk = c->chk.s;
    RunParams rp = {};
    for (int s = 0; s < 4; ++s)
        if (prm->schema_mask & (1u << s)) rp.schema[rp.S++] = s;
    if (rp.S == 0 || (prm->schema_mask & ~15u)) return fail(c, MMALIGN_EINVAL, "schema_mask=0x%x selects no valid schema", prm->schema_mask);
    if (prm->candidates != MMALIGN_CAND_SAME_PAGE && prm->candidates != MMALIGN_CAND_ALL)
        return fail(c, MMALIGN_EINVAL, "candidates=%d is not MMALIGN_CAND_SAME_PAGE/ALL", prm->candidates);
    if (prm->n_k < 1 || prm->n_k > kMaxK) return fail(c, MMALIGN_EINVAL, "n_k=%d must be in 1..8", prm->n_k);
    rp.n_k = prm->n_k;
    for (int q = 0; q < rp.n_k; ++q) {
        if (prm->k_list[q] < 1 || prm->k_list[q] > 256) return fail(c, MMALIGN_EINVAL, "k_list[%d]=%d must be in 1..256", q, prm->k_list[q]);
        rp.k_list[q] = prm->k_list[q];
        if (rp.k_list[q] > rp.kmax) rp.kmax = rp.k_list[q];
    }
    if (prm->mrr_cutoff < 0 || prm->mrr_cutoff > 256) return fail(c, MMALIGN_EINVAL, "mrr_cutoff=%d must be in 0..256", prm->mrr_cutoff);
    rp.mrr_cutoff = prm->mrr_cutoff;
    rp.kneed = rp.kmax > rp.mrr_cutoff ? rp.kmax : rp.mrr_cutoff;
    rp.candidates = prm->candidates;
    rp.lam_lex = prm->lam_lex; rp.lam_pos = prm->lam_pos; rp.lam_comb = prm->lam_comb;
    rp.n_terms = c->n_terms;
    rp.col_offset = c->col_offset;
    if (prm->eps_scale < 0.f || prm->eps_scale > 1e6f) return fail(c, MMALIGN_EINVAL, "eps_scale=%g must be in 0..1e6 (0 = 1)", (double)prm->eps_scale);
    rp.eps_scale = prm->eps_scale > 0.f ? prm->eps_scale : 1.f;
    const bool needs_terms = (prm->schema_mask & (MMALIGN_LEXICAL | MMALIGN_COMBINED)) != 0;
    if (needs_terms && !chk.terms && chk.n > 0) return fail(c, MMALIGN_EINVAL, "lexical schema requested but the chunks have no term sets");
    if (prm->path < 0 || prm->path > 2) return fail(c, MMALIGN_EINVAL, "path=%d unknown", prm->path);
    if (prm->n_ranks < 0 || prm->n_ranks > 64) return fail(c, MMALIGN_EINVAL, "n_ranks=%d must be in 0..64", prm->n_ranks);
    *out = rp;
    return MMALIGN_OK;
}

// Range checks of the sharded parameters; (0, 0) means everything.
static int resolve_range(mmalign_ctx *c, int64_t lo, int64_t cnt, int64_t total, const char *what, int64_t *o_lo, int64_t *o_cnt)
{
    if (lo == 0 && cnt == 0) { *o_lo = 0; *o_cnt = total; return MMALIGN_OK; }
    if (lo < 0 || cnt < 0 || lo + cnt > total) return fail(c, MMALIGN_EINVAL, "%s [%lld, +%lld) is outside 0..%lld", what, (long long)lo, (long long)cnt, (long long)total);
    *o_lo = lo; *o_cnt = cnt;
    return MMALIGN_OK;
}

// Rows [row0, row0 + n_rows) with their slice of the pair arrays (two 8-byte reads of the pair index).
static int make_row_range(mmalign_ctx *c, int64_t row0, int64_t n_rows, cudaStream_t st, RowRange *out)
{
    RowRange r;
    r.row0 = row0; r.n_rows = n_rows;
    int64_t h[2] = {0, 0};
    CU(c, cudaMemcpyAsync(&h[0], c->px.offsets + row0, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(&h[1], c->px.offsets + row0 + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    r.pair0 = h[0]; r.P_out = h[1] - h[0];
    *out = r;
    return MMALIGN_OK;
}

// K1 of image rows [row0, +n_rows) against chunk rows [col0, +n_cols) of the current corpus; the lists (row and
// column indices relative to the ranges) stay in the context for the rescoring / export pass
static int run_fused(mmalign_ctx *c, const RunParams &rp, int kprime_req, int n_ranks, int64_t row0, int64_t n_rows,
                     int64_t col0, int64_t n_cols, cudaStream_t st, FusedPlan *plan_out)
{
    Side img = c->img.s, chk = c->chk.s;
    if (img.D % 64 != 0) return fail(c, MMALIGN_EINVAL, "the fused path needs D %% 64 == 0 (D=%d); use MMALIGN_PATH_EXACT", img.D);
    if (rp.kneed > 256) return fail(c, MMALIGN_ELIMIT, "kneed=%d exceeds 256", rp.kneed);
    alignas(64) CUtensorMap tmap_a = c->img.tmap, tmap_b = c->chk.tmap;
    char msg[256];
    if (row0 != 0 || n_rows != img.n) {
        if (encode_tensor_map(&tmap_a, img.emb_bf16 + row0 * img.D, n_rows, img.D, 128, msg, sizeof msg)) return fail(c, MMALIGN_ECUDA, "%s", msg);
        img.n = n_rows;
    }
    if (col0 != 0 || n_cols != chk.n) {
        if (encode_tensor_map(&tmap_b, chk.emb_bf16 + col0 * chk.D, n_cols, chk.D, 256, msg, sizeof msg)) return fail(c, MMALIGN_ECUDA, "%s", msg);
        chk.n = n_cols;
    }
    c->lists_col0 = col0;
    FusedPlan plan;
    const int prc = fused_plan(img.n, chk.n, img.D, rp.kneed, kprime_req, c->sm_count, n_ranks, &plan);
    if (prc) return fail(c, MMALIGN_ELIMIT, "no fused plan for N=%lld M=%lld D=%d K'=%d (code %d)", (long long)img.n, (long long)chk.n, img.D, kprime_req, prc);
    CU(c, c->list_keys.reserve((size_t)plan.n_lists * plan.cap * sizeof(uint64_t)));
    CU(c, c->list_tau.reserve((size_t)plan.n_lists * sizeof(float)));
    CU(c, c->list_count.reserve((size_t)plan.n_lists * sizeof(int32_t)));
    CU(c, c->fail_rows.reserve((size_t)img.n * sizeof(int32_t)));
    c->lists = CandLists();
    c->lists.keys = (uint64_t *)c->list_keys.p; c->lists.tau = (float *)c->list_tau.p; c->lists.count = (int32_t *)c->list_count.p;
    CU(c, launch_fused(img, chk, plan, &tmap_a, &tmap_b, c->lists, nullptr, st));
    c->lists_valid = true;
    *plan_out = plan;
    return MMALIGN_OK;
}

// mmalign_run and mmalign_rescore_slab: rank image rows [slab_row0, +slab_rows) (0, 0 = all).  `imported` = the
// candidate lists received from the ranks' fused passes (global chunk indices); without it the fused kernel runs here.
static int run_impl(mmalign_ctx *c, const mmalign_params *prm, mmalign_out *uo, void *stream, const CandLists *imported)
{
    Trace tr("run");
    int rc = ensure_index(c);
    if (rc) return rc;
    tr.mark("pair index");
    cudaStream_t st = (cudaStream_t)stream;
    const Side &img = c->img.s, &chk = c->chk.s;
    const int64_t M = chk.n;
    RunParams rp;
    if ((rc = parse_params(c, prm, &rp))) return rc;
    CU(c, cudaSetDevice(c->device));
    int64_t row0, N;  // the slab; per-row outputs are [S][N][..], per-pair outputs [S][P]
    if ((rc = resolve_range(c, prm->slab_row0, prm->slab_rows, img.n, "slab rows", &row0, &N))) return rc;
    RowRange range;
    if ((rc = make_row_range(c, row0, N, st, &range))) return rc;
    const int64_t P = range.P_out;
    // ---- outputs
    Stager sg{c, st};
    Outputs out = {};
    int64_t *d_hits = nullptr, *d_np = nullptr, *d_stats = nullptr;
    double *d_rr = nullptr, *d_sim = nullptr;
    const size_t SNK = (size_t)rp.S * N * rp.kmax, SP = (size_t)rp.S * P;
    if ((uo->topk_idx == nullptr) != (uo->topk_score == nullptr)) return fail(c, MMALIGN_EINVAL, "topk_idx and topk_score go together");
    sg.map(uo->topk_idx, SNK, &out.topk_idx);
    sg.map(uo->topk_score, SNK, &out.topk_score);
    sg.map(uo->pair_rank, SP, &out.pair_rank);
    sg.map(uo->pair_sim, (size_t)P, &out.pair_sim);
    sg.map(uo->hits, (size_t)rp.S * rp.n_k, &d_hits);
    sg.map(uo->rr_sum, (size_t)rp.S, &d_rr);
    sg.map(uo->sim_sum, 1, &d_sim);
    sg.map(uo->num_pairs, 1, &d_np);
    sg.map(uo->stats, 8, &d_stats);
    sg.map(uo->pair_score, SP, &out.pair_score);
    if ((uo->deep_idx == nullptr) != (uo->deep_score == nullptr)) return fail(c, MMALIGN_EINVAL, "deep_idx and deep_score go together");
    sg.map(uo->deep_idx, (size_t)rp.S * N * rp.kneed, &out.deep_idx);
    sg.map(uo->deep_score, (size_t)rp.S * N * rp.kneed, &out.deep_score);
    // metric sums need the per-pair arrays even if the caller did not ask for them
    const bool want_sums = uo->hits || uo->rr_sum || uo->sim_sum;
    if ((rc = sg.commit())) return rc;
    DevBuf extra;  // scratch per-pair arrays when only the sums were requested
    if (want_sums && (!out.pair_rank || !out.pair_sim) && P > 0) {
        CU(c, extra.reserve(SP * sizeof(int32_t) + 256 + (size_t)P * sizeof(double)));
        if (!out.pair_rank) out.pair_rank = (int32_t *)extra.p;
        if (!out.pair_sim) out.pair_sim = (double *)((char *)extra.p + ((SP * sizeof(int32_t) + 255) & ~(size_t)255));
    }
    tr.mark("params + staging");
    // ---- small device state
    int32_t *fail_count = (int32_t *)c->small.p;
    unsigned long long *cand_counter = (unsigned long long *)((char *)c->small.p + 8);
    int32_t *error_flag = (int32_t *)((char *)c->small.p + 16);
    int32_t *k_list_dev = (int32_t *)((char *)c->small.p + 64);
    CU(c, cudaMemsetAsync(c->small.p, 0, 64, st));
    CU(c, cudaMemcpyAsync(k_list_dev, rp.k_list, sizeof(int32_t) * kMaxK, cudaMemcpyHostToDevice, st));
    if (out.pair_rank && SP) CU(c, cudaMemsetAsync(out.pair_rank, 0, SP * sizeof(int32_t), st));
    long long launches = 2, fused_launches = 0, kprime_used = 0;
    // ---- scoring
    CU(c, cudaEventRecord(c->ev[0], st));
    for (int q = 1; q < 4; ++q) CU(c, cudaEventRecord(c->ev[q], st));  // overwritten by the phases that run
    if (N > 0) {
        if (rp.candidates == MMALIGN_CAND_SAME_PAGE) {
            CU(c, launch_rescore(img, chk, c->px, rp, nullptr, nullptr, out, nullptr, nullptr, nullptr, cand_counter,
                                 error_flag, nullptr, nullptr, range, st));
            launches += 1;
        } else if (prm->path == MMALIGN_PATH_EXACT || M == 0) {
            CU(c, launch_exact_scan(img, chk, c->px, rp, nullptr, nullptr, N, out, error_flag, range, nullptr, st));
            launches += 1;
        } else {
            CU(c, c->fail_thr.reserve((size_t)img.n * sizeof(unsigned long long)));
            CU(c, c->scan_buf.reserve(scan_scratch_bytes()));
            CU(c, c->scan_cnt.reserve(sizeof(int32_t) * kScanSlots));
            ScanScratch pre;
            pre.thr = (unsigned long long *)c->fail_thr.p; pre.buf = c->scan_buf.p; pre.cnt = (int32_t *)c->scan_cnt.p;
            if (imported) {
                CU(c, c->fail_rows.reserve((size_t)img.n * sizeof(int32_t)));
                kprime_used = imported->kprime;
            } else {
                FusedPlan plan;
                if ((rc = run_fused(c, rp, prm->kprime, 1, row0, N, 0, M, st, &plan))) return rc;
                fused_launches = 1;
                launches += 1;
                kprime_used = plan.kprime;
            }
            const CandLists &L = imported ? *imported : c->lists;
            CU(c, cudaEventRecord(c->ev[1], st));
            CU(c, launch_rescore(img, chk, c->px, rp, &L, c->chk.err_max, out, (int32_t *)c->fail_rows.p, fail_count,
                                 pre.thr, cand_counter, error_flag, nullptr, nullptr, range, st));
            CU(c, cudaEventRecord(c->ev[2], st));
            CU(c, launch_exact_scan(img, chk, c->px, rp, (int32_t *)c->fail_rows.p, fail_count, 0, out, error_flag, range, &pre, st));
            CU(c, cudaEventRecord(c->ev[3], st));
            launches += 3;
        }
    }
    tr.mark("launch scoring");
    // ---- metric sums
    if (want_sums) {
        CU(c, c->metrics_scratch.reserve(metrics_scratch_bytes(rp.S, rp.n_k)));
        CU(c, launch_reduce_metrics(out.pair_rank, out.pair_sim, rp.S, P, k_list_dev, rp.n_k, rp.mrr_cutoff, d_hits,
                                    d_rr, d_sim, c->metrics_scratch.p, st));
        launches += 2;
    }
    CU(c, cudaEventRecord(c->ev[4], st));
    // ---- status, stats
    struct { int32_t fail; int32_t pad; unsigned long long cand; int32_t err; } h = {};
    CU(c, cudaMemcpyAsync(&h, c->small.p, 24, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    tr.mark("kernels done");
    if (h.err) { extra.release(); return fail(c, MMALIGN_ELIMIT, "an image has more than 512 same-page chunks (capacity limit of this build)"); }
    if (prm->path == MMALIGN_PATH_FUSED && h.fail > 0) {
        extra.release();
        return fail(c, MMALIGN_ELIMIT, "%d rows were not certified by the fused path (MMALIGN_PATH_FUSED forbids the exact rescan)", h.fail);
    }
    if (d_np) CU(c, cudaMemcpyAsync(d_np, &P, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (d_stats) {
        float t_fused = 0.f, t_resc = 0.f, t_scan = 0.f;
        if (fused_launches || imported) {
            if (fused_launches) cudaEventElapsedTime(&t_fused, c->ev[0], c->ev[1]);
            cudaEventElapsedTime(&t_resc, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&t_scan, c->ev[2], c->ev[3]);
        }
        const int64_t stats[8] = {h.fail, (int64_t)h.cand, imported ? 1 : fused_launches, launches + (imported ? 2 : 0), kprime_used,
                                  imported ? c->last_fused_us : (int64_t)(t_fused * 1000.f), (int64_t)(t_resc * 1000.f),
                                  (int64_t)(t_scan * 1000.f)};
        CU(c, cudaMemcpyAsync(d_stats, stats, sizeof stats, cudaMemcpyHostToDevice, st));
    }
    if ((rc = sg.copy_back())) { extra.release(); return rc; }
    CU(c, cudaStreamSynchronize(st));
    extra.release();
    tr.mark("copy back");
    return MMALIGN_OK;
}

extern "C" int mmalign_run(mmalign_ctx *c, const mmalign_params *prm, mmalign_out *uo, void *stream)
{
    if (!c || !prm || !uo) return fail(c, MMALIGN_EINVAL, "mmalign_run: NULL argument");
    return run_impl(c, prm, uo, stream, nullptr);
}

extern "C" int mmalign_rescore_slab(mmalign_ctx *c, const mmalign_params *prm, const uint64_t *keys, const int32_t *count,
                                    const float *tau, int32_t n_src, int64_t list_rows, int32_t stride, mmalign_out *uo,
                                    void *stream)
{
    if (!c || !prm || !uo || !keys || !count || !tau) return fail(c, MMALIGN_EINVAL, "mmalign_rescore_slab: NULL argument");
    if (n_src < 1 || n_src > 64 || stride < 1) return fail(c, MMALIGN_EINVAL, "mmalign_rescore_slab: bad n_src/stride");
    if (!is_device_ptr(keys) || !is_device_ptr(count) || !is_device_ptr(tau))
        return fail(c, MMALIGN_EINVAL, "mmalign_rescore_slab: the received lists must be device pointers");
    if (prm->candidates != MMALIGN_CAND_ALL || prm->path == MMALIGN_PATH_EXACT)
        return fail(c, MMALIGN_EINVAL, "mmalign_rescore_slab is the second half of the fused path (MMALIGN_CAND_ALL)");
    if (!c->img.ready) return fail(c, MMALIGN_ESTATE, "set_images and set_chunks must be called first");
    int64_t row0, rows;
    int rc = resolve_range(c, prm->slab_row0, prm->slab_rows, c->img.s.n, "slab rows", &row0, &rows);
    if (rc) return rc;
    if (list_rows < rows) return fail(c, MMALIGN_EINVAL, "mmalign_rescore_slab: list_rows=%lld is smaller than the slab (%lld rows)", (long long)list_rows, (long long)rows);
    FusedPlan plan;  // K' as the ranks' fused passes chose it (same arguments, same answer)
    RunParams rp;
    if ((rc = parse_params(c, prm, &rp))) return rc;
    if (fused_plan(c->img.s.n, c->chk.s.n > 0 ? c->chk.s.n : 1, c->img.s.D, rp.kneed, prm->kprime, c->sm_count, 1, &plan))
        return fail(c, MMALIGN_ELIMIT, "no fused plan");
    CandLists L;
    L.imp_keys = keys; L.imp_count = count; L.imp_tau = tau;
    L.imp_src = n_src; L.imp_stride = stride; L.imp_rows = list_rows;
    L.kprime = plan.kprime;
    return run_impl(c, prm, uo, stream, &L);
}

extern "C" int mmalign_num_pairs_range(mmalign_ctx *c, int64_t row0, int64_t rows, int64_t *num_pairs)
{
    if (!c || !num_pairs) return fail(c, MMALIGN_EINVAL, "mmalign_num_pairs_range: NULL argument");
    int rc = ensure_index(c);
    if (rc) return rc;
    int64_t r0, n;
    if ((rc = resolve_range(c, row0, rows, c->img.s.n, "rows", &r0, &n))) return rc;
    RowRange range;
    if ((rc = make_row_range(c, r0, n, 0, &range))) return rc;
    *num_pairs = range.P_out;
    return MMALIGN_OK;
}

extern "C" int mmalign_list_stride(mmalign_ctx *c, int32_t *stride)
{
    if (!c || !stride) return fail(c, MMALIGN_EINVAL, "mmalign_list_stride: NULL argument");
    // an empty shard exports nothing
    *stride = c->lists_valid ? 2 * c->lists.n_splits * (c->lists.kprime_list + kListSlack) : 1;
    return MMALIGN_OK;
}

extern "C" int mmalign_export_lists(mmalign_ctx *c, int32_t n_dest, int64_t slab_rows, int32_t stride, uint64_t *keys,
                                    int32_t *count, float *tau, void *stream)
{
    if (!c || !keys || !count || !tau) return fail(c, MMALIGN_EINVAL, "mmalign_export_lists: NULL argument");
    if (!c->img.ready || !c->chk.ready) return fail(c, MMALIGN_ESTATE, "set_images and set_chunks must be called first");
    if (n_dest < 1 || slab_rows < 1 || stride < 1 || (int64_t)n_dest * slab_rows < c->img.s.n)
        return fail(c, MMALIGN_EINVAL, "mmalign_export_lists: n_dest x slab_rows must cover the %lld image rows", (long long)c->img.s.n);
    if (!is_device_ptr(keys) || !is_device_ptr(count) || !is_device_ptr(tau))
        return fail(c, MMALIGN_EINVAL, "mmalign_export_lists writes device buffers (they feed an all-to-all)");
    CU(c, cudaSetDevice(c->device));
    CandLists L = c->lists_valid ? c->lists : CandLists();  // an empty shard: no lists, complete above -inf
    CU(c, launch_export_lists(L, c->img.s.n, n_dest, slab_rows, stride, c->lists_col0, keys, count, tau, (cudaStream_t)stream));
    return MMALIGN_OK;
}

// ---------------------------------------------------------------------------------------------
// Sharded (multi-GPU) run in passes; the caller performs the collectives in between (distributed.py).
// ---------------------------------------------------------------------------------------------
extern "C" int mmalign_fused_pass(mmalign_ctx *c, const mmalign_params *prm, float *tau_row, void *stream)
{
    if (!c || !prm) return fail(c, MMALIGN_EINVAL, "mmalign_fused_pass: NULL argument");
    int rc = ensure_index(c);
    if (rc) return rc;
    RunParams rp;
    if ((rc = parse_params(c, prm, &rp))) return rc;
    if (rp.candidates != MMALIGN_CAND_ALL) return fail(c, MMALIGN_EINVAL, "mmalign_fused_pass is for MMALIGN_CAND_ALL");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    const int64_t N = c->img.s.n;
    int64_t col0, n_cols;
    if ((rc = resolve_range(c, prm->shard_col0, prm->shard_cols, c->chk.s.n, "shard columns", &col0, &n_cols))) return rc;
    if (prm->shard_col0 != 0 && prm->shard_cols == 0) n_cols = 0;  // an empty shard at the end of the table
    Stager sg{c, st};
    float *d_tau = nullptr;
    sg.map(tau_row, (size_t)N, &d_tau);
    if ((rc = sg.commit())) return rc;
    c->last_fused_us = 0;
    if (N == 0) return MMALIGN_OK;
    if (n_cols == 0) {  // an empty shard holds every one of its (zero) columns
        if (d_tau) {
            std::vector<float> inf((size_t)N, -INFINITY);
            CU(c, cudaMemcpyAsync(d_tau, inf.data(), sizeof(float) * N, cudaMemcpyHostToDevice, st));
            CU(c, cudaStreamSynchronize(st));
        }
        c->lists_valid = false;
        c->lists_col0 = col0;
    } else {
        FusedPlan plan;
        CU(c, cudaEventRecord(c->ev[0], st));
        if ((rc = run_fused(c, rp, prm->kprime, prm->n_ranks > 1 ? prm->n_ranks : 1, 0, N, col0, n_cols, st, &plan))) return rc;
        CU(c, cudaEventRecord(c->ev[1], st));
        if (d_tau) CU(c, launch_row_tau(c->lists, N, d_tau, st));
    }
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    if (n_cols > 0) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        c->last_fused_us = (int64_t)(ms * 1000.f);
    }
    return MMALIGN_OK;
}

static int device_outputs_only(mmalign_ctx *c, const mmalign_out *uo, const char *what)
{
    const void *ptrs[] = {uo->topk_idx, uo->topk_score, uo->pair_rank, uo->pair_sim, uo->pair_score, uo->deep_idx, uo->deep_score};
    for (const void *p : ptrs)
        if (p && !is_device_ptr(p)) return fail(c, MMALIGN_EINVAL, "%s takes device output pointers (it runs between collectives)", what);
    if (uo->hits || uo->rr_sum || uo->sim_sum) return fail(c, MMALIGN_EINVAL, "%s does not reduce metrics; use mmalign_reduce_metrics after the rank exchange", what);
    return MMALIGN_OK;
}

static void outputs_from(const mmalign_out *uo, Outputs *out)
{
    *out = Outputs();
    out->topk_idx = uo->topk_idx; out->topk_score = uo->topk_score; out->pair_rank = uo->pair_rank;
    out->pair_sim = uo->pair_sim; out->pair_score = uo->pair_score; out->deep_idx = uo->deep_idx; out->deep_score = uo->deep_score;
}

extern "C" int mmalign_rescore_pass(mmalign_ctx *c, const mmalign_params *prm, const float *tau_global,
                                    float eps_chunk_global, mmalign_out *uo, int32_t *cert_count, void *stream)
{
    if (!c || !prm || !uo || !tau_global || !cert_count) return fail(c, MMALIGN_EINVAL, "mmalign_rescore_pass: NULL argument");
    int rc = ensure_index(c);
    if (rc) return rc;
    RunParams rp;
    if ((rc = parse_params(c, prm, &rp))) return rc;
    if ((rc = device_outputs_only(c, uo, "mmalign_rescore_pass"))) return rc;
    if (!is_device_ptr(tau_global) || !is_device_ptr(cert_count)) return fail(c, MMALIGN_EINVAL, "mmalign_rescore_pass: tau_global and cert_count must be device pointers");
    const int64_t N = c->img.s.n, M = c->chk.s.n, P = c->px.P;
    if (M > 0 && !c->lists_valid) return fail(c, MMALIGN_ESTATE, "mmalign_fused_pass must run before mmalign_rescore_pass");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    Outputs out;
    outputs_from(uo, &out);
    int32_t *error_flag = (int32_t *)((char *)c->small.p + 16);
    unsigned long long *cand_counter = (unsigned long long *)((char *)c->small.p + 8);
    float *eps_dev = (float *)((char *)c->small.p + 128);
    CU(c, cudaMemsetAsync(c->small.p, 0, 64, st));
    CU(c, cudaMemcpyAsync(eps_dev, &eps_chunk_global, sizeof(float), cudaMemcpyHostToDevice, st));
    if (out.pair_rank && P) CU(c, cudaMemsetAsync(out.pair_rank, 0, (size_t)rp.S * P * sizeof(int32_t), st));
    if (N > 0) {
        if (M > 0) {
            CU(c, launch_rescore(c->img.s, c->chk.s, c->px, rp, &c->lists, eps_dev, out, nullptr, nullptr, nullptr, cand_counter,
                                 error_flag, tau_global, cert_count, RowRange(), st));
        } else {  // empty shard: no entries, nothing certified here
            CU(c, cudaMemsetAsync(cert_count, 0, (size_t)rp.S * N * sizeof(int32_t), st));
            CU(c, launch_exact_scan(c->img.s, c->chk.s, c->px, rp, nullptr, nullptr, N, out, error_flag, RowRange(), nullptr, st));
        }
    }
    CU(c, cudaEventRecord(c->ev[2], st));
    struct { int32_t fail; int32_t pad; unsigned long long cand; int32_t err; } h = {};
    CU(c, cudaMemcpyAsync(&h, c->small.p, 24, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    if (h.err) return fail(c, MMALIGN_ELIMIT, "an image has more than 512 same-page chunks, or a row's lists exceed the rescoring capacity");
    if (uo->stats && !is_device_ptr(uo->stats)) {
        float t_fused = 0.f, t_resc = 0.f;
        if (M > 0) { cudaEventElapsedTime(&t_fused, c->ev[0], c->ev[1]); cudaEventElapsedTime(&t_resc, c->ev[1], c->ev[2]); }
        const int64_t stats[8] = {0, (int64_t)h.cand, M > 0, 4, c->lists.kprime, (int64_t)(t_fused * 1000.f), (int64_t)(t_resc * 1000.f), 0};
        memcpy(uo->stats, stats, sizeof stats);
    }
    return MMALIGN_OK;
}

extern "C" int mmalign_rescan_rows(mmalign_ctx *c, const mmalign_params *prm, const int32_t *rows, int64_t n_rows,
                                   mmalign_out *uo, void *stream)
{
    if (!c || !prm || !uo || (n_rows > 0 && !rows)) return fail(c, MMALIGN_EINVAL, "mmalign_rescan_rows: NULL argument");
    int rc = ensure_index(c);
    if (rc) return rc;
    RunParams rp;
    if ((rc = parse_params(c, prm, &rp))) return rc;
    if ((rc = device_outputs_only(c, uo, "mmalign_rescan_rows"))) return rc;
    if (n_rows == 0) return MMALIGN_OK;
    if (!is_device_ptr(rows)) return fail(c, MMALIGN_EINVAL, "mmalign_rescan_rows: rows must be a device pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    Outputs out;
    outputs_from(uo, &out);
    int32_t *error_flag = (int32_t *)((char *)c->small.p + 16);
    CU(c, cudaMemsetAsync(c->small.p, 0, 64, st));
    CU(c, launch_exact_scan(c->img.s, c->chk.s, c->px, rp, rows, nullptr, n_rows, out, error_flag, RowRange(), nullptr, st));
    int32_t err = 0;
    CU(c, cudaMemcpyAsync(&err, error_flag, sizeof err, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    if (err) return fail(c, MMALIGN_ELIMIT, "an image has more than 512 same-page chunks");
    return MMALIGN_OK;
}

extern "C" int mmalign_chunk_err_max(mmalign_ctx *c, float *err_max)
{
    if (!c || !err_max) return fail(c, MMALIGN_EINVAL, "mmalign_chunk_err_max: NULL argument");
    if (!c->chk.ready) return fail(c, MMALIGN_ESTATE, "set_chunks must be called first");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpy(err_max, c->chk.err_max, sizeof(float), cudaMemcpyDeviceToHost));
    return MMALIGN_OK;
}

extern "C" int mmalign_alignments(mmalign_ctx *c, uint32_t schema, double *rec, void *stream)
{
    if (!c || !rec) return fail(c, MMALIGN_EINVAL, "mmalign_alignments: NULL argument");
    const bool raw = (schema & MMALIGN_RAW_SCORES) != 0;
    const int s = schema_index(schema & ~MMALIGN_RAW_SCORES);
    if (s < 0) return fail(c, MMALIGN_EINVAL, "schema=0x%x must be exactly one MMALIGN_* schema bit", schema);
    int rc = ensure_index(c);
    if (rc) return rc;
    if ((s == 1 || s == 3) && !c->chk.s.terms && c->chk.s.n > 0)
        return fail(c, MMALIGN_EINVAL, "lexical schema requested but the chunks have no term sets");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    Stager sg{c, st};
    double *d = nullptr;
    sg.map(rec, (size_t)c->px.P * 3, &d);
    if ((rc = sg.commit())) return rc;
    if (c->px.P > 0) CU(c, launch_alignments(c->img.s, c->chk.s, c->px, s, c->n_terms, raw, d, st));
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}

extern "C" int mmalign_merge_topk(mmalign_ctx *c, const int64_t *in_idx, const double *in_score, int32_t G,
                                  int64_t n_lists, int32_t K, int64_t *out_idx, double *out_score, void *stream)
{
    if (!c || !in_idx || !in_score || !out_idx || !out_score) return fail(c, MMALIGN_EINVAL, "mmalign_merge_topk: NULL argument");
    if (G < 1 || G > 16 || K < 1 || n_lists < 0) return fail(c, MMALIGN_EINVAL, "mmalign_merge_topk: bad G/K/n_lists");
    if (!is_device_ptr(in_idx) || !is_device_ptr(in_score) || !is_device_ptr(out_idx) || !is_device_ptr(out_score))
        return fail(c, MMALIGN_EINVAL, "mmalign_merge_topk takes device pointers (it runs between two collectives)");
    CU(c, cudaSetDevice(c->device));
    CU(c, launch_merge_topk(in_idx, in_score, G, n_lists, K, out_idx, out_score, (cudaStream_t)stream));
    return MMALIGN_OK;
}

extern "C" int mmalign_count_beating(mmalign_ctx *c, const int64_t *deep_idx, const double *deep_score, int64_t N,
                                     int32_t S, int32_t K, int64_t n_q, const int64_t *q_image,
                                     const int64_t *q_chunk, const double *q_score, int32_t *counts, void *stream)
{
    if (!c || !deep_idx || !deep_score || (n_q > 0 && (!q_image || !q_chunk || !q_score || !counts)))
        return fail(c, MMALIGN_EINVAL, "mmalign_count_beating: NULL argument");
    if (S < 1 || S > 4 || K < 1 || N < 0 || n_q < 0) return fail(c, MMALIGN_EINVAL, "mmalign_count_beating: bad S/K/N/n_q");
    CU(c, cudaSetDevice(c->device));
    CU(c, launch_count_beating(deep_idx, deep_score, N, S, K, n_q, q_image, q_chunk, q_score, counts, (cudaStream_t)stream));
    return MMALIGN_OK;
}

extern "C" int mmalign_reduce_metrics(mmalign_ctx *c, const int32_t *pair_rank, const double *pair_sim, int32_t S,
                                      int64_t P, const int32_t *k_list, int32_t n_k, int32_t mrr_cutoff,
                                      int64_t *hits, double *rr_sum, double *sim_sum, void *stream)
{
    if (!c || !pair_rank || !k_list) return fail(c, MMALIGN_EINVAL, "mmalign_reduce_metrics: NULL argument");
    if (S < 1 || S > 4 || n_k < 1 || n_k > kMaxK || P < 0) return fail(c, MMALIGN_EINVAL, "mmalign_reduce_metrics: bad S/n_k/P");
    if (!is_device_ptr(pair_rank) || (pair_sim && !is_device_ptr(pair_sim)))
        return fail(c, MMALIGN_EINVAL, "mmalign_reduce_metrics: pair arrays must be device pointers");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    Stager sg{c, st};
    int64_t *d_hits = nullptr;
    double *d_rr = nullptr, *d_sim = nullptr;
    sg.map(hits, (size_t)S * n_k, &d_hits);
    sg.map(rr_sum, (size_t)S, &d_rr);
    sg.map(sim_sum, 1, &d_sim);
    int rc;
    if ((rc = sg.commit())) return rc;
    int32_t *k_list_dev = (int32_t *)((char *)c->small.p + 64);
    CU(c, cudaMemcpyAsync(k_list_dev, k_list, sizeof(int32_t) * n_k, cudaMemcpyDefault, st));
    CU(c, c->metrics_scratch.reserve(metrics_scratch_bytes(S, n_k)));
    CU(c, launch_reduce_metrics(pair_rank, pair_sim, S, P, k_list_dev, n_k, mrr_cutoff, d_hits, d_rr, d_sim,
                                c->metrics_scratch.p, st));
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}

extern "C" int mmalign_term_bitsets(mmalign_ctx *c, const uint8_t *text, const int64_t *text_off, int64_t m,
                                    const uint8_t *terms, const int64_t *term_off, int32_t n_terms, int32_t term_words,
                                    uint64_t *bits, void *stream)
{
    if (!c || !text_off || !bits || m < 0 || (n_terms > 0 && (!terms || !term_off)))
        return fail(c, MMALIGN_EINVAL, "mmalign_term_bitsets: NULL argument");
    if (n_terms < 0 || n_terms > 4096) return fail(c, MMALIGN_ELIMIT, "mmalign_term_bitsets: n_terms=%d must be in 0..4096", n_terms);
    if (term_words < 1 || (int64_t)term_words * 64 < n_terms || term_words > 64)
        return fail(c, MMALIGN_EINVAL, "mmalign_term_bitsets: term_words=%d does not hold %d terms (1..64 words)", term_words, n_terms);
    if (is_device_ptr(terms) || is_device_ptr(term_off)) return fail(c, MMALIGN_EINVAL, "mmalign_term_bitsets: terms and term_off are host arrays");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    if (m == 0) return MMALIGN_OK;
    // the term table: tiny, prepared on the host
    TermTableHost h;
    static const int64_t zero_off[1] = {0};
    const int trc = build_term_table(terms, n_terms > 0 ? term_off : zero_off, n_terms, term_words, &h);
    if (trc) return fail(c, MMALIGN_EINVAL, "mmalign_term_bitsets: bad term offsets (code %d)", trc);
    const size_t tb = n_terms > 0 ? (size_t)term_off[n_terms] : 0;
    // one packed upload: term bytes | hash (8-byte slots) | offsets | short-term buckets | groups | empty-term mask
    auto pad16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_hash = pad16(tb), o_off = pad16(o_hash + h.hash.size() * 4), o_bs = pad16(o_off + h.off.size() * 4),
                 o_bt = pad16(o_bs + h.bucket_start.size() * 4), o_gt = pad16(o_bt + h.bucket_term.size() * 4),
                 o_al = pad16(o_gt + h.group_term.size() * 4), total = o_al + h.always.size() * 4;
    std::vector<char> packed(total, 0);
    if (tb) memcpy(packed.data(), terms, tb);
    memcpy(packed.data() + o_hash, h.hash.data(), h.hash.size() * 4);
    memcpy(packed.data() + o_off, h.off.data(), h.off.size() * 4);
    memcpy(packed.data() + o_bs, h.bucket_start.data(), h.bucket_start.size() * 4);
    memcpy(packed.data() + o_bt, h.bucket_term.data(), h.bucket_term.size() * 4);
    memcpy(packed.data() + o_gt, h.group_term.data(), h.group_term.size() * 4);
    memcpy(packed.data() + o_al, h.always.data(), h.always.size() * 4);
    CU(c, c->term_table.reserve(total));
    char *d = (char *)c->term_table.p;
    CU(c, cudaMemcpyAsync(d, packed.data(), total, cudaMemcpyHostToDevice, st));
    CU(c, cudaStreamSynchronize(st));  // `packed` is pageable host memory that goes out of scope
    // the texts
    const int64_t *d_off = text_off;
    int64_t total_text = 0;
    if (is_device_ptr(text_off)) {
        CU(c, cudaMemcpyAsync(&total_text, text_off + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
    } else {
        total_text = text_off[m];
        CU(c, c->text_off.reserve(sizeof(int64_t) * (m + 1)));
        CU(c, cudaMemcpyAsync(c->text_off.p, text_off, sizeof(int64_t) * (m + 1), cudaMemcpyHostToDevice, st));
        d_off = (const int64_t *)c->text_off.p;
    }
    if (total_text < 0 || (total_text > 0 && !text)) return fail(c, MMALIGN_EINVAL, "mmalign_term_bitsets: bad text offsets");
    const uint8_t *d_text = text;
    if (total_text > 0 && !is_device_ptr(text)) {
        CU(c, c->text_bytes.reserve((size_t)total_text));
        CU(c, cudaMemcpyAsync(c->text_bytes.p, text, (size_t)total_text, cudaMemcpyHostToDevice, st));
        d_text = (const uint8_t *)c->text_bytes.p;
    }
    Stager sg{c, st};
    uint64_t *d_bits = nullptr;
    sg.map(bits, (size_t)m * term_words, &d_bits);
    int rc;
    if ((rc = sg.commit())) return rc;
    CU(c, launch_term_bitsets(d_text, d_off, m, (const uint8_t *)d, (const int32_t *)(d + o_off), (const int32_t *)(d + o_bs),
                              (const int32_t *)(d + o_bt), (const uint32_t *)(d + o_hash), h.hash_bits,
                              (const int32_t *)(d + o_gt), (const uint32_t *)(d + o_al), term_words, d_bits, st));
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}

extern "C" int mmalign_debug_scores(mmalign_ctx *c, float *out, void *stream)
{
    if (!c || !out) return fail(c, MMALIGN_EINVAL, "mmalign_debug_scores: NULL argument");
    int rc = ensure_index(c);
    if (rc) return rc;
    const Side &img = c->img.s, &chk = c->chk.s;
    if (img.n == 0 || chk.n == 0) return MMALIGN_OK;
    if (img.D % 64 != 0) return fail(c, MMALIGN_EINVAL, "the fused kernel needs D %% 64 == 0");
    if ((double)img.n * (double)chk.n > 2.7e8) return fail(c, MMALIGN_ELIMIT, "mmalign_debug_scores is for small problems (N*M <= 2.7e8)");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    FusedPlan plan;
    if (fused_plan(img.n, chk.n, img.D, 10, 0, c->sm_count, 1, &plan)) return fail(c, MMALIGN_ELIMIT, "no fused plan");
    Stager sg{c, st};
    float *d = nullptr;
    sg.map(out, (size_t)img.n * chk.n, &d);
    if ((rc = sg.commit())) return rc;
    CandLists L;
    CU(c, launch_fused(img, chk, plan, &c->img.tmap, &c->chk.tmap, L, d, st));
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}
