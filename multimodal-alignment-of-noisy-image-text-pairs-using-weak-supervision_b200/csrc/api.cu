// api.cu -- the C ABI of include/mmalign.h: context, uploads, and the
// K0 -> K1 -> K2 -> exact scan -> K4 pipeline.  No CPU fallback anywhere: every
// result is produced by the kernels in prep.cu / fused_tc.cu / rescore.cu.
#include "common.cuh"
#include <algorithm>
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
    explicit Trace(const char *w) : on(enabled()), t0(std::chrono::steady_clock::now()), what(w) {}
    static bool enabled() { static const bool e = getenv("MMALIGN_TRACE") != nullptr; return e; }
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

constexpr int kMaxGroups = 4;

struct SideStore {
    Side s;
    DevBuf up[4];               // uploads of host inputs: emb, key, bbox, terms (reused across set_* calls)
    DevBuf bf16, norm2, err, errmax;
    float *err_max = nullptr;   // [1] max rounding-error norm over the rows (chunks)
    alignas(64) CUtensorMap tmap;
    alignas(64) CUtensorMap tmap_half;  // chunks: 128-row boxes (the B halves of the CTA-pair kernel)
    bool ready = false;
    // asynchronous ingest: host embedding rows arrive in pieces on the context's copy stream
    int64_t piece_rows = 0;
    int n_pieces = 0;                    // 0 = the rows were on the device already
    std::vector<cudaEvent_t> ev_piece;   // [>= n_pieces] piece p is on the device (recorded on the copy stream)
    const float *pending_emb = nullptr;  // page-locked host rows whose pieces [pieces_queued, n_pieces) are not queued yet
    int pieces_queued = 0;
    // chunks: the table in column groups (whole pieces), each announcing its K0 -- the first slab of a run is
    // contracted group by group while the later groups still travel
    int n_groups = 1;
    int64_t group_row[kMaxGroups + 1] = {};
    cudaEvent_t ev_group[kMaxGroups] = {};
    cudaEvent_t ev_small = nullptr;      // keys / boxes / term sets are on the device (copy stream)
    cudaEvent_t ev_caller = nullptr;     // whatever the caller had queued on the legacy default stream at set_* time
    cudaEvent_t ev_ready = nullptr;      // the K0 launches queued so far have run (chunks: all rows, at set_* time)
    bool ev_ready_set = false;
    int64_t prep_lo = 0, prep_hi = 0;    // rows whose K0 outputs (bf16, norm2, err) are queued or done
    void release()
    {
        for (DevBuf &b : up) b.release();
        bf16.release(); norm2.release(); err.release(); errmax.release();
        s = Side();
        err_max = nullptr;
        ready = false;
        n_pieces = 0; prep_lo = prep_hi = 0; ev_ready_set = false;
        pending_emb = nullptr; pieces_queued = 0; n_groups = 1;
    }
    void destroy_events()
    {
        for (cudaEvent_t e : ev_piece) if (e) cudaEventDestroy(e);
        ev_piece.clear();
        for (cudaEvent_t *e : {&ev_small, &ev_caller, &ev_ready, &ev_group[0], &ev_group[1], &ev_group[2], &ev_group[3]}) {
            if (*e) cudaEventDestroy(*e);
            *e = nullptr;
        }
    }
};

constexpr int kMaxSlabs = 512;          // slabs of one mmalign_run (their failure counters live in `small`)
constexpr int kSlabWaves = 4;           // auto slab = this many waves of 128-row blocks over the SMs
constexpr size_t kSmallBytes = 8192;    // 0: fail_count | 8: cand_counter | 16: error_flag | 24: eps violations | 64: k_list | 128: eps | 1024: slab counters | 4096: slab counters of the rows left to the block-per-row rescoring

struct mmalign_ctx {
    int device = 0;
    int sm_count = 148;
    char err[512] = "";
    SideStore img, chk;
    int64_t n_terms = 0, col_offset = 0;
    PairIndex px;
    bool px_ready = false;
    bool chk_consumed = true;      // a run has read the chunk table since the last set_chunks
    int k2_sms = 0;                // SMs left to the exact rescoring of slab s while slab s+1 is contracted (0 = no overlap)
    int epi_sleep_ns = 0;          // mmalign_set_option: pause between polls of the epilogue's accumulator barrier
    int compact_one = 0;           // mmalign_set_option: routine list compaction all flagged lists at once (0, default), one list per tile gap (1),
                                   // split over two gaps (2: loads under the next tile's filtering) -- measured alike
    int cta_pairs = 0;             // the fused kernel on CTA pairs: 1 = tcgen05.mma.cta_group::2, 2 = B multicast (mmalign_set_option)
    size_t piece_bytes = (size_t)64 << 20;  // host embedding rows travel in pieces of about this size (mmalign_set_option)
    DevBuf px_offsets, px_sorted, px_start, px_scratch;
    DevBuf list_keys, list_tau, list_count;
    DevBuf list_keys2, list_tau2, list_count2;  // second set: slab s+1 is contracted while slab s is re-scored
    DevBuf fail_rows, fail_thr, scan_buf, scan_cnt, small;
    DevBuf big_rows, k2_scratch;   // rows the warp-per-row rescoring hands to the block-per-row kernel; its per-row records
    DevBuf metrics_scratch, stage; // stage: device copies of host outputs
    DevBuf term_table, text_off, text_bytes;  // mmalign_term_bitsets: term table, uploads of host texts
    DevBuf copy_off, copy_len, half_up;       // mmalign_copy_decode: field tables; set_*_half: upload of host halves
    CandLists lists;               // written by the last fused pass
    bool lists_valid = false;
    int64_t lists_col0 = 0;        // first chunk row of the column range the lists were built on
    int64_t lists_rows = 0, lists_cols = 0;  // image rows / chunk columns the lists cover
    int64_t last_fused_us = 0;     // CUDA-event time of the last mmalign_fused_pass
    cudaEvent_t ev[5] = {};        // fused pass / rescore pass timing
    // streams of the context (non-blocking): uploads, chunk preparation, pair index, downloads
    cudaStream_t s_in = nullptr, s_prep = nullptr, s_idx = nullptr, s_out = nullptr;
    cudaStream_t s_k1 = nullptr, s_k2 = nullptr;  // contraction (high priority) and rescoring when the two overlap
    cudaEvent_t ev_tmp = nullptr;
    std::vector<cudaEvent_t> ev_slab;  // 5 per slab: start, after fused, after rescore, after exact scan (= slab done), rescore start
    cudaEvent_t resc_wait = nullptr;   // mmalign_rescore_after: one-shot
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

enum PtrKind { kPtrNull, kPtrDevice, kPtrPinned, kPtrPageable };
static PtrKind ptr_kind(const void *p)
{
    if (!p) return kPtrNull;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return kPtrPageable; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return kPtrDevice;
    return a.type == cudaMemoryTypeHost ? kPtrPinned : kPtrPageable;
}
static bool is_device_ptr(const void *p) { return ptr_kind(p) == kPtrDevice; }

extern "C" int mmalign_abi_version(void) { return MMALIGN_ABI_VERSION; }

extern "C" const char *mmalign_last_error(const mmalign_ctx *ctx) { return ctx ? ctx->err : g_err; }

extern "C" void mmalign_destroy(mmalign_ctx *c);

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
    bool ok = c->small.reserve(kSmallBytes) == cudaSuccess;
    for (cudaEvent_t &ev : c->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    for (cudaStream_t *s : {&c->s_in, &c->s_prep, &c->s_idx, &c->s_out, &c->s_k2})
        ok = ok && cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) == cudaSuccess;
    {   // the contraction's CTAs are dispatched before the rescoring's when both kernels are ready
        int least = 0, greatest = 0;
        ok = ok && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&c->s_k1, cudaStreamNonBlocking, greatest) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&c->ev_tmp, cudaEventDisableTiming) == cudaSuccess;
    for (SideStore *ss : {&c->img, &c->chk})
        for (cudaEvent_t *ev : {&ss->ev_small, &ss->ev_caller, &ss->ev_ready, &ss->ev_group[0], &ss->ev_group[1],
                                &ss->ev_group[2], &ss->ev_group[3]})
            ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { mmalign_destroy(c); return fail(nullptr, MMALIGN_ECUDA, "creating the context's streams / events / scratch failed"); }
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
    c->img.destroy_events();
    c->chk.destroy_events();
    DevBuf *bufs[] = {&c->px_offsets, &c->px_sorted, &c->px_start, &c->px_scratch, &c->list_keys, &c->list_tau, &c->list_count,
                      &c->list_keys2, &c->list_tau2, &c->list_count2, &c->fail_rows, &c->fail_thr, &c->big_rows, &c->k2_scratch, &c->scan_buf, &c->scan_cnt, &c->small, &c->metrics_scratch, &c->stage,
                      &c->term_table, &c->text_off, &c->text_bytes, &c->copy_off, &c->copy_len, &c->half_up};
    for (DevBuf *b : bufs) b->release();
    for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_slab) if (e) cudaEventDestroy(e);
    if (c->ev_tmp) cudaEventDestroy(c->ev_tmp);
    for (cudaStream_t s : {c->s_in, c->s_prep, c->s_idx, c->s_out, c->s_k1, c->s_k2}) if (s) cudaStreamDestroy(s);
    delete c;
}

extern "C" int mmalign_set_option(mmalign_ctx *c, const char *name, int64_t value)
{
    if (!c || !name) return fail(c, MMALIGN_EINVAL, "mmalign_set_option: NULL argument");
    if (!strcmp(name, "piece_bytes")) {
        if (value < 1024 || value > ((int64_t)1 << 34)) return fail(c, MMALIGN_EINVAL, "piece_bytes=%lld must be in 1 KiB..16 GiB", (long long)value);
        c->piece_bytes = (size_t)value;
        return MMALIGN_OK;
    }
    if (!strcmp(name, "k2_sms")) {
        if (value < 0 || value > c->sm_count / 2) return fail(c, MMALIGN_EINVAL, "k2_sms=%lld must be in 0..%d", (long long)value, c->sm_count / 2);
        c->k2_sms = (int)value;
        return MMALIGN_OK;
    }
    if (!strcmp(name, "compact_one")) {
        if (value < 0 || value > 2) return fail(c, MMALIGN_EINVAL, "compact_one=%lld must be 0, 1 or 2", (long long)value);
        c->compact_one = (int)value;
        return MMALIGN_OK;
    }
    if (!strcmp(name, "epi_sleep_ns")) {
        if (value < 0 || value > 100000) return fail(c, MMALIGN_EINVAL, "epi_sleep_ns=%lld must be in 0..100000", (long long)value);
        c->epi_sleep_ns = (int)value;
        return MMALIGN_OK;
    }
    if (!strcmp(name, "cta_pairs")) {
        if (value < 0 || value > 2) return fail(c, MMALIGN_EINVAL, "cta_pairs=%lld must be 0, 1 or 2", (long long)value);
        c->cta_pairs = (int)value;
        return MMALIGN_OK;
    }
    return fail(c, MMALIGN_EINVAL, "mmalign_set_option: unknown option '%s'", name);
}

extern "C" int mmalign_sync(mmalign_ctx *c)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_sync: ctx is NULL");
    CU(c, cudaSetDevice(c->device));
    for (cudaStream_t s : {c->s_in, c->s_prep, c->s_idx, c->s_out}) CU(c, cudaStreamSynchronize(s));
    return MMALIGN_OK;
}

// host array -> the side's upload buffer (copy stream); device array: borrowed
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

// `b` waits for what `a` has queued so far
static int order_after(mmalign_ctx *c, cudaStream_t b, cudaStream_t a)
{
    CU(c, cudaEventRecord(c->ev_tmp, a));
    CU(c, cudaStreamWaitEvent(b, c->ev_tmp, 0));
    return MMALIGN_OK;
}


// queues the uploads of pieces [pieces_queued, upto) of a side's host embedding rows on the copy stream
static int queue_pieces(mmalign_ctx *c, SideStore &ss, int upto)
{
    if (!ss.pending_emb) return MMALIGN_OK;
    if (upto > ss.n_pieces) upto = ss.n_pieces;
    const int64_t n = ss.s.n, D = ss.s.D, pr = ss.piece_rows;
    for (int p = ss.pieces_queued; p < upto; ++p) {
        const int64_t r0 = (int64_t)p * pr, r1 = r0 + pr < n ? r0 + pr : n;
        CU(c, cudaMemcpyAsync((float *)ss.up[0].p + r0 * D, ss.pending_emb + r0 * D, (size_t)(r1 - r0) * D * sizeof(float),
                              cudaMemcpyHostToDevice, c->s_in));
        CU(c, cudaEventRecord(ss.ev_piece[p], c->s_in));
    }
    if (upto > ss.pieces_queued) ss.pieces_queued = upto;
    if (ss.pieces_queued >= ss.n_pieces) ss.pending_emb = nullptr;
    return MMALIGN_OK;
}

static int set_side(mmalign_ctx *c, SideStore &ss, const float *emb, const uint64_t *key, const double *bbox,
                    const uint64_t *terms, int64_t n, int D, int term_words, int box_rows, const char *what, bool is_chunks)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "%s: ctx is NULL", what);
    CU(c, cudaSetDevice(c->device));
    if (n < 0 || n > 0x7FFFFFF0ll) return fail(c, MMALIGN_EINVAL, "%s: row count %lld out of range", what, (long long)n);
    if (D <= 0 || D % 4 != 0 || D > 4096) return fail(c, MMALIGN_EINVAL, "%s: D=%d must be a multiple of 4 in 4..4096", what, D);
    if (n > 0 && (!emb || !key)) return fail(c, MMALIGN_EINVAL, "%s: emb and page_key are required", what);
    if (term_words < 0 || (terms && term_words == 0)) return fail(c, MMALIGN_EINVAL, "%s: bad term_words", what);
    Trace tr(what);
    int rc;
    // the side's buffers may still be read by preparation / index work queued by an earlier set_* call
    if ((rc = order_after(c, c->s_in, c->s_prep))) return rc;
    if ((rc = order_after(c, c->s_in, c->s_idx))) return rc;
    CU(c, cudaEventRecord(ss.ev_caller, 0));  // device inputs: complete on the legacy default stream (include/mmalign.h)
    ss.ready = false;
    c->px_ready = false;
    c->lists_valid = false;
    ss.s = Side();
    Side &s = ss.s;
    s.n = n; s.D = D; s.term_words = term_words;
    cudaStream_t sin = c->s_in;
    if ((rc = adopt(c, ss.up[1], key, (size_t)n, &s.key, sin))) return rc;
    if ((rc = adopt(c, ss.up[3], terms, (size_t)n * term_words, &s.terms, sin))) return rc;
    if (bbox) {
        if ((rc = adopt(c, ss.up[2], bbox, (size_t)n * 4, &s.bbox, sin))) return rc;
    } else if (n > 0) {  // missing boxes: all zero -> positional score 0 (insert_clip_embeddings.py:161)
        CU(c, ss.up[2].reserve((size_t)n * 4 * sizeof(double)));
        CU(c, cudaMemsetAsync(ss.up[2].p, 0, (size_t)n * 4 * sizeof(double), sin));
        s.bbox = static_cast<const double *>(ss.up[2].p);
    }
    CU(c, cudaEventRecord(ss.ev_small, sin));
    const size_t nn = n > 0 ? (size_t)n : 1;
    CU(c, ss.bf16.reserve(nn * D * sizeof(__nv_bfloat16))); s.emb_bf16 = (__nv_bfloat16 *)ss.bf16.p;
    CU(c, ss.norm2.reserve(nn * sizeof(float))); s.norm2 = (float *)ss.norm2.p;
    CU(c, ss.err.reserve(nn * sizeof(float))); s.err = (float *)ss.err.p;
    CU(c, ss.errmax.reserve(sizeof(float))); ss.err_max = (float *)ss.errmax.p;
    // embedding rows: borrowed on the device, or uploaded piece by piece (each piece announces itself with an event).
    // Order on the copy stream: the chunk table goes first (the contraction of any query slab needs all of it),
    // preceded only by the image pieces of the first query slab; a page-locked image table set BEFORE the chunks
    // therefore waits for set_chunks (or for its first consumer) to queue its pieces.  Pageable rows cannot wait.
    ss.n_pieces = 0;
    ss.pending_emb = nullptr;
    ss.pieces_queued = 0;
    ss.prep_lo = ss.prep_hi = 0;
    ss.ev_ready_set = false;
    ss.n_groups = 1;
    if (n > 0 && is_device_ptr(emb)) {
        s.emb = emb;
    } else if (n > 0) {
        CU(c, ss.up[0].reserve((size_t)n * D * sizeof(float)));
        s.emb = static_cast<const float *>(ss.up[0].p);
        int64_t pr = (int64_t)(c->piece_bytes / ((size_t)D * sizeof(float)));
        pr = pr < 256 ? 256 : pr / 256 * 256;  // whole column tiles of the fused kernel
        ss.piece_rows = pr;
        ss.n_pieces = (int)((n + pr - 1) / pr);
        while ((int)ss.ev_piece.size() < ss.n_pieces) {
            cudaEvent_t e = nullptr;
            CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ss.ev_piece.push_back(e);
        }
        ss.pending_emb = emb;
        const bool pinned = ptr_kind(emb) == kPtrPinned;
        if (!is_chunks) {
            if (!pinned || (c->chk.ready && !c->chk_consumed)) { if ((rc = queue_pieces(c, ss, ss.n_pieces))) return rc; }
        } else {
            SideStore &im = c->img;
            if (im.pending_emb) {  // the first query slab's rows, then the chunks, then the other image rows
                const int64_t first = (int64_t)kSlabWaves * c->sm_count * 128;
                if ((rc = queue_pieces(c, im, (int)((first + im.piece_rows - 1) / im.piece_rows)))) return rc;
            }
            if ((rc = queue_pieces(c, ss, ss.n_pieces))) return rc;
            if ((rc = queue_pieces(c, im, im.n_pieces))) return rc;
        }
    }
    if (is_chunks && c->img.pending_emb && ss.n_pieces == 0)
        if ((rc = queue_pieces(c, c->img, c->img.n_pieces))) return rc;
    if (n > 0 && D % 64 == 0) {
        char msg[256];
        if (encode_tensor_map(&ss.tmap, s.emb_bf16, n, D, box_rows, msg, sizeof msg) ||
            (is_chunks && encode_tensor_map(&ss.tmap_half, s.emb_bf16, n, D, 128, msg, sizeof msg)))
            return fail(c, MMALIGN_ECUDA, "%s: %s", what, msg);
    }
    // chunks: K0 runs now, piece by piece behind the uploads.  Images: K0 runs in the consumer (mmalign_run prepares
    // slab s just before it contracts it, so that the upload of later slabs overlaps the kernels of earlier ones).
    if (is_chunks) {
        cudaStream_t sp = c->s_prep;
        CU(c, cudaStreamWaitEvent(sp, ss.ev_caller, 0));
        ss.group_row[0] = 0; ss.group_row[1] = n;
        if (ss.n_pieces == 0) {
            CU(c, launch_prep(s, 0, n, c->sm_count, sp));
        } else {
            ss.n_groups = ss.n_pieces >= 2 * kMaxGroups ? kMaxGroups : 1;
            int g = 0;
            for (int p = 0; p < ss.n_pieces; ++p) {
                const int64_t r0 = (int64_t)p * ss.piece_rows, r1 = r0 + ss.piece_rows < n ? r0 + ss.piece_rows : n;
                CU(c, cudaStreamWaitEvent(sp, ss.ev_piece[p], 0));
                CU(c, launch_prep(s, r0, r1 - r0, c->sm_count, sp));
                if (p + 1 == (g + 1) * ss.n_pieces / ss.n_groups) {  // last piece of group g
                    CU(c, cudaEventRecord(ss.ev_group[g], sp));
                    ss.group_row[++g] = r1;
                }
            }
        }
        CU(c, reduce_max_float(s.err, n, ss.err_max, sp));
        CU(c, cudaEventRecord(ss.ev_ready, sp));
        ss.ev_ready_set = true;
        ss.prep_lo = 0; ss.prep_hi = n;
        c->chk_consumed = false;
    }
    tr.mark("queued upload + prep");
    ss.ready = true;
    return MMALIGN_OK;
}

extern "C" int mmalign_set_images(mmalign_ctx *c, const float *emb, const uint64_t *key, const double *bbox,
                                  const uint64_t *terms, int64_t n, int32_t D, int32_t term_words)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_set_images: ctx is NULL");
    return set_side(c, c->img, emb, key, bbox, terms, n, D, term_words, 128, "mmalign_set_images", false);
}

extern "C" int mmalign_set_chunks(mmalign_ctx *c, const float *emb, const uint64_t *key, const double *bbox,
                                  const uint64_t *terms, int64_t m, int32_t D, int32_t term_words,
                                  int64_t n_terms, int64_t col_offset)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_set_chunks: ctx is NULL");
    if (n_terms < 0 || n_terms > (int64_t)term_words * 64)
        return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks: n_terms=%lld does not fit %d term words", (long long)n_terms, term_words);
    if (col_offset < 0) return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks: negative col_offset");
    int rc = set_side(c, c->chk, emb, key, bbox, terms, m, D, term_words, 256, "mmalign_set_chunks", true);
    if (rc) return rc;
    c->n_terms = n_terms;
    c->col_offset = col_offset;
    return MMALIGN_OK;
}

extern "C" int mmalign_prep_rows(mmalign_ctx *c, const float *emb, int64_t n, int32_t D, void *bf16_out, float *norm2_out,
                                 float *err_out, void *stream)
{
    if (!c || (n > 0 && (!emb || !bf16_out || !norm2_out || !err_out))) return fail(c, MMALIGN_EINVAL, "mmalign_prep_rows: NULL argument");
    if (n < 0 || D <= 0 || D % 4 != 0 || D > 4096) return fail(c, MMALIGN_EINVAL, "mmalign_prep_rows: bad n / D");
    if (n == 0) return MMALIGN_OK;
    if (!is_device_ptr(emb) || !is_device_ptr(bf16_out) || !is_device_ptr(norm2_out) || !is_device_ptr(err_out))
        return fail(c, MMALIGN_EINVAL, "mmalign_prep_rows takes device pointers (it feeds an all-gather)");
    CU(c, cudaSetDevice(c->device));
    Side s;
    s.n = n; s.D = D; s.emb = emb; s.emb_bf16 = (__nv_bfloat16 *)bf16_out; s.norm2 = norm2_out; s.err = err_out;
    CU(c, launch_prep(s, 0, n, c->sm_count, (cudaStream_t)stream));
    return MMALIGN_OK;
}

extern "C" int mmalign_set_chunks_prepared(mmalign_ctx *c, const float *emb, const uint64_t *key, const double *bbox,
                                           const uint64_t *terms, const void *bf16, const float *norm2, const float *err,
                                           int64_t m, int32_t D, int32_t term_words, int64_t n_terms, int64_t col_offset,
                                           void *stream)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_set_chunks_prepared: ctx is NULL");
    CU(c, cudaSetDevice(c->device));
    if (m < 0 || m > 0x7FFFFFF0ll) return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks_prepared: row count out of range");
    if (D <= 0 || D % 4 != 0 || D > 4096) return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks_prepared: D=%d must be a multiple of 4 in 4..4096", D);
    if (n_terms < 0 || n_terms > (int64_t)term_words * 64 || term_words < 0 || (terms && term_words == 0) || col_offset < 0)
        return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks_prepared: bad term_words / n_terms / col_offset");
    if (m > 0) {
        const void *need[] = {emb, key, bbox, bf16, norm2, err};
        for (const void *p : need)
            if (!is_device_ptr(p)) return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks_prepared takes device pointers (emb, page_key, bbox, bf16, norm2, err)");
        if (terms && !is_device_ptr(terms)) return fail(c, MMALIGN_EINVAL, "mmalign_set_chunks_prepared: terms must be a device pointer");
    }
    SideStore &ss = c->chk;
    int rc;
    if ((rc = order_after(c, c->s_in, c->s_prep))) return rc;
    if ((rc = order_after(c, c->s_in, c->s_idx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ss.ready = false; c->px_ready = false; c->lists_valid = false;
    ss.s = Side();
    Side &s = ss.s;
    s.n = m; s.D = D; s.term_words = term_words;
    s.emb = emb; s.key = key; s.bbox = bbox; s.terms = terms;
    s.emb_bf16 = (__nv_bfloat16 *)const_cast<void *>(bf16); s.norm2 = const_cast<float *>(norm2); s.err = const_cast<float *>(err);
    ss.n_pieces = 0; ss.n_groups = 1; ss.pending_emb = nullptr; ss.pieces_queued = 0;
    ss.group_row[0] = 0; ss.group_row[1] = m;
    c->chk_consumed = false;
    if ((rc = queue_pieces(c, c->img, c->img.n_pieces))) return rc;
    CU(c, ss.errmax.reserve(sizeof(float))); ss.err_max = (float *)ss.errmax.p;
    // the caller's stream carries the exchange that filled the prepared operands and the keys
    CU(c, cudaEventRecord(ss.ev_caller, st));
    CU(c, cudaEventRecord(ss.ev_small, st));
    CU(c, cudaStreamWaitEvent(c->s_prep, ss.ev_caller, 0));
    CU(c, reduce_max_float(s.err, m, ss.err_max, c->s_prep));
    CU(c, cudaEventRecord(ss.ev_ready, c->s_prep));
    ss.ev_ready_set = true;
    ss.prep_lo = 0; ss.prep_hi = m;
    if (m > 0 && D % 64 == 0) {
        char msg[256];
        if (encode_tensor_map(&ss.tmap, s.emb_bf16, m, D, 256, msg, sizeof msg) ||
            encode_tensor_map(&ss.tmap_half, s.emb_bf16, m, D, 128, msg, sizeof msg))
            return fail(c, MMALIGN_ECUDA, "mmalign_set_chunks_prepared: %s", msg);
    }
    c->n_terms = n_terms;
    c->col_offset = col_offset;
    ss.ready = true;
    return MMALIGN_OK;
}

extern "C" int mmalign_rescore_after(mmalign_ctx *c, void *event)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_rescore_after: ctx is NULL");
    c->resc_wait = (cudaEvent_t)event;
    return MMALIGN_OK;
}

// `st` waits for the side's small arrays, for the caller's producers and for the K0 launches queued so far
static int wait_side(mmalign_ctx *c, SideStore &ss, cudaStream_t st)
{
    CU(c, cudaStreamWaitEvent(st, ss.ev_small, 0));
    CU(c, cudaStreamWaitEvent(st, ss.ev_caller, 0));
    if (ss.ev_ready_set) CU(c, cudaStreamWaitEvent(st, ss.ev_ready, 0));
    return MMALIGN_OK;
}

// K0 of image rows [lo, hi) on `st` (behind the upload pieces that hold them), unless it is queued already
static int prepare_images(mmalign_ctx *c, int64_t lo, int64_t hi, cudaStream_t st, long long *launched = nullptr)
{
    SideStore &ss = c->img;
    if (lo >= hi) return MMALIGN_OK;
    if (lo >= ss.prep_lo && hi <= ss.prep_hi) return MMALIGN_OK;  // (wait_side ordered `st` behind those launches)
    if (ss.n_pieces > 0) {
        const int p0 = (int)(lo / ss.piece_rows), p1 = (int)((hi - 1) / ss.piece_rows);
        for (int p = p0; p <= p1 && p < ss.n_pieces; ++p) CU(c, cudaStreamWaitEvent(st, ss.ev_piece[p], 0));
    }
    CU(c, launch_prep(ss.s, lo, hi - lo, c->sm_count, st));
    if (launched) *launched += 1;
    CU(c, cudaEventRecord(ss.ev_ready, st));
    ss.ev_ready_set = true;
    if (ss.prep_lo == ss.prep_hi) { ss.prep_lo = lo; ss.prep_hi = hi; }
    else if (lo <= ss.prep_hi && hi >= ss.prep_lo) { ss.prep_lo = lo < ss.prep_lo ? lo : ss.prep_lo; ss.prep_hi = hi > ss.prep_hi ? hi : ss.prep_hi; }
    else { ss.prep_lo = lo; ss.prep_hi = hi; }
    return MMALIGN_OK;
}

// every input of a kernel that reads both tables whole (the sharded passes, debug hooks)
static int wait_tables(mmalign_ctx *c, cudaStream_t st)
{
    int rc;
    if ((rc = queue_pieces(c, c->img, c->img.n_pieces))) return rc;
    c->chk_consumed = true;
    if ((rc = wait_side(c, c->img, st))) return rc;
    if ((rc = wait_side(c, c->chk, st))) return rc;
    return prepare_images(c, 0, c->img.s.n, st);
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
    // the index needs the page keys only: it is built on its own stream while the embedding rows still travel
    for (SideStore *ss : {&c->img, &c->chk}) {
        CU(c, cudaStreamWaitEvent(c->s_idx, ss->ev_small, 0));
        CU(c, cudaStreamWaitEvent(c->s_idx, ss->ev_caller, 0));
    }
    CU(c, build_pair_index(c->img.s, c->chk.s, c->px, c->px_scratch.p, sb, c->s_idx));
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
    enum Kind { kSmall, kPerRow, kPerPair };  // per-row: [S][rows][w]; per-pair: [S][P] (S = 1: [P])
    mmalign_ctx *c;
    cudaStream_t st;
    struct Item { void *host; void *dev; size_t bytes; Kind kind; size_t unit; int S; };
    std::vector<Item> items;
    size_t used = 0;
    std::vector<std::pair<void **, size_t>> pending;  // (slot to patch, offset)
    template <typename T> int map(T *user, size_t count, T **dev, Kind kind = kSmall, size_t unit = 0, int S = 1)
    {
        *dev = nullptr;
        if (!user || count == 0) return 0;
        if (is_device_ptr(user)) { *dev = user; return 0; }
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        items.push_back({user, nullptr, count * sizeof(T), kind, unit, S});
        pending.push_back({(void **)dev, used});
        used += bytes;
        return 0;
    }
    bool any_large() const
    {
        for (auto &it : items) if (it.kind != kSmall) return true;
        return false;
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
    int copy_back(cudaStream_t s, bool small_only = false)
    {
        for (auto &it : items)
            if (!small_only || it.kind == kSmall) CU(c, cudaMemcpyAsync(it.host, it.dev, it.bytes, cudaMemcpyDeviceToHost, s));
        return 0;
    }
    int copy_back() { return copy_back(st); }
    // rows [r0, r1) of N and pairs [p0, p1) of P (positions within the run's output window)
    int copy_slab(cudaStream_t s, int64_t r0, int64_t r1, int64_t N, int64_t p0, int64_t p1, int64_t P)
    {
        for (auto &it : items) {
            if (it.kind == kSmall) continue;
            const int64_t lo = it.kind == kPerRow ? r0 : p0, hi = it.kind == kPerRow ? r1 : p1, tot = it.kind == kPerRow ? N : P;
            if (hi <= lo) continue;
            for (int si = 0; si < it.S; ++si) {
                const size_t off = ((size_t)si * tot + lo) * it.unit;
                CU(c, cudaMemcpyAsync((char *)it.host + off, (char *)it.dev + off, (size_t)(hi - lo) * it.unit,
                                      cudaMemcpyDeviceToHost, s));
            }
        }
        return 0;
    }
};

extern "C" int mmalign_get_pairs(mmalign_ctx *c, int64_t *pair_offsets, int64_t *pair_chunk)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_get_pairs: ctx is NULL");
    int rc = ensure_index(c);
    if (rc) return rc;
    cudaStream_t st = c->s_idx;
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
    if (prm->pipeline_rows < -1) return fail(c, MMALIGN_EINVAL, "pipeline_rows=%d must be >= -1", prm->pipeline_rows);
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

// Pair offsets at the given image rows (a handful of 8-byte reads of the pair index, one synchronisation)
static int pair_offsets_at(mmalign_ctx *c, const std::vector<int64_t> &rows, std::vector<int64_t> *out)
{
    out->assign(rows.size(), 0);
    for (size_t q = 0; q < rows.size(); ++q)
        CU(c, cudaMemcpyAsync(out->data() + q, c->px.offsets + rows[q], sizeof(int64_t), cudaMemcpyDeviceToHost, c->s_idx));
    CU(c, cudaStreamSynchronize(c->s_idx));
    return MMALIGN_OK;
}

// Rows [row0, row0 + n_rows) with their slice of the pair arrays
static int make_row_range(mmalign_ctx *c, int64_t row0, int64_t n_rows, RowRange *out)
{
    RowRange r;
    r.row0 = row0; r.n_rows = n_rows;
    std::vector<int64_t> h;
    int rc = pair_offsets_at(c, {row0, row0 + n_rows}, &h);
    if (rc) return rc;
    r.pair0 = h[0]; r.P_out = h[1] - h[0];
    *out = r;
    return MMALIGN_OK;
}

// K1 of image rows [row0, +n_rows) against chunk rows [col0, +n_cols) of the current corpus into `lists` (row and
// column indices relative to the ranges); the list buffers of the context must hold plan.n_lists lists
static int reserve_lists(mmalign_ctx *c, const FusedPlan &plan)
{
    CU(c, c->list_keys.reserve((size_t)plan.n_lists * plan.cap * sizeof(uint64_t)));
    CU(c, c->list_tau.reserve((size_t)plan.n_lists * sizeof(float)));
    CU(c, c->list_count.reserve((size_t)plan.n_lists * sizeof(int32_t)));
    return MMALIGN_OK;
}

static int launch_fused_range(mmalign_ctx *c, const FusedPlan &plan, int64_t row0, int64_t n_rows, int64_t col0, int64_t n_cols,
                              cudaStream_t st, CandLists *lists, int64_t list_base = 0, int64_t col_base = 0, int set = 0)
{
    Side img = c->img.s, chk = c->chk.s;
    alignas(64) CUtensorMap tmap_a = c->img.tmap, tmap_b = plan.pairs ? c->chk.tmap_half : c->chk.tmap;
    char msg[256];
    if (row0 != 0 || n_rows != img.n) {
        if (encode_tensor_map(&tmap_a, img.emb_bf16 + row0 * img.D, n_rows, img.D, 128, msg, sizeof msg)) return fail(c, MMALIGN_ECUDA, "%s", msg);
        img.n = n_rows;
    }
    if (col0 != 0 || n_cols != chk.n) {
        if (encode_tensor_map(&tmap_b, chk.emb_bf16 + col0 * chk.D, n_cols, chk.D, plan.pairs ? 128 : 256, msg, sizeof msg)) return fail(c, MMALIGN_ECUDA, "%s", msg);
        chk.n = n_cols;
    }
    *lists = CandLists();
    lists->keys = (uint64_t *)(set ? c->list_keys2.p : c->list_keys.p) + list_base * plan.cap;
    lists->tau = (float *)(set ? c->list_tau2.p : c->list_tau.p) + list_base;
    lists->count = (int32_t *)(set ? c->list_count2.p : c->list_count.p) + list_base;
    CU(c, launch_fused(img, chk, plan, &tmap_a, &tmap_b, *lists, nullptr, st, col_base));
    return MMALIGN_OK;
}

static int plan_fused(mmalign_ctx *c, const RunParams &rp, int kprime_req, int n_ranks, int64_t n_rows, int64_t n_cols, FusedPlan *plan,
                      int sms = 0)
{
    if (c->img.s.D % 64 != 0) return fail(c, MMALIGN_EINVAL, "the fused path needs D %% 64 == 0 (D=%d); use MMALIGN_PATH_EXACT", c->img.s.D);
    if (rp.kneed > 256) return fail(c, MMALIGN_ELIMIT, "kneed=%d exceeds 256", rp.kneed);
    const int prc = fused_plan(n_rows, n_cols, c->img.s.D, rp.kneed, kprime_req, sms > 0 ? sms : c->sm_count, n_ranks, plan, c->cta_pairs);
    if (prc) return fail(c, MMALIGN_ELIMIT, "no fused plan for N=%lld M=%lld D=%d K'=%d (code %d)", (long long)n_rows, (long long)n_cols, c->img.s.D, kprime_req, prc);
    plan->epi_sleep_ns = c->epi_sleep_ns;
    plan->compact_one = c->compact_one;
    return MMALIGN_OK;
}

// the fused pass of the sharded entry points: one launch over [row0, +n_rows) x [col0, +n_cols), lists kept in the context
static int run_fused(mmalign_ctx *c, const RunParams &rp, int kprime_req, int n_ranks, int64_t row0, int64_t n_rows,
                     int64_t col0, int64_t n_cols, cudaStream_t st, FusedPlan *plan_out)
{
    FusedPlan plan;
    int rc;
    if ((rc = plan_fused(c, rp, kprime_req, n_ranks, n_rows, n_cols, &plan))) return rc;
    if ((rc = reserve_lists(c, plan))) return rc;
    CU(c, c->fail_rows.reserve((size_t)c->img.s.n * sizeof(int32_t)));
    if ((rc = launch_fused_range(c, plan, row0, n_rows, col0, n_cols, st, &c->lists))) return rc;
    c->lists_col0 = col0;
    c->lists_rows = n_rows; c->lists_cols = n_cols;
    c->lists_valid = true;
    *plan_out = plan;
    return MMALIGN_OK;
}

static int slab_events(mmalign_ctx *c, int n_slabs)
{
    while ((int)c->ev_slab.size() < 5 * n_slabs) {
        cudaEvent_t e = nullptr;
        CU(c, cudaEventCreate(&e));
        c->ev_slab.push_back(e);
    }
    return MMALIGN_OK;
}

// mmalign_run and mmalign_rescore_slab: rank image rows [slab_row0, +slab_rows) (0, 0 = all).  `imported` = the
// candidate lists received from the ranks' fused passes (global chunk indices); without it the fused kernel runs here.
//
// The rows are processed in pipeline slabs that share one set of device output arrays: slab s is prepared (K0),
// contracted (K1), re-scored (K2) and, where its certificate fails, scanned exactly on the run's stream while the
// copy stream still uploads the embedding rows of later slabs (queued by mmalign_set_images) and the download
// stream already returns the results of earlier ones.  Every kernel of every slab is queued before the first
// download is, so a pageable destination (whose copies block the host) stalls nothing on the device.
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
    int64_t row0, N;  // the rows of this run; per-row outputs are [S][N][..], per-pair outputs [S][P]
    if ((rc = resolve_range(c, prm->slab_row0, prm->slab_rows, img.n, "slab rows", &row0, &N))) return rc;
    // ---- outputs
    if ((uo->topk_idx == nullptr) != (uo->topk_score == nullptr)) return fail(c, MMALIGN_EINVAL, "topk_idx and topk_score go together");
    if ((uo->deep_idx == nullptr) != (uo->deep_score == nullptr)) return fail(c, MMALIGN_EINVAL, "deep_idx and deep_score go together");
    const bool fused_path = rp.candidates == MMALIGN_CAND_ALL && prm->path != MMALIGN_PATH_EXACT && M > 0;
    // ---- pipeline slabs
    const void *large_out[] = {uo->topk_idx, uo->pair_rank, uo->pair_sim, uo->pair_score, uo->deep_idx};
    bool host_out = false;
    for (const void *p : large_out) host_out = host_out || (p && !is_device_ptr(p));
    int64_t slab_rows = N > 0 ? N : 1;
    const int64_t wave_rows = (int64_t)c->sm_count * 128;  // one 128-row block per SM
    bool whole_waves_first = false;
    // k2_sms > 0: the exact rescoring of slab s runs on a few SMs left free by the contraction of slab s+1
    int k1_sms = c->sm_count;
    if (fused_path && !imported && c->k2_sms > 0 && prm->pipeline_rows >= 0) {
        k1_sms = (c->sm_count - c->k2_sms) & ~1;
        const int64_t w = (int64_t)k1_sms * 128;
        if (prm->pipeline_rows > 0 ? N > (int64_t)prm->pipeline_rows : N >= 4 * w) {
            if (prm->pipeline_rows == 0) slab_rows = N >= 2 * kSlabWaves * w ? kSlabWaves * w : 2 * w;
        } else k1_sms = c->sm_count;  // a single slab: nothing to overlap
    }
    const bool overlap = k1_sms != c->sm_count;
    if (imported || prm->pipeline_rows < 0) { /* one slab */ }
    else if (prm->pipeline_rows > 0) slab_rows = ((int64_t)prm->pipeline_rows + 127) / 128 * 128;
    else if (overlap) { /* chosen above */ }
    else if ((c->img.n_pieces > 0 || host_out) && N >= 4 * wave_rows) {  // auto: whole waves, at least two slabs
        slab_rows = N >= 2 * kSlabWaves * wave_rows ? kSlabWaves * wave_rows : 2 * wave_rows;
    } else if (fused_path && N > wave_rows && N % wave_rows != 0) {
        // Nothing to overlap, but the row blocks do not fill whole waves of the persistent kernel: one slab of whole
        // waves (one unit = one row block against every column: two lists per row), and the remainder on its own,
        // cut by columns to fill the GPU.  Cutting ALL rows by columns instead would give every row several times the
        // lists and their warm-up (measured at 8 GPUs, 977 row blocks per rank: 108.6 -> ms per rank).
        whole_waves_first = true;
    }
    if (N > slab_rows * kMaxSlabs) slab_rows = ((N + kMaxSlabs - 1) / kMaxSlabs + 127) / 128 * 128;
    std::vector<int64_t> bounds, poff;
    if (N > 0 && whole_waves_first) {
        bounds = {row0, row0 + N / wave_rows * wave_rows, row0 + N};
    } else if (N > 0) {
        for (int64_t r = 0; r < N; r += slab_rows) bounds.push_back(row0 + r);
        bounds.push_back(row0 + N);
    } else {
        bounds.push_back(row0);
    }
    const int n_slabs = (int)bounds.size() - 1;
    if ((rc = pair_offsets_at(c, bounds, &poff))) return rc;
    const int64_t pair0 = poff[0], P = poff[n_slabs] - poff[0];
    Stager sg{c, st};
    Outputs out = {};
    int64_t *d_hits = nullptr;
    double *d_rr = nullptr, *d_sim = nullptr;
    const size_t SNK = (size_t)rp.S * N * rp.kmax, SP = (size_t)rp.S * P;
    sg.map(uo->topk_idx, SNK, &out.topk_idx, Stager::kPerRow, (size_t)rp.kmax * 8, rp.S);
    sg.map(uo->topk_score, SNK, &out.topk_score, Stager::kPerRow, (size_t)rp.kmax * 8, rp.S);
    sg.map(uo->pair_rank, SP, &out.pair_rank, Stager::kPerPair, 4, rp.S);
    sg.map(uo->pair_sim, (size_t)P, &out.pair_sim, Stager::kPerPair, 8, 1);
    sg.map(uo->hits, (size_t)rp.S * rp.n_k, &d_hits);
    sg.map(uo->rr_sum, (size_t)rp.S, &d_rr);
    sg.map(uo->sim_sum, 1, &d_sim);
    sg.map(uo->pair_score, SP, &out.pair_score, Stager::kPerPair, 8, rp.S);
    sg.map(uo->deep_idx, (size_t)rp.S * N * rp.kneed, &out.deep_idx, Stager::kPerRow, (size_t)rp.kneed * 8, rp.S);
    sg.map(uo->deep_score, (size_t)rp.S * N * rp.kneed, &out.deep_score, Stager::kPerRow, (size_t)rp.kneed * 8, rp.S);
    // metric sums need the per-pair arrays even if the caller did not ask for them
    const bool want_sums = uo->hits || uo->rr_sum || uo->sim_sum;
    if ((rc = sg.commit())) return rc;
    DevBuf extra;  // scratch per-pair arrays when only the sums were requested
    if (want_sums && (!out.pair_rank || !out.pair_sim) && P > 0) {
        CU(c, extra.reserve(SP * sizeof(int32_t) + 256 + (size_t)P * sizeof(double)));
        if (!out.pair_rank) out.pair_rank = (int32_t *)extra.p;
        if (!out.pair_sim) out.pair_sim = (double *)((char *)extra.p + ((SP * sizeof(int32_t) + 255) & ~(size_t)255));
    }
    // ---- plans and scratch of every slab, before anything is queued (a growing buffer would stall the pipeline)
    std::vector<FusedPlan> plans((size_t)n_slabs), gplans;
    if (fused_path) {
        CU(c, c->fail_rows.reserve((size_t)(img.n > 0 ? img.n : 1) * sizeof(int32_t)));
        CU(c, c->fail_thr.reserve((size_t)(img.n > 0 ? img.n : 1) * sizeof(unsigned long long)));
        CU(c, c->big_rows.reserve((size_t)(img.n > 0 ? img.n : 1) * sizeof(int32_t)));
        {   // per-row records of the warp-per-row rescoring: the rescoring of one slab at a time uses them
            int64_t most_rows = 1;
            for (int s = 0; s < n_slabs; ++s) most_rows = std::max<int64_t>(most_rows, bounds[s + 1] - bounds[s]);
            CU(c, c->k2_scratch.reserve(k2_scratch_bytes(most_rows)));
        }
        CU(c, c->scan_buf.reserve(scan_scratch_bytes()));
        CU(c, c->scan_cnt.reserve(sizeof(int32_t) * kScanSlots));
        if (!imported) {
            FusedPlan big = {};
            // The first slab of a pipelined run over a chunk table that is still travelling is contracted column
            // group by column group (set_side): its rows get n_groups times the lists, each over one group.
            if (n_slabs > 1 && c->chk.n_groups > 1 && !c->chk_consumed) {
                const SideStore &ck = c->chk;
                gplans.resize((size_t)ck.n_groups);
                bool same = true;
                for (int g = 0; g < ck.n_groups && same; ++g) {
                    if ((rc = plan_fused(c, rp, prm->kprime, ck.n_groups, bounds[1] - bounds[0], ck.group_row[g + 1] - ck.group_row[g], &gplans[g], k1_sms))) return rc;
                    same = gplans[g].cap == gplans[0].cap && gplans[g].kprime_list == gplans[0].kprime_list &&
                           gplans[g].n_splits == gplans[0].n_splits && gplans[g].n_lists == gplans[0].n_lists;
                }
                if (!same) gplans.clear();
            }
            for (int s = 0; s < n_slabs; ++s) {
                if ((rc = plan_fused(c, rp, prm->kprime, 1, bounds[s + 1] - bounds[s], M, &plans[s], k1_sms))) return rc;
                if (s == 0 && !gplans.empty()) { plans[0] = gplans[0]; plans[0].n_lists = gplans[0].n_lists * (int64_t)gplans.size(); }
                if ((size_t)plans[s].n_lists * plans[s].cap >= (size_t)big.n_lists * big.cap) { big.cap = plans[s].cap; big.n_lists = plans[s].n_lists; }
            }
            FusedPlan most = big;
            for (int s = 0; s < n_slabs; ++s) if (plans[s].n_lists > most.n_lists) most.n_lists = plans[s].n_lists;
            if ((rc = reserve_lists(c, big))) return rc;
            CU(c, c->list_tau.reserve((size_t)most.n_lists * sizeof(float)));
            CU(c, c->list_count.reserve((size_t)most.n_lists * sizeof(int32_t)));
            if (overlap) {
                CU(c, c->list_keys2.reserve((size_t)big.n_lists * big.cap * sizeof(uint64_t)));
                CU(c, c->list_tau2.reserve((size_t)most.n_lists * sizeof(float)));
                CU(c, c->list_count2.reserve((size_t)most.n_lists * sizeof(int32_t)));
            }
        }
    }
    if ((rc = slab_events(c, n_slabs))) return rc;
    tr.mark("params + staging");
    // ---- small device state
    int32_t *slab_fail = (int32_t *)((char *)c->small.p + 1024);  // [n_slabs] rows of the slab that missed the certificate
    int32_t *slab_big = (int32_t *)((char *)c->small.p + 4096);   // [n_slabs] rows of the slab left to the block-per-row rescoring
    unsigned long long *cand_counter = (unsigned long long *)((char *)c->small.p + 8);
    int32_t *error_flag = (int32_t *)((char *)c->small.p + 16);
    int32_t *k_list_dev = (int32_t *)((char *)c->small.p + 64);
    if ((rc = queue_pieces(c, c->img, c->img.n_pieces))) return rc;
    if ((rc = wait_side(c, c->img, st))) return rc;
    CU(c, cudaStreamWaitEvent(st, c->chk.ev_small, 0));
    CU(c, cudaStreamWaitEvent(st, c->chk.ev_caller, 0));
    bool chunks_waited = false;  // the whole chunk table prepared (its K0 runs behind the uploads on its own stream)
    auto need_chunks = [&]() -> int {
        if (!chunks_waited && c->chk.ev_ready_set) CU(c, cudaStreamWaitEvent(st, c->chk.ev_ready, 0));
        chunks_waited = true;
        return MMALIGN_OK;
    };
    c->chk_consumed = true;
    CU(c, cudaMemsetAsync(c->small.p, 0, 64, st));
    CU(c, cudaMemsetAsync(slab_fail, 0, sizeof(int32_t) * kMaxSlabs, st));
    CU(c, cudaMemsetAsync(slab_big, 0, sizeof(int32_t) * kMaxSlabs, st));
    CU(c, cudaMemcpyAsync(k_list_dev, rp.k_list, sizeof(int32_t) * kMaxK, cudaMemcpyHostToDevice, st));
    if (out.pair_rank && SP) CU(c, cudaMemsetAsync(out.pair_rank, 0, SP * sizeof(int32_t), st));
    long long launches = 0, fused_launches = 0, kprime_used = 0;
    ScanScratch pre;
    // ---- the slabs.  Without overlap everything runs on the caller's stream.  With it (k2_sms > 0) the contraction
    // runs on a high-priority stream with a grid that leaves k2_sms SMs free, and the exact rescoring of slab s
    // (HBM-bound row gathers) runs on those SMs from a second stream while slab s+1 is contracted; two sets of list
    // buffers alternate.  The rescoring of the last slab has the GPU to itself.
    cudaStream_t sk1 = overlap ? c->s_k1 : st, sk2 = overlap ? c->s_k2 : st;
    if (overlap) {
        if ((rc = order_after(c, sk1, st))) return rc;
        if ((rc = order_after(c, sk2, st))) return rc;
    }
    for (int s = 0; s < n_slabs; ++s) {
        const int64_t r0 = bounds[s], rows = bounds[s + 1] - bounds[s];
        RowRange range;
        range.row0 = r0; range.n_rows = rows;
        range.pair0 = pair0; range.P_out = P;
        range.o_row0 = row0; range.o_rows = N;
        cudaEvent_t *ev = &c->ev_slab[5 * (size_t)s];
        if (overlap) {  // (prepare_images orders sk1 behind the uploads it needs)
            CU(c, cudaStreamWaitEvent(sk1, c->img.ev_small, 0));
            CU(c, cudaStreamWaitEvent(sk1, c->img.ev_caller, 0));
            if (c->img.ev_ready_set) CU(c, cudaStreamWaitEvent(sk1, c->img.ev_ready, 0));
        }
        if ((rc = prepare_images(c, r0, r0 + rows, sk1, &launches))) return rc;
        if (overlap && s >= 2) CU(c, cudaStreamWaitEvent(sk1, c->ev_slab[5 * (size_t)(s - 2) + 3], 0));  // its list buffers are free again
        CU(c, cudaEventRecord(ev[0], sk1));
        if (rp.candidates == MMALIGN_CAND_SAME_PAGE) {
            if ((rc = need_chunks())) return rc;
            CU(c, launch_rescore(img, chk, c->px, rp, nullptr, nullptr, out, nullptr, nullptr, nullptr, cand_counter,
                                 error_flag, nullptr, nullptr, range, st));
            launches += 1;
        } else if (!fused_path) {
            if ((rc = need_chunks())) return rc;
            CU(c, launch_exact_scan(img, chk, c->px, rp, nullptr, nullptr, rows, out, error_flag, range, nullptr, st));
            launches += 1;
        } else {
            CandLists L;
            const int set = overlap ? (s & 1) : 0;
            if (imported) {
                if ((rc = need_chunks())) return rc;
                L = *imported;
                kprime_used = imported->kprime;
            } else if (s == 0 && !gplans.empty()) {
                const SideStore &ck = c->chk;
                for (int g = 0; g < ck.n_groups; ++g) {
                    CU(c, cudaStreamWaitEvent(sk1, ck.ev_group[g], 0));
                    CandLists Lg;
                    if ((rc = launch_fused_range(c, gplans[g], r0, rows, ck.group_row[g], ck.group_row[g + 1] - ck.group_row[g], sk1, &Lg,
                                                 (int64_t)g * gplans[0].n_lists, ck.group_row[g], set))) return rc;
                    if (g == 0) L = Lg;
                }
                L.n_splits = gplans[0].n_splits * ck.n_groups;  // one row's lists: n_groups x (splits x 2 halves)
                fused_launches += ck.n_groups;
                launches += ck.n_groups;
                kprime_used = gplans[0].kprime;
                if ((rc = need_chunks())) return rc;
            } else {
                if ((rc = need_chunks())) return rc;
                if (overlap) CU(c, cudaStreamWaitEvent(sk1, c->chk.ev_ready, 0));
                if ((rc = launch_fused_range(c, plans[s], r0, rows, 0, M, sk1, &L, 0, 0, set))) return rc;
                fused_launches += 1;
                launches += 1;
                kprime_used = plans[s].kprime;
            }
            CU(c, cudaEventRecord(ev[1], sk1));
            if (overlap) {
                CU(c, cudaStreamWaitEvent(sk2, ev[1], 0));
                if (c->chk.ev_ready_set) CU(c, cudaStreamWaitEvent(sk2, c->chk.ev_ready, 0));
            }
            if (s == 0 && c->resc_wait) {  // the exact rescoring reads the fp32 master rows: mmalign_rescore_after
                CU(c, cudaStreamWaitEvent(sk2, c->resc_wait, 0));
                c->resc_wait = nullptr;
            }
            CU(c, cudaEventRecord(ev[4], sk2));
            int32_t *fail_rows = (int32_t *)c->fail_rows.p + r0;
            long long k2_launched = 0;
            pre.thr = (unsigned long long *)c->fail_thr.p + r0; pre.buf = c->scan_buf.p; pre.cnt = (int32_t *)c->scan_cnt.p;
            // beside the next slab's contraction the rescoring gets the SMs that were left free: 8 CTAs on each
            const int64_t k2_grid = overlap && s + 1 < n_slabs ? (int64_t)(c->sm_count - k1_sms) * 8 : 0;
            CU(c, launch_rescore(img, chk, c->px, rp, &L, c->chk.err_max, out, fail_rows, slab_fail + s, pre.thr,
                                 cand_counter, error_flag, nullptr, nullptr, range, sk2, k2_grid,
                                 (int32_t *)c->big_rows.p + r0, slab_big + s, c->k2_scratch.p, &k2_launched));
            CU(c, cudaEventRecord(ev[2], sk2));
            CU(c, launch_exact_scan(img, chk, c->px, rp, fail_rows, slab_fail + s, 0, out, error_flag, range, &pre, sk2));
            launches += k2_launched + 2;  // + the two stages of the exact scan
        }
        CU(c, cudaEventRecord(ev[3], fused_path ? sk2 : st));
    }
    if (overlap && n_slabs > 0) {  // the caller's stream continues behind both
        if ((rc = order_after(c, st, sk1))) return rc;
        if ((rc = order_after(c, st, sk2))) return rc;
    }
    c->lists_valid = false;  // (the context's lists cover the last slab only)
    tr.mark("queued scoring");
    // ---- metric sums
    if (want_sums) {
        CU(c, c->metrics_scratch.reserve(metrics_scratch_bytes(rp.S, rp.n_k)));
        CU(c, launch_reduce_metrics(out.pair_rank, out.pair_sim, rp.S, P, k_list_dev, rp.n_k, rp.mrr_cutoff, d_hits,
                                    d_rr, d_sim, c->metrics_scratch.p, st));
        launches += 2;
    }
    // ---- results of the slabs travel while later slabs compute
    const bool stream_out = sg.any_large();
    if (stream_out) {
        for (int s = 0; s < n_slabs; ++s) {
            CU(c, cudaStreamWaitEvent(c->s_out, c->ev_slab[5 * (size_t)s + 3], 0));
            if ((rc = sg.copy_slab(c->s_out, bounds[s] - row0, bounds[s + 1] - row0, N, poff[s] - pair0, poff[s + 1] - pair0, P))) return rc;
        }
    }
    // ---- status, stats
    struct { int32_t fail; int32_t pad; unsigned long long cand; int32_t err; int32_t pad2; unsigned long long viol; } h = {};
    std::vector<int32_t> h_fail((size_t)(n_slabs > 0 ? n_slabs : 1), 0);
    CU(c, cudaMemcpyAsync(&h, c->small.p, 32, cudaMemcpyDeviceToHost, st));
    if (n_slabs > 0) CU(c, cudaMemcpyAsync(h_fail.data(), slab_fail, sizeof(int32_t) * n_slabs, cudaMemcpyDeviceToHost, st));
    if ((rc = sg.copy_back(st, true))) { extra.release(); return rc; }
    CU(c, cudaStreamSynchronize(st));
    if (stream_out) CU(c, cudaStreamSynchronize(c->s_out));
    tr.mark("kernels + copies done");
    extra.release();
    int64_t n_fail = 0;
    for (int s = 0; s < n_slabs; ++s) n_fail += h_fail[s];
    if (h.err) return fail(c, MMALIGN_ELIMIT, "an image has more than 512 same-page chunks (capacity limit of this build)");
    if (prm->path == MMALIGN_PATH_FUSED && n_fail > 0)
        return fail(c, MMALIGN_ELIMIT, "%lld rows were not certified by the fused path (MMALIGN_PATH_FUSED forbids the exact rescan)", (long long)n_fail);
    if (uo->num_pairs) CU(c, cudaMemcpy(uo->num_pairs, &P, sizeof(int64_t), cudaMemcpyDefault));
    if (uo->stats) {
        double t_fused = 0.0, t_resc = 0.0, t_scan = 0.0;
        if (!fused_path) {  // same-page candidates: the rescoring kernel alone; exact path: the scan alone
            for (int s = 0; s < n_slabs; ++s) {
                float a = 0.f;
                cudaEventElapsedTime(&a, c->ev_slab[5 * (size_t)s], c->ev_slab[5 * (size_t)s + 3]);
                (rp.candidates == MMALIGN_CAND_SAME_PAGE ? t_resc : t_scan) += a;
            }
        } else {
            for (int s = 0; s < n_slabs; ++s) {
                float a = 0.f, b = 0.f, d = 0.f;
                const cudaEvent_t *ev = &c->ev_slab[5 * (size_t)s];
                if (fused_launches) cudaEventElapsedTime(&a, ev[0], ev[1]);
                cudaEventElapsedTime(&b, ev[4], ev[2]);
                cudaEventElapsedTime(&d, ev[2], ev[3]);
                t_fused += a; t_resc += b; t_scan += d;
            }
        }
        const int64_t stats[16] = {n_fail, (int64_t)h.cand, imported ? 1 : fused_launches, launches + (imported ? 2 : 0), kprime_used,
                                   imported ? c->last_fused_us : (int64_t)(t_fused * 1000.0), (int64_t)(t_resc * 1000.0),
                                   (int64_t)(t_scan * 1000.0), (int64_t)h.viol, n_slabs, overlap ? c->sm_count - k1_sms : 0};
        CU(c, cudaMemcpy(uo->stats, stats, sizeof stats, cudaMemcpyDefault));
    }
    tr.mark("status");
    return MMALIGN_OK;
}

// Checked build only (make check): out[0] = 1, then {violations, first failing line} of fused_tc.cu, rescore.cu,
// prep.cu, ingest.cu.  The release build reports out[0] = 0 and zeros.
extern "C" int mmalign_check_report(uint32_t *out)
{
    if (!out) return MMALIGN_EINVAL;
#ifdef MMALIGN_CHECKED
    out[0] = 1u;
#else
    out[0] = 0u;
#endif
    out[1] = out[2] = out[3] = out[4] = out[5] = out[6] = out[7] = out[8] = 0u;
    if (cudaDeviceSynchronize() != cudaSuccess) return MMALIGN_EDEVICE;
    unsigned int w[2];
    int (*readers[4])(unsigned int *) = {check_read_fused, check_read_rescore, check_read_prep, check_read_ingest};
    for (int q = 0; q < 4; ++q) {
        if (readers[q](w) != 0) return MMALIGN_EDEVICE;
        out[1 + 2 * q] = w[0]; out[2 + 2 * q] = w[1];
    }
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
    if ((rc = make_row_range(c, r0, n, &range))) return rc;
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
    if ((rc = wait_tables(c, st))) return rc;
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
    // the lists' columns are relative to the range of the fused pass; this pass reads them as table rows
    if (M > 0 && (c->lists_col0 != 0 || c->lists_cols != M || c->lists_rows != N))
        return fail(c, MMALIGN_ESTATE, "mmalign_rescore_pass needs lists of a fused pass over the whole tables (the last one covered "
                    "%lld rows x columns [%lld, +%lld)); lists of a column shard go through mmalign_export_lists",
                    (long long)c->lists_rows, (long long)c->lists_col0, (long long)c->lists_cols);
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    if ((rc = wait_tables(c, st))) return rc;
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
        const int64_t stats[16] = {0, (int64_t)h.cand, M > 0, 4, c->lists.kprime, (int64_t)(t_fused * 1000.f), (int64_t)(t_resc * 1000.f), 0};
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
    if ((rc = wait_tables(c, st))) return rc;
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
    CU(c, cudaStreamSynchronize(c->s_prep));
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
    if ((rc = wait_tables(c, st))) return rc;
    FusedPlan plan;
    if (fused_plan(img.n, chk.n, img.D, 10, 0, c->sm_count, 1, &plan, c->cta_pairs)) return fail(c, MMALIGN_ELIMIT, "no fused plan");
    Stager sg{c, st};
    float *d = nullptr;
    sg.map(out, (size_t)img.n * chk.n, &d);
    if ((rc = sg.commit())) return rc;
    CandLists L;
    CU(c, launch_fused(img, chk, plan, &c->img.tmap, plan.pairs ? &c->chk.tmap_half : &c->chk.tmap, L, d, st));
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}

extern "C" int mmalign_debug_operands(mmalign_ctx *c, void *img_bf16, void *chk_bf16, void *stream)
{
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "mmalign_debug_operands: ctx is NULL");
    if (!c->img.ready || !c->chk.ready) return fail(c, MMALIGN_ESTATE, "set_images and set_chunks must be called first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = wait_tables(c, st))) return rc;
    const Side &img = c->img.s, &chk = c->chk.s;
    if (img_bf16 && img.n > 0) CU(c, cudaMemcpyAsync(img_bf16, img.emb_bf16, (size_t)img.n * img.D * 2, cudaMemcpyDefault, st));
    if (chk_bf16 && chk.n > 0) CU(c, cudaMemcpyAsync(chk_bf16, chk.emb_bf16, (size_t)chk.n * chk.D * 2, cudaMemcpyDefault, st));
    CU(c, cudaStreamSynchronize(st));
    return MMALIGN_OK;
}

// ---- half-precision encoder rows ---------------------------------------------------
static int set_side_half(mmalign_ctx *c, bool chunks, const void *emb, int32_t dtype, const uint64_t *key, const double *bbox,
                         const uint64_t *terms, int64_t n, int32_t D, int32_t term_words, int64_t n_terms, int64_t col_offset)
{
    const char *what = chunks ? "mmalign_set_chunks_half" : "mmalign_set_images_half";
    if (!c) return fail(nullptr, MMALIGN_EINVAL, "%s: ctx is NULL", what);
    if (dtype == MMALIGN_F32)
        return chunks ? mmalign_set_chunks(c, (const float *)emb, key, bbox, terms, n, D, term_words, n_terms, col_offset)
                      : mmalign_set_images(c, (const float *)emb, key, bbox, terms, n, D, term_words);
    if (dtype != MMALIGN_F16 && dtype != MMALIGN_BF16) return fail(c, MMALIGN_EINVAL, "%s: emb_dtype=%d unknown", what, dtype);
    if (n < 0 || n > 0x7FFFFFF0ll || D <= 0 || D % 4 != 0 || D > 4096) return fail(c, MMALIGN_EINVAL, "%s: bad n / D", what);
    if (n > 0 && !emb) return fail(c, MMALIGN_EINVAL, "%s: emb is required", what);
    CU(c, cudaSetDevice(c->device));
    SideStore &ss = chunks ? c->chk : c->img;
    // the widened rows live in the side's upload buffer; earlier work may still read it
    int rc;
    if ((rc = mmalign_sync(c))) return rc;
    CU(c, cudaStreamSynchronize(0));
    const size_t count = (size_t)n * D;
    CU(c, ss.up[0].reserve((count ? count : 1) * sizeof(float)));
    const void *src = emb;
    if (count && !is_device_ptr(emb)) {
        CU(c, c->half_up.reserve(count * 2));
        CU(c, cudaMemcpyAsync(c->half_up.p, emb, count * 2, cudaMemcpyHostToDevice, 0));
        src = c->half_up.p;
    }
    CU(c, launch_widen_rows(src, dtype, (int64_t)count, (float *)ss.up[0].p, 0));
    const float *master = n > 0 ? (const float *)ss.up[0].p : nullptr;
    if (n == 0) {
        static const float dummy[4] = {0.f, 0.f, 0.f, 0.f};
        master = dummy;  // (never read: n == 0)
    }
    return chunks ? mmalign_set_chunks(c, master, key, bbox, terms, n, D, term_words, n_terms, col_offset)
                  : mmalign_set_images(c, master, key, bbox, terms, n, D, term_words);
}

extern "C" int mmalign_set_images_half(mmalign_ctx *c, const void *emb, int32_t dtype, const uint64_t *key, const double *bbox,
                                       const uint64_t *terms, int64_t n, int32_t D, int32_t term_words)
{
    return set_side_half(c, false, emb, dtype, key, bbox, terms, n, D, term_words, 0, 0);
}

extern "C" int mmalign_set_chunks_half(mmalign_ctx *c, const void *emb, int32_t dtype, const uint64_t *key, const double *bbox,
                                       const uint64_t *terms, int64_t m, int32_t D, int32_t term_words, int64_t n_terms,
                                       int64_t col_offset)
{
    return set_side_half(c, true, emb, dtype, key, bbox, terms, m, D, term_words, n_terms, col_offset);
}

// ---- pgvector interop: binary COPY streams -----------------------------------------
extern "C" int64_t mmalign_copy_scan(const uint8_t *data, int64_t n_bytes, int32_t n_cols, int64_t *field_off, int32_t *field_len,
                                     int64_t cap)
{
    if (!data || n_bytes < 0 || n_cols < 1 || n_cols > 1600 || (field_off && !field_len)) return -1;
    return copy_scan(data, n_bytes, n_cols, field_off, field_len, cap);
}

extern "C" int mmalign_copy_decode(mmalign_ctx *c, const uint8_t *data, int64_t n_bytes, const int64_t *field_off,
                                   const int32_t *field_len, int64_t n, int32_t n_cols, int32_t vec_col, int32_t bbox_col,
                                   int32_t page_col, int32_t D, float *emb, double *bbox, int32_t *page, uint8_t *page_null,
                                   void *stream)
{
    if (!c || !data || !field_off || !field_len) return fail(c, MMALIGN_EINVAL, "mmalign_copy_decode: NULL argument");
    if (n < 0 || n_cols < 1 || vec_col >= n_cols || bbox_col >= n_cols || page_col >= n_cols || (emb && (vec_col < 0 || D <= 0)))
        return fail(c, MMALIGN_EINVAL, "mmalign_copy_decode: bad column indices / D");
    if (n == 0) return MMALIGN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    const uint8_t *d_data = data;
    if (!is_device_ptr(data)) {
        CU(c, c->text_bytes.reserve((size_t)n_bytes + 8));  // (+8: the unaligned 32-bit reads look one word past a field's end)
        CU(c, cudaMemcpyAsync(c->text_bytes.p, data, (size_t)n_bytes, cudaMemcpyHostToDevice, st));
        d_data = (const uint8_t *)c->text_bytes.p;
    }
    const int64_t *d_off = field_off;
    const int32_t *d_len = field_len;
    if (!is_device_ptr(field_off)) {
        CU(c, c->copy_off.reserve((size_t)n * n_cols * sizeof(int64_t)));
        CU(c, cudaMemcpyAsync(c->copy_off.p, field_off, (size_t)n * n_cols * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        d_off = (const int64_t *)c->copy_off.p;
    }
    if (!is_device_ptr(field_len)) {
        CU(c, c->copy_len.reserve((size_t)n * n_cols * sizeof(int32_t)));
        CU(c, cudaMemcpyAsync(c->copy_len.p, field_len, (size_t)n * n_cols * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        d_len = (const int32_t *)c->copy_len.p;
    }
    Stager sg{c, st};
    float *d_emb = nullptr;
    double *d_bbox = nullptr;
    int32_t *d_page = nullptr;
    uint8_t *d_null = nullptr;
    sg.map(emb, (size_t)n * (D > 0 ? D : 0), &d_emb);
    sg.map(bbox, (size_t)n * 4, &d_bbox);
    sg.map(page, (size_t)n, &d_page);
    sg.map(page_null, (size_t)n, &d_null);
    int rc;
    if ((rc = sg.commit())) return rc;
    int32_t *err = (int32_t *)((char *)c->small.p + 192);
    CU(c, cudaMemsetAsync(err, 0, sizeof(int32_t), st));
    CU(c, launch_copy_decode(d_data, d_off, d_len, n, n_cols, vec_col, bbox_col, page_col, D, d_emb, d_bbox, d_page, d_null, err, st));
    int32_t h_err = 0;
    CU(c, cudaMemcpyAsync(&h_err, err, sizeof h_err, cudaMemcpyDeviceToHost, st));
    if ((rc = sg.copy_back())) return rc;
    CU(c, cudaStreamSynchronize(st));
    if (h_err) return fail(c, MMALIGN_EINVAL, "mmalign_copy_decode: a vector field is NULL or not of dimension %d", D);
    return MMALIGN_OK;
}
