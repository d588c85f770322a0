// fused_tc.cu -- K1: the image x chunk similarity contraction on the 5th-gen
// tensor cores, with a streaming per-row top-K' in the epilogue, so the N x M
// score matrix never reaches HBM.  (The reference ranks with one SQL statement
// per image: src/evaluate_alignments.py:126-136.)
//
// One persistent CTA per SM, 12 warps, warp-specialised:
//   warp 0      TMA producer: A row block (128 images x D, bf16, resident in shared
//               memory for the whole sweep when D <= 512) and B k-slices (256 chunks
//               x 64) through a ring of mbarrier-guarded stages, SWIZZLE_128B.
//   warp 1      MMA issuer: tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16,
//               fp32 accumulators in TMEM, double-buffered (2 x 256 columns).
//   warp 2      TMEM allocator.
//   warps 4-11  epilogue: tcgen05.ld 32 columns at a time (one thread = one image
//               row of one accumulator half), compare against the row's running
//               threshold tau, append survivors to the row's candidate list in
//               global memory; a full list is compacted by the whole warp to its
//               best K' entries (ballot / popc / redux selection), which raises tau.
// A list is complete above its final tau: every column the thread saw with score >
// tau is in it.  rescore.cu re-scores the lists exactly and certifies each row.
#include "common.cuh"
#include <cuda.h>
#include <math_constants.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>

namespace mma {
MMA_CHECK_DECL

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kFusedThreads = 384;
constexpr uint32_t kABlockBytes = BM * BK * 2;   // 16 KiB: one 64-wide k block of the A tile
constexpr uint32_t kBStageBytes = BN * BK * 2;   // 32 KiB
constexpr int kMaxStages = 8;
constexpr int kMaxPairStages = 12;      // the pair kernel's ring (16 / 32 KiB stages)
constexpr size_t kSmemLimit = 232448;            // 227 KiB opt-in maximum per CTA

// ---------------------------------------------------------------------------
// Profiling build (-DMMALIGN_PROFILE_EPI, tools/k1_epilogue_profile.py): where the epilogue warps spend their cycles.
// ---------------------------------------------------------------------------
#ifdef MMALIGN_PROFILE_EPI
__device__ unsigned long long g_epi_prof[16];
#define EPI_T(var) const long long var = clock64()
#define EPI_ADD(slot, expr) prof[slot] += (expr)
extern "C" int mmalign_profile_counters(unsigned long long *out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, g_epi_prof, sizeof(g_epi_prof));
    if (e == cudaSuccess && reset) {
        unsigned long long z[16] = {};
        e = cudaMemcpyToSymbol(g_epi_prof, z, sizeof z);
    }
    return (int)e;
}
#else
#define EPI_T(var)
#define EPI_ADD(slot, expr)
#endif

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out.  The wait
// loops (MMA issuer on `full`, the eight epilogue warps on `tmem_full`) are 37 % of the executed warp-instructions
// on the ncu source page; the hint was measured to leave the kernel time unchanged (204 ms either way at
// 250k x 1M x 512), so the waiting is not what limits the kernel -- kept because it costs nothing.
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0;
}
// A wait that lasts seconds is a protocol deadlock, not a long wait (every legitimate wait of this kernel is a
// few microseconds): trap, so that the launch fails with an error instead of hanging the GPU.
__device__ __noinline__ void mbar_watchdog(uint64_t &t0)
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t0 == 0) t0 = t;
    else if (t - t0 > 8000000000ull) __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity))
        if ((++polls & 0xFFFu) == 0) mbar_watchdog(t0);
}
// Polling with a pause: the eight epilogue warps wait for an accumulator most of the time (they are ahead of the
// tensor pipe); a spinning wait is executed instructions, i.e. power, on a kernel that runs against the power cap.
__device__ __forceinline__ void mbar_wait_paused(uint32_t bar, uint32_t parity, uint32_t ns)
{
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (ns) __nanosleep(ns);
        if ((++polls & 0xFFFu) == 0) mbar_watchdog(t0);
    }
}
// For the single-thread roles: their spin loops would otherwise steal issue slots from the
// epilogue warps that share their scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(40);
        if ((++polls & 0xFFFu) == 0) mbar_watchdog(t0);
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// The registers of the pending load are in/out operands, so no use of them can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t *v)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

// K-major, SWIZZLE_128B shared-memory operand descriptor (sm_100 "version 1"):
// rows are 128 bytes, 8-row swizzle atoms are 1024 bytes apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);

struct FusedArgs {
    int64_t N, M;
    int num_kb;            // D / 64
    int64_t n_row_blocks, n_tiles;
    int n_splits, tiles_per_split;
    int stages;
    uint64_t *keys;        // [n_lists][cap] entries: lo = column, hi = fp32 score bits
    float *tau;
    int32_t *count;
    int kprime;
    float *dump;           // debug: write raw scores [N][M] instead of lists
    float tau_init;        // -inf; tuning builds (-DMMALIGN_TUNING) can start the lists at a threshold
    int skip_final;        // tuning builds: no end-of-unit compaction
    uint32_t col_base;     // added to the column index of every entry (a launch over a column group of the table)
    uint32_t epi_sleep_ns; // the epilogue warps poll their accumulator barrier this many ns apart (0 = spin)
    int compact_one;       // routine compaction: 0 = every flagged list at once (default), 1 = one list per gap between tiles, 2 = split over two gaps
    int diag;              // tuning builds (-DMMALIGN_TUNING, MMALIGN_K1_DIAG): 1 = the epilogue hands every accumulator back
                           // unread, 2 = no operand loads after the ring's first fill; always 0 in the release build
};
#ifdef MMALIGN_TUNING
#define K1_DIAG(bit) ((P.diag & (bit)) != 0)
#else
#define K1_DIAG(bit) false
#endif

// ---------------------------------------------------------------------------
// Candidate lists.  An entry is 8 bytes: lo = chunk column, hi = fp32 score bits.
// Thread (row, half) owns one list of CAP slots; `n` entries are live.
// ---------------------------------------------------------------------------
constexpr int kSlack = kListSlack;  // a compaction may keep up to K' + kSlack entries (saves bisection steps)
constexpr int kCompactMargin = 72;  // routine compaction when fewer than this many slots are free (one tile adds <= 128 - ...)

// Warp-cooperative compaction of the lists of the lanes that ask for it: keeps the best ~K'
// entries and raises the lane's threshold tau to the smallest kept score.
// `one`: at most one list per call (the routine compaction between two tiles: a warp that compacts several lists in
// one go -- 6.5k cycles measured -- comes late to its next accumulator, and the accumulator's hand-back, i.e. the
// tensor pipe, waits for the slowest of the epilogue warps; the lists left over keep their flag and take the next
// gaps -- 72 free slots last for hundreds of tiles).
template <int KPL>
__device__ __noinline__ void compact_lists(uint2 *my_list, int &n, float &tau, int kprime, bool need, bool one = false)
{
    constexpr int CAP = 32 * KPL;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned pending = __ballot_sync(0xFFFFFFFFu, need);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const int cnt = __shfl_sync(0xFFFFFFFFu, n, src);
        MMA_CHECK(cnt >= 0 && cnt <= CAP);  // a list never holds more than its capacity
        uint2 *L = reinterpret_cast<uint2 *>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(my_list), src));
        __syncwarp();
#ifdef MMALIGN_PROFILE_EPI
        const long long q0_ = clock64();
#endif
        uint32_t h[KPL], c[KPL];  // ordered score key, column
        uint32_t hmin = 0xFFFFFFFFu, hmax = 0u;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int idx = q * 32 + lane;
            h[q] = 0u; c[q] = 0u;
            if (idx < cnt) {
                const uint2 e = __ldcg(L + idx);
                c[q] = e.x;
                h[q] = f32_ordered(__uint_as_float(e.y));  // >= 1 for every real score
                hmin = min(hmin, h[q]);
                hmax = max(hmax, h[q]);
            }
        }
        uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, hmin);
        uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, hmax);
#ifdef MMALIGN_PROFILE_EPI
        const long long q1_ = clock64();
#endif
        int c_lo = cnt;  // #{key >= lo}
        // largest t with #{key >= t} >= kprime, stopping early once the count is within the slack
        while (lo < hi && c_lo > kprime + kSlack) {
            const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
            int cc = 0;
#pragma unroll
            for (int q = 0; q < KPL; ++q) cc += (h[q] >= mid);
            cc = __reduce_add_sync(0xFFFFFFFFu, cc);
            if (cc >= kprime) { lo = mid; c_lo = cc; } else hi = mid - 1u;
        }
        const uint32_t t = lo;
#ifdef MMALIGN_PROFILE_EPI
        const long long q2_ = clock64();
#endif
        const bool drop_ties = c_lo > CAP - 64;  // a wall of equal scores: keep only what is above it
        int base = 0;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const bool keep = drop_ties ? h[q] > t : h[q] >= t;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
            if (keep) L[base + __popc(m & lt_mask)] = make_uint2(c[q], __float_as_uint(f32_unordered(h[q])));
            base += __popc(m);
        }
        __syncwarp();
        MMA_CHECK(base <= cnt && base <= CAP);
        if (lane == src) { n = base; tau = fmaxf(tau, f32_unordered(t)); }
#ifdef MMALIGN_PROFILE_EPI
        if (lane == 0) {
            const long long q3_ = clock64();
            atomicAdd(&g_epi_prof[12], (unsigned long long)(q1_ - q0_)); atomicAdd(&g_epi_prof[13], (unsigned long long)(q2_ - q1_));
            atomicAdd(&g_epi_prof[14], (unsigned long long)(q3_ - q2_)); atomicAdd(&g_epi_prof[15], 1ull);
        }
#endif
        if (one) break;
    }
}

// Split compaction (the routine case, lists of up to 256 entries).  Compacting a list takes ~4k cycles (load 1.7k,
// bisection 1.1k, write-back 1.2k: tools/k1_epilogue_profile.py) against ~2.2k cycles of slack between two tiles, so
// a warp that compacts arrives late at its next accumulator, and the accumulator's hand-back -- i.e. the tensor pipe
// -- waits for the slowest of the epilogue warps; with eight (sixteen on CTA pairs) warps and 0.09 compactions per warp
// and tile that is most tiles.  Split: the loads of ONE flagged list are issued in the gap after a tile and consumed in
// the gap after the next one, so their latency runs under that tile's filtering.  What the owner appended in between
// (behind the snapshot) is moved down behind the compacted entries.
template <int KPL>
__device__ __forceinline__ void compact_issue(const uint2 *my_list, int n, bool need, int &src, int &cnt, uint2 (&e)[KPL])
{
    const int lane = threadIdx.x & 31;
    const unsigned pending = __ballot_sync(0xFFFFFFFFu, need);
    src = -1;
    if (!pending) return;
    src = __ffs(pending) - 1;
    cnt = __shfl_sync(0xFFFFFFFFu, n, src);
    const uint2 *L = reinterpret_cast<const uint2 *>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(my_list), src));
    __syncwarp();
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int idx = q * 32 + lane;
        e[q] = make_uint2(0u, 0u);
        if (idx < cnt) e[q] = __ldcg(L + idx);
    }
}

template <int KPL>
__device__ __forceinline__ void compact_complete(uint2 *my_list, int &n, float &tau, int kprime, int src, int cnt,
                                                 const uint2 (&e)[KPL])
{
    constexpr int CAP = 32 * KPL;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2 *L = reinterpret_cast<uint2 *>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(my_list), src));
    uint32_t h[KPL];
    uint32_t hmin = 0xFFFFFFFFu, hmax = 0u;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        h[q] = 0u;
        if (q * 32 + lane < cnt) {
            h[q] = f32_ordered(__uint_as_float(e[q].y));
            hmin = min(hmin, h[q]);
            hmax = max(hmax, h[q]);
        }
    }
    uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, hmin);
    uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, hmax);
    int c_lo = cnt;
    while (lo < hi && c_lo > kprime + kSlack) {  // (as compact_lists)
        const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
        int cc = 0;
#pragma unroll
        for (int q = 0; q < KPL; ++q) cc += (h[q] >= mid);
        cc = __reduce_add_sync(0xFFFFFFFFu, cc);
        if (cc >= kprime) { lo = mid; c_lo = cc; } else hi = mid - 1u;
    }
    const uint32_t t = lo;
    const bool drop_ties = c_lo > CAP - 64;
    int base = 0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const bool keep = drop_ties ? h[q] > t : h[q] >= t;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) L[base + __popc(m & lt_mask)] = e[q];
        base += __popc(m);
    }
    __syncwarp();
    MMA_CHECK(base <= cnt && base <= CAP);
    if (lane == src) {
        const float t_new = f32_unordered(t);
        int w = base;
        for (int q = cnt; q < n; ++q) {  // appended since the snapshot: a handful
            const uint2 x = my_list[q];
            if (__uint_as_float(x.y) > t_new) my_list[w++] = x;
        }
        n = w;
        tau = fmaxf(tau, t_new);
    }
    __syncwarp();
}

__device__ __forceinline__ float max8(const uint32_t *v)
{
    const float a = fmaxf(fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), __uint_as_float(v[2]));
    const float b = fmaxf(fmaxf(__uint_as_float(v[3]), __uint_as_float(v[4])), __uint_as_float(v[5]));
    return fmaxf(fmaxf(a, b), fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])));
}

// One 32-column chunk of one accumulator row: append every score above tau to the row's list.
// Appends the columns of one 8-column group whose score exceeds tau to the thread's list; returns the new write pointer.
__device__ __noinline__ uint2 *append_group(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                            uint32_t a6, uint32_t a7, uint32_t c0, float tau, uint2 *wp)
{
    if (__uint_as_float(a0) > tau) { *wp = make_uint2(c0, a0); ++wp; }
    if (__uint_as_float(a1) > tau) { *wp = make_uint2(c0 + 1, a1); ++wp; }
    if (__uint_as_float(a2) > tau) { *wp = make_uint2(c0 + 2, a2); ++wp; }
    if (__uint_as_float(a3) > tau) { *wp = make_uint2(c0 + 3, a3); ++wp; }
    if (__uint_as_float(a4) > tau) { *wp = make_uint2(c0 + 4, a4); ++wp; }
    if (__uint_as_float(a5) > tau) { *wp = make_uint2(c0 + 5, a5); ++wp; }
    if (__uint_as_float(a6) > tau) { *wp = make_uint2(c0 + 6, a6); ++wp; }
    if (__uint_as_float(a7) > tau) { *wp = make_uint2(c0 + 7, a7); ++wp; }
    return wp;
}

#ifdef MMALIGN_PROFILE_EPI
#define PC_PROF , long long *prof
#define PC_PROF_ARG , prof
#else
#define PC_PROF
#define PC_PROF_ARG
#endif
#ifdef MMALIGN_TUNING
__device__ int g_hit_mode;  // tuning builds (MMALIGN_K1_HIT): 1 = a hit group appends its maximum only (one store, no call: lists NOT valid), 2 = hits are found but nothing is stored
#endif
template <int KPL>
__device__ __forceinline__ void process_chunk(uint32_t *v, int64_t col0, int64_t M, uint2 *list, int &n, float &tau,
                                              int kprime, uint32_t col_base, int &pend PC_PROF)
{
    constexpr int CAP = 32 * KPL;
    EPI_T(p0_);
    if (col0 + 32 > M) {  // ragged last tile: TMA zero-filled these columns
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (col0 + k >= M) v[k] = 0xFF800000u;  // -inf
    }
    if (__any_sync(0xFFFFFFFFu, n > CAP - 32)) {  // (the first tiles of a sweep; a snapshot taken for a split compaction is stale now)
        compact_lists<KPL>(list, n, tau, kprime, n > CAP - 32);
        pend = -1;
    }
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = max8(v + 8 * q);
    const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
    EPI_T(p1_);
    EPI_ADD(8, p1_ - p0_);  // maxima + vote
    if (__any_sync(0xFFFFFFFFu, m > tau)) {
        EPI_ADD(9, 1);      // chunks with a hit
        // Appends are rare once tau has risen, and the sixteen call sites of a tile (4 chunks x 4 groups) would each
        // carry their own copy of the append code: kept out of line, ONE copy serves them all and stays in the
        // instruction cache (the inlined copies were measured at ~430 cycles per 8-column group with a hit, most
        // of it instruction fetch).
        uint2 *wp = list + n;
        const uint32_t c0 = (uint32_t)col0 + col_base;
#ifdef MMALIGN_TUNING
        if (g_hit_mode) {  // (diagnosis: what a hit path of a few instructions would buy)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (g[q] > tau) { if (g_hit_mode == 1) *wp = make_uint2(c0 + 8 * q, __float_as_uint(g[q])); ++wp; }
            n = (int)(wp - list);
            return;
        }
#endif
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (__any_sync(0xFFFFFFFFu, g[q] > tau)) {
                EPI_ADD(10, 1);  // 8-column groups with a hit
                wp = append_group(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5],
                                  v[8 * q + 6], v[8 * q + 7], c0 + 8 * q, tau, wp);
            }
        }
        n = (int)(wp - list);
        MMA_CHECK(n >= 0 && n <= CAP);  // 32 appends at most, behind a compaction that left 32 slots free
        EPI_T(p2_);
        EPI_ADD(11, p2_ - p1_);  // the hit path
    }
}

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// B multicast (the MC variant of the kernel below): this CTA's half of a B k-slice lands at the same offset of BOTH
// CTAs of the cluster and completes bytes on the full barrier of each; a commit frees the stage in both.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void *tmap, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar)  // arrives on `bar` in both CTAs when this CTA's MMAs retire
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// ---------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------
// MC = true: launched as clusters of two CTAs (the two SMs of a TPC) that contract two row blocks against the SAME
// column tiles in lock step and share the B ring: each CTA loads half of every B k-slice and multicasts it into
// both, so every B byte leaves L2 once per pair instead of once per SM -- the operand traffic that paces the plain
// kernel (tools/k1_diag.py) halves -- while MMAs, accumulators and epilogues stay private to each CTA (unlike the
// cta_group::2 kernel further down, whose pair meets on one accumulator barrier every tile).  The CTAs are coupled
// through the ring only: a stage is refilled when BOTH have consumed it (empty barriers count two commits).
template <int KPL, bool A_RES, bool MC>
__device__ __forceinline__ void fused_body(const CUtensorMap &tmap_a, const CUtensorMap &tmap_b, const FusedArgs &P)
{
    constexpr int CAP = 32 * KPL;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1 KiB alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = P.num_kb;
    const uint32_t a_bytes = A_RES ? (uint32_t)num_kb * kABlockBytes : 0u;
    const uint32_t stage_bytes = kBStageBytes + (A_RES ? 0u : kABlockBytes);
    const uint32_t sA = smem_base;
    const uint32_t sStage = smem_base + a_bytes;
    const uint32_t sBar = sStage + (uint32_t)P.stages * stage_bytes;
    // barriers (8 bytes each): full[stages], empty[stages], tmem_full[2], tmem_empty[2], a_full, a_empty
    const uint32_t bar_full = sBar, bar_empty = sBar + 8u * kMaxStages;
    const uint32_t bar_tfull = sBar + 16u * kMaxStages, bar_tempty = bar_tfull + 16u;
    const uint32_t bar_afull = bar_tempty + 16u, bar_aempty = bar_afull + 8u;
    const uint32_t tmem_slot = bar_aempty + 8u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, MC ? 2 : 1); }  // MC: a commit of each CTA
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8u * a, 1); mbar_init(bar_tempty + 8u * a, 8); }  // one arrival per epilogue warp
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();  // the peer's barriers exist before anything signals them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // work units: (row block, column split); MC: (pair of row blocks, column split), this CTA takes block 2p + rank
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    const int64_t u_first = MC ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
    const int64_t u_step = MC ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
    const int64_t n_rbu = MC ? P.n_row_blocks / 2 : P.n_row_blocks;  // (the plan pads the row blocks to whole pairs)
    const int64_t n_units = n_rbu * P.n_splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
            int stage = 0;
            uint32_t phase = 0, uphase = 0;
            for (int64_t u = u_first; u < n_units; u += u_step) {
                const int64_t rb = MC ? 2 * (u % n_rbu) + rank : u % n_rbu;
                const int sp = (int)(u / n_rbu);
                const int64_t t0 = (int64_t)sp * P.tiles_per_split;
                const int64_t t1 = min(t0 + P.tiles_per_split, P.n_tiles);
                if (A_RES) {
                    mbar_wait_relaxed(bar_aempty, uphase ^ 1u);  // previous unit's MMAs have drained A
                    mbar_expect_tx(bar_afull, a_bytes);
                    for (int kb = 0; kb < num_kb; ++kb)
                        tma_load_2d(sA + kb * kABlockBytes, &tmap_a, bar_afull, kb * BK, (int)(rb * BM));
                    uphase ^= 1u;
                }
                for (int64_t t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait_relaxed(bar_empty + 8u * stage, phase ^ 1u);
                        const uint32_t dst = sStage + stage * stage_bytes;
                        if (K1_DIAG(2) && (t > t0 || kb >= P.stages || u != u_first)) {  // (diagnosis: the MMAs re-read what the ring holds)
                            mbar_arrive(bar_full + 8u * stage);
                        } else {
                            mbar_expect_tx(bar_full + 8u * stage, stage_bytes);  // (MC: half of it arrives from the peer)
                            if (MC) tma_load_2d_mc(dst + rank * (kBStageBytes / 2), &tmap_b, bar_full + 8u * stage, kb * BK,
                                                   (int)(t * BN + rank * (BN / 2)));
                            else tma_load_2d(dst, &tmap_b, bar_full + 8u * stage, kb * BK, (int)(t * BN));
                            if (!A_RES)
                                tma_load_2d(dst + kBStageBytes, &tmap_a, bar_full + 8u * stage, kb * BK, (int)(rb * BM));
                        }
                        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, uphase = 0;
            for (int64_t u = u_first; u < n_units; u += u_step) {
                const int sp = (int)(u / n_rbu);
                const int64_t t0 = (int64_t)sp * P.tiles_per_split;
                const int64_t t1 = min(t0 + P.tiles_per_split, P.n_tiles);
                if (A_RES) { mbar_wait(bar_afull, uphase); uphase ^= 1u; }
                for (int64_t t = t0; t < t1; ++t) {
                    mbar_wait_relaxed(bar_tempty + 8u * acc, acc_phase ^ 1u);  // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(bar_full + 8u * stage, phase);
                        tc_fence_after();
                        const uint32_t b_addr = sStage + stage * stage_bytes;
                        const uint32_t a_addr = A_RES ? sA + kb * kABlockBytes : b_addr + kBStageBytes;
                        const uint64_t adesc = umma_desc_sw128(a_addr);
                        const uint64_t bdesc = umma_desc_sw128(b_addr);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)  // +32 bytes per K=16 step inside the swizzle atom
                            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                        (uint32_t)((kb | k) != 0));
                        if (MC) tc_commit_mc(bar_empty + 8u * stage);  // frees the stage (in both CTAs) when these MMAs retire
                        else tc_commit(bar_empty + 8u * stage);
                        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit(bar_tfull + 8u * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
                if (A_RES) tc_commit(bar_aempty);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int quad = warp & 3;          // TMEM lane quadrant this warp may read
        const int half = (warp - 4) >> 2;   // which 128 accumulator columns
        const int r = quad * 32 + lane;     // image row within the block
        int acc = 0;
        uint32_t acc_phase = 0;
#ifdef MMALIGN_PROFILE_EPI
        long long prof[12] = {};
        const long long prof_t0 = clock64();
#endif
        for (int64_t u = u_first; u < n_units; u += u_step) {
            const int64_t rb = MC ? 2 * (u % n_rbu) + rank : u % n_rbu;
            const int sp = (int)(u / n_rbu);
            const int64_t t0 = (int64_t)sp * P.tiles_per_split;
            const int64_t t1 = min(t0 + P.tiles_per_split, P.n_tiles);
            const int64_t list_id = (((int64_t)sp * P.n_row_blocks + rb) * 2 + half) * 128 + r;
            uint2 *const list = reinterpret_cast<uint2 *>(P.keys + list_id * CAP);  // this thread's candidate list
            float tau = P.tau_init;
            int n = 0;
            int pend = -1, pend_cnt = 0;  // split compaction in flight: the lane whose list was snapshot, its length then
            uint2 pend_e[KPL <= 8 ? KPL : 1];
            const int64_t row = rb * BM + r;
            for (int64_t t = t0; t < t1; ++t) {
                EPI_T(w0_);
                mbar_wait_paused(bar_tfull + 8u * acc, acc_phase, P.epi_sleep_ns);
                EPI_T(w1_);
                EPI_ADD(0, w1_ - w0_);  // waiting for the tensor pipe
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * 128);
                const int64_t col0 = t * BN + half * 128;
                if (K1_DIAG(1) || (K1_DIAG(8) && quad == 1) || (K1_DIAG(16) && quad == 2)) {  // (diagnosis: the contraction alone; 8 / 16: without
                                                                                                  //  the epilogue warps of quadrant 1 (the MMA warp's scheduler) / 2)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
                } else if (K1_DIAG(32)) {  // (diagnosis: the accumulator is read, nothing is done with it)
                    uint32_t va[32], vb[32];
                    __syncwarp();
                    tmem_ld32(taddr, va); tmem_ld32(taddr + 32, vb);
                    tmem_ld_wait(va); tmem_ld_wait(vb);
                    uint32_t x = va[0] ^ vb[31];
                    tmem_ld32(taddr + 64, va); tmem_ld32(taddr + 96, vb);
                    tmem_ld_wait(va); tmem_ld_wait(vb);
                    x ^= va[5] ^ vb[7];
                    if (x == 0x12345678u) n = 1;  // (keeps the loads)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
                } else if (P.dump) {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t v[32];
                        __syncwarp();
                        tmem_ld32(taddr + ch * 32, v);
                        tmem_ld_wait(v);
                        if (row < P.N)
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (col0 + ch * 32 + k < P.M) P.dump[row * P.M + col0 + ch * 32 + k] = __uint_as_float(v[k]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
                } else {
                    // ping-pong: the load of chunk c+1 is in flight while chunk c is filtered
                    uint32_t va[32], vb[32];
                    __syncwarp();
                    EPI_T(c0_);
                    tmem_ld32(taddr, va);
                    tmem_ld_wait(va);
                    EPI_T(c1_);
                    tmem_ld32(taddr + 32, vb);
                    process_chunk<KPL>(va, col0, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c2_);
                    tmem_ld_wait(vb);
                    EPI_T(c3_);
                    tmem_ld32(taddr + 64, va);
                    process_chunk<KPL>(vb, col0 + 32, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c4_);
                    tmem_ld_wait(va);
                    EPI_T(c5_);
                    tmem_ld32(taddr + 96, vb);
                    process_chunk<KPL>(va, col0 + 64, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c6_);
                    tmem_ld_wait(vb);
                    EPI_T(c7_);
                    tc_fence_before();  // last read of this accumulator: hand it back to the MMA warp (every lane's loads
                    __syncwarp();       // have completed; one arrival per warp)
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
                    process_chunk<KPL>(vb, col0 + 96, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    // Routine compaction happens HERE, after the accumulator went back to the MMA warp, so that its
                    // global-memory latency is off the MMA critical path (the check inside process_chunk only fires
                    // when a single tile overflows the remaining room, i.e. in the first tiles of a sweep).
                    __syncwarp();
                    EPI_T(c8_);
                    bool split = false;
                    if constexpr (KPL <= 8) {
                        if (P.compact_one == 2) {
                            split = true;
                            if (pend >= 0) {
                                compact_complete<KPL>(list, n, tau, P.kprime, pend, pend_cnt, pend_e);
                                EPI_ADD(5, 1);
                            }
                            compact_issue<KPL>(list, n, n > CAP - kCompactMargin, pend, pend_cnt, pend_e);
                        }
                    }
                    if (!split && __any_sync(0xFFFFFFFFu, n > CAP - kCompactMargin)) {
                        compact_lists<KPL>(list, n, tau, P.kprime, n > CAP - kCompactMargin, P.compact_one != 0);
                        EPI_ADD(5, 1);
                    }
                    EPI_T(c9_);
                    EPI_ADD(1, (c1_ - c0_) + (c3_ - c2_) + (c5_ - c4_) + (c7_ - c6_));  // waiting for TMEM loads
                    EPI_ADD(2, (c2_ - c1_) + (c4_ - c3_) + (c6_ - c5_) + (c8_ - c7_));  // filtering (incl. appends)
                    EPI_ADD(3, c9_ - c8_);                                              // routine compaction
                    EPI_ADD(4, 1);                                                      // tiles
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (!P.dump && !P.skip_final) {  // leave at most K' + slack entries per list for the rescoring kernel
                __syncwarp();
                pend = -1;
                if (__any_sync(0xFFFFFFFFu, n > P.kprime + kSlack))
                    compact_lists<KPL>(list, n, tau, P.kprime, n > P.kprime + kSlack);
            }
            if (!P.dump) {
                MMA_CHECK(list_id >= 0 && list_id < P.n_row_blocks * P.n_splits * 256 && n <= CAP);
                P.tau[list_id] = tau;
                P.count[list_id] = n;
            }
        }
#ifdef MMALIGN_PROFILE_EPI
        if (lane == 0) {
            prof[6] = clock64() - prof_t0;
            for (int q = 0; q < 12; ++q) if (q != 7) atomicAdd(&g_epi_prof[q], (unsigned long long)prof[q]);
            atomicAdd(&g_epi_prof[7], 1ull);
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();  // neither CTA leaves while the other may still write its ring or signal its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int KPL, bool A_RES>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_score_topk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const FusedArgs P)
{
    fused_body<KPL, A_RES, false>(tmap_a, tmap_b, P);
}

template <int KPL, bool A_RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFusedThreads, 1)
fused_score_topk_mc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                           const FusedArgs P)
{
    fused_body<KPL, A_RES, true>(tmap_a, tmap_b, P);
}


// ---------------------------------------------------------------------------
// The same kernel on CTA pairs: tcgen05.mma.cta_group::2, M = 256 (two 128-image row blocks, one per CTA of a
// 2-CTA cluster = the two SMs of a TPC), N = 256.  Each CTA holds its own A rows and HALF of every B k-slice
// (128 chunks x 64): 16 KiB of B per stage and CTA instead of 32, so the ring is twice as deep in k and every B
// byte is read from L2 by one SM of the pair instead of both.
//   * both CTAs run a TMA producer (warp 0); every load signals the LEADER's (cluster rank 0) full barrier
//     (.cta_group::2 load, barrier address with the peer bit cleared); the leader expects the bytes of both
//   * only the leader issues MMAs (warp 1); its tcgen05.commit is multicast to the barriers of both CTAs
//     (stage free, accumulator ready, A rows free)
//   * both CTAs run the epilogue on their own TMEM (their 128 rows of the 256 x 256 accumulator); the 2 x 256
//     epilogue threads hand an accumulator back by arriving (one lane per warp) on the leader's tmem_empty barrier
// Row blocks are handed out in pairs (2p, 2p + 1); an odd count is padded with a phantom block whose rows TMA fills
// with zeros and whose lists nobody reads.
// ---------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // shared::cluster address of the same offset in cluster rank 0
constexpr uint32_t kBHalfBytes = (BN / 2) * BK * 2;  // 16 KiB: this CTA's half of a B k-slice
constexpr uint32_t kIdescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)((2 * BM) >> 4) << 24);

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void *tmap, uint32_t leader_bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(leader_bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar)  // arrives on `bar` in both CTAs when the MMAs retire
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint32_t bar)  // on the barrier at this offset in cluster rank 0
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(0u) : "memory");
}

template <int KPL, bool A_RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFusedThreads, 1)
fused_score_topk_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                             const FusedArgs P)
{
    constexpr int CAP = 32 * KPL;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_kb = P.num_kb;
    const uint32_t a_bytes = A_RES ? (uint32_t)num_kb * kABlockBytes : 0u;
    const uint32_t stage_bytes = kBHalfBytes + (A_RES ? 0u : kABlockBytes);
    const uint32_t sA = smem_base;
    const uint32_t sStage = smem_base + a_bytes;
    const uint32_t sBar = sStage + (uint32_t)P.stages * stage_bytes;
    const uint32_t bar_full = sBar, bar_empty = sBar + 8u * kMaxPairStages;
    const uint32_t bar_tfull = sBar + 16u * kMaxPairStages, bar_tempty = bar_tfull + 16u;
    const uint32_t bar_afull = bar_tempty + 16u, bar_aempty = bar_afull + 8u;
    const uint32_t tmem_slot = bar_aempty + 8u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8u * a, 1); mbar_init(bar_tempty + 8u * a, 16); }  // one arrival per epilogue warp of the pair
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist before anything signals them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int64_t n_pairs = P.n_row_blocks / 2;  // the plan pads n_row_blocks to an even count
    const int64_t n_units = n_pairs * P.n_splits;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
            int stage = 0;
            uint32_t phase = 0, uphase = 0;
            for (int64_t u = cluster_id; u < n_units; u += n_clusters) {
                const int64_t rb = 2 * (u % n_pairs) + rank;
                const int sp = (int)(u / n_pairs);
                const int64_t t0 = (int64_t)sp * P.tiles_per_split;
                const int64_t t1 = min(t0 + P.tiles_per_split, P.n_tiles);
                if (A_RES) {
                    mbar_wait_relaxed(bar_aempty, uphase ^ 1u);  // (own copy; the leader's commit reaches both)
                    if (leader) mbar_expect_tx(bar_afull, 2u * a_bytes);
                    for (int kb = 0; kb < num_kb; ++kb)
                        tma_load_2d_pair(sA + kb * kABlockBytes, &tmap_a, bar_afull, kb * BK, (int)(rb * BM));
                    uphase ^= 1u;
                }
                for (int64_t t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait_relaxed(bar_empty + 8u * stage, phase ^ 1u);
                        const uint32_t dst = sStage + stage * stage_bytes;
                        if (K1_DIAG(2) && (t > t0 || kb >= P.stages || u != cluster_id)) {  // (diagnosis: the MMAs re-read what the ring holds)
                            if (leader) mbar_arrive(bar_full + 8u * stage);
                        } else {
                            if (leader) mbar_expect_tx(bar_full + 8u * stage, 2u * stage_bytes);
                            tma_load_2d_pair(dst, &tmap_b, bar_full + 8u * stage, kb * BK, (int)(t * BN + rank * (BN / 2)));
                            if (!A_RES)
                                tma_load_2d_pair(dst + kBHalfBytes, &tmap_a, bar_full + 8u * stage, kb * BK, (int)(rb * BM));
                        }
                        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader only) =====================
        if (leader && lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, uphase = 0;
            for (int64_t u = cluster_id; u < n_units; u += n_clusters) {
                const int sp = (int)(u / n_pairs);
                const int64_t t0 = (int64_t)sp * P.tiles_per_split;
                const int64_t t1 = min(t0 + P.tiles_per_split, P.n_tiles);
                if (A_RES) { mbar_wait(bar_afull, uphase); uphase ^= 1u; }
                for (int64_t t = t0; t < t1; ++t) {
                    mbar_wait_relaxed(bar_tempty + 8u * acc, acc_phase ^ 1u);  // both epilogues have drained it
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(bar_full + 8u * stage, phase);
                        tc_fence_after();
                        const uint32_t b_addr = sStage + stage * stage_bytes;
                        const uint32_t a_addr = A_RES ? sA + kb * kABlockBytes : b_addr + kBHalfBytes;
                        const uint64_t adesc = umma_desc_sw128(a_addr);
                        const uint64_t bdesc = umma_desc_sw128(b_addr);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdescPair,
                                             (uint32_t)((kb | k) != 0));
                        tc_commit_pair(bar_empty + 8u * stage);
                        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit_pair(bar_tfull + 8u * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
                if (A_RES) tc_commit_pair(bar_aempty);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int quad = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
#ifdef MMALIGN_PROFILE_EPI
        long long prof[12] = {};
        const long long prof_t0 = clock64();
#endif
        for (int64_t u = cluster_id; u < n_units; u += n_clusters) {
            const int64_t rb = 2 * (u % n_pairs) + rank;
            const int sp = (int)(u / n_pairs);
            const int64_t t0 = (int64_t)sp * P.tiles_per_split;
            const int64_t t1 = min(t0 + P.tiles_per_split, P.n_tiles);
            const int64_t list_id = (((int64_t)sp * P.n_row_blocks + rb) * 2 + half) * 128 + r;
            uint2 *const list = reinterpret_cast<uint2 *>(P.keys + list_id * CAP);
            float tau = P.tau_init;
            int n = 0;
            int pend = -1, pend_cnt = 0;  // split compaction in flight: the lane whose list was snapshot, its length then
            uint2 pend_e[KPL <= 8 ? KPL : 1];
            const int64_t row = rb * BM + r;
            for (int64_t t = t0; t < t1; ++t) {
                EPI_T(w0_);
                mbar_wait_paused(bar_tfull + 8u * acc, acc_phase, P.epi_sleep_ns);
                EPI_T(w1_);
                EPI_ADD(0, w1_ - w0_);  // waiting for the tensor pipe
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * 128);
                const int64_t col0 = t * BN + half * 128;
                if (K1_DIAG(1)) {  // (diagnosis: the contraction alone)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_rank0(bar_tempty + 8u * acc);
                } else if (K1_DIAG(32)) {  // (diagnosis: the accumulator is read, nothing is done with it)
                    uint32_t va[32], vb[32];
                    __syncwarp();
                    tmem_ld32(taddr, va); tmem_ld32(taddr + 32, vb);
                    tmem_ld_wait(va); tmem_ld_wait(vb);
                    uint32_t x = va[0] ^ vb[31];
                    tmem_ld32(taddr + 64, va); tmem_ld32(taddr + 96, vb);
                    tmem_ld_wait(va); tmem_ld_wait(vb);
                    x ^= va[5] ^ vb[7];
                    if (x == 0x12345678u) n = 1;  // (keeps the loads)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_rank0(bar_tempty + 8u * acc);
                } else if (P.dump) {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t v[32];
                        __syncwarp();
                        tmem_ld32(taddr + ch * 32, v);
                        tmem_ld_wait(v);
                        if (row < P.N)
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (col0 + ch * 32 + k < P.M) P.dump[row * P.M + col0 + ch * 32 + k] = __uint_as_float(v[k]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_rank0(bar_tempty + 8u * acc);
                } else if (K1_DIAG(4)) {  // (diagnosis: what an epilogue that hands the accumulator back right after its loads would
                    uint32_t va[32], vb[32];  //  buy -- the chunks are processed from two buffers, so the lists are NOT valid)
                    __syncwarp();
                    tmem_ld32(taddr, va);
                    tmem_ld32(taddr + 32, vb);
                    tmem_ld_wait(va);
                    tmem_ld_wait(vb);
                    tmem_ld32(taddr + 64, va);
                    tmem_ld32(taddr + 96, vb);
                    tmem_ld_wait(va);
                    tmem_ld_wait(vb);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_rank0(bar_tempty + 8u * acc);
                    process_chunk<KPL>(va, col0, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    process_chunk<KPL>(vb, col0 + 32, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    for (int k = 0; k < 32; ++k) { va[k] ^= 0x00000100u; vb[k] ^= 0x00000100u; }
                    process_chunk<KPL>(va, col0 + 64, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    process_chunk<KPL>(vb, col0 + 96, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    if (__any_sync(0xFFFFFFFFu, n > CAP - kCompactMargin))
                        compact_lists<KPL>(list, n, tau, P.kprime, n > CAP - kCompactMargin);
                } else {
                    uint32_t va[32], vb[32];
                    __syncwarp();
                    EPI_T(c0_);
                    tmem_ld32(taddr, va);
                    tmem_ld_wait(va);
                    EPI_T(c1_);
                    tmem_ld32(taddr + 32, vb);
                    process_chunk<KPL>(va, col0, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c2_);
                    tmem_ld_wait(vb);
                    EPI_T(c3_);
                    tmem_ld32(taddr + 64, va);
                    process_chunk<KPL>(vb, col0 + 32, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c4_);
                    tmem_ld_wait(va);
                    EPI_T(c5_);
                    tmem_ld32(taddr + 96, vb);
                    process_chunk<KPL>(va, col0 + 64, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c6_);
                    tmem_ld_wait(vb);
                    EPI_T(c7_);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_rank0(bar_tempty + 8u * acc);  // 2 x 8 arrivals free the accumulator for the pair
                    process_chunk<KPL>(vb, col0 + 96, P.M, list, n, tau, P.kprime, P.col_base, pend PC_PROF_ARG);
                    __syncwarp();
                    EPI_T(c8_);
                    bool split = false;
                    if constexpr (KPL <= 8) {
                        if (P.compact_one == 2) {
                            split = true;
                            if (pend >= 0) {
                                compact_complete<KPL>(list, n, tau, P.kprime, pend, pend_cnt, pend_e);
                                EPI_ADD(5, 1);
                            }
                            compact_issue<KPL>(list, n, n > CAP - kCompactMargin, pend, pend_cnt, pend_e);
                        }
                    }
                    if (!split && __any_sync(0xFFFFFFFFu, n > CAP - kCompactMargin)) {
                        compact_lists<KPL>(list, n, tau, P.kprime, n > CAP - kCompactMargin, P.compact_one != 0);
                        EPI_ADD(5, 1);
                    }
                    EPI_T(c9_);
                    EPI_ADD(1, (c1_ - c0_) + (c3_ - c2_) + (c5_ - c4_) + (c7_ - c6_));  // waiting for TMEM loads
                    EPI_ADD(2, (c2_ - c1_) + (c4_ - c3_) + (c6_ - c5_) + (c8_ - c7_));  // filtering (incl. appends)
                    EPI_ADD(3, c9_ - c8_);                                              // routine compaction
                    EPI_ADD(4, 1);                                                      // tiles
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (!P.dump && !P.skip_final) {
                __syncwarp();
                pend = -1;
                if (__any_sync(0xFFFFFFFFu, n > P.kprime + kSlack))
                    compact_lists<KPL>(list, n, tau, P.kprime, n > P.kprime + kSlack);
            }
            if (!P.dump) {
                MMA_CHECK(list_id >= 0 && list_id < P.n_row_blocks * P.n_splits * 256 && n <= CAP);
                P.tau[list_id] = tau;
                P.count[list_id] = n;
            }
        }
#ifdef MMALIGN_PROFILE_EPI
        if (lane == 0) {
            prof[6] = clock64() - prof_t0;
            for (int q = 0; q < 12; ++q) if (q != 7) atomicAdd(&g_epi_prof[q], (unsigned long long)prof[q]);
            atomicAdd(&g_epi_prof[7], 1ull);
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA leaves (or frees its TMEM) while the other may still signal its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
int fused_plan(int64_t N, int64_t M, int D, int kneed, int kprime_req, int sm_count, int n_ranks, FusedPlan *plan, int cluster)
{
    if (D % BK != 0 || D < BK || N <= 0 || M <= 0) return -1;
    FusedPlan p = {};
    p.n_row_blocks = (N + BM - 1) / BM;
    p.pairs = cluster != 0 && sm_count % 2 == 0;  // clusters of two CTAs: whole pairs of row blocks, an even grid
    p.mc = p.pairs && cluster == 2;                // ... that share the B ring by multicast (cta_group::1 MMAs)
    if (p.pairs) p.n_row_blocks = (p.n_row_blocks + 1) & ~(int64_t)1;  // whole pairs; a phantom block's lists are never read
    const int64_t n_tiles = (M + BN - 1) / BN;
    // column splits: the fewest that keep the last wave of persistent CTAs >= 95 % full
    // (up to 16 splits keep a row's 2 x splits lists within what the warp-per-row rescoring takes, rescore.cu; more only
    // when they are needed to fill the GPU at all)
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64 && s <= n_tiles; ++s) {
        if (s == 17 && best_eff >= 0.85) break;
        const int64_t units = p.n_row_blocks * s;
        const int64_t waves = (units + sm_count - 1) / sm_count;
        const double eff = (double)units / (double)(waves * sm_count);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
        if (eff >= 0.95) { best = s; break; }
    }
    p.n_splits = best;
    const int64_t tps = (n_tiles + p.n_splits - 1) / p.n_splits;
    p.n_splits = (int)((n_tiles + tps - 1) / tps);
    p.tiles_per_split = (int)tps;
    // Depth to which the UNION of a row's lists must be complete.  The certificate of rescore.cu needs
    // (exact kneed-th score) - (approximate K'-th score) > eps ~ 0.004 (bf16 rounding bound); for
    // near-isotropic embeddings at D >= 512 that takes K' ~ 1.9 x kneed (tools/gap_diag.py).
    p.kprime = kprime_req > 0 ? kprime_req : (kneed * 48 / 25 > kneed + 50 ? kneed * 48 / 25 : kneed + 50);
    if (p.kprime < kneed) p.kprime = kneed;
    // A row has 2 * n_splits lists over disjoint columns on each of the n_ranks GPUs; the union of their best
    // k entries is complete to a depth of about n_lists * k, so each list keeps its share plus 15 % for imbalance.
#ifdef MMALIGN_TUNING  // tuning builds only: size the lists as for a sharded run
    if (const char *e = getenv("MMALIGN_PLAN_RANKS")) n_ranks = atoi(e);
#endif
    const int lists_per_row = 2 * p.n_splits * (n_ranks > 1 ? n_ranks : 1);
    // The union is complete above the LARGEST of the lists' thresholds.  A list's k-th best score sits at
    // global rank ~ L*k with relative spread 1/sqrt(k), and the largest of L of them about z_L = sqrt(2 ln L)
    // spreads above the mean, so the union's depth is ~ L*k*(1 - z_L/sqrt(k)); solve that for depth = K'.
    {
        const double L = (double)lists_per_row, share = (double)p.kprime / L;
        const double z = sqrt(2.0 * log(L)) + 0.4;  // + 0.4: one uncertified row in 1e6 still costs a rescan launch
        const double rk = 0.5 * (z + sqrt(z * z + 4.0 * share));
        p.kprime_list = (int)ceil(rk * rk) + 2;
    }
    if (p.kprime_list > p.kprime) p.kprime_list = p.kprime;
    if (p.kprime_list < 16) p.kprime_list = 16;
    // A list is compacted back to kprime_list (+ kSlack) entries when fewer than kCompactMargin slots are free after a
    // tile; the room between those two marks is how many insertions one compaction buys.  Measured (K1 alone): 6 slots
    // of room (the 4-GPU share in 128-entry lists) cost 6.5 % against 64; 300 slots in 512-entry lists gain nothing.
    int room = 64;
    bool room_forced = false;
#ifdef MMALIGN_TUNING
    if (const char *e = getenv("MMALIGN_CAP_ROOM")) { room = atoi(e); room_forced = true; }
#endif
    int need = p.kprime_list + kSlack + kCompactMargin + room;
    if (need > 512) need = p.kprime_list + kSlack + kCompactMargin + 32;
    // A few slots short of the full room is worth the next smaller capacity (half the keys per lane in every
    // compaction, half the list memory): 48 slots of room are accepted for it.
    if (!room_forced && need > 256 && need - room + 48 <= 256) need = 256;
    if (!room_forced && need > 128 && need - room + 48 <= 128) need = 128;
    if (need <= 128) p.cap = 128;
    else if (need <= 256) p.cap = 256;
    else if (need <= 512) p.cap = 512;
    else return -2;
    p.n_lists = p.n_row_blocks * p.n_splits * 256;
    p.a_resident = D <= 512;
#ifdef MMALIGN_TUNING
    if (const char *e = getenv("MMALIGN_A_RESIDENT")) p.a_resident = p.a_resident && atoi(e) != 0;
#endif
    const size_t a_bytes = p.a_resident ? (size_t)(D / BK) * kABlockBytes : 0;
    const bool pair_mma = p.pairs && !p.mc;  // the cta_group::2 kernel: half a B k-slice per CTA and stage
    const size_t stage_bytes = (pair_mma ? kBHalfBytes : kBStageBytes) + (p.a_resident ? 0 : kABlockBytes);
    const size_t fixed = 1024 /*alignment slack*/ + a_bytes + 256 /*barriers*/;
    int stages = (int)((kSmemLimit - fixed) / stage_bytes);
    if (stages > (pair_mma ? kMaxPairStages : kMaxStages)) stages = pair_mma ? kMaxPairStages : kMaxStages;
    if (stages < 2) return -3;
    p.stages = stages;
    p.smem_bytes = fixed + (size_t)stages * stage_bytes;
    const int64_t units = p.n_row_blocks * p.n_splits;
    p.grid = (int)(units < sm_count ? units : sm_count);
    if (p.pairs) p.grid &= ~1;  // whole clusters
    *plan = p;
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int encode_tensor_map(void *tmap_out, const void *base, int64_t rows, int D, int box_rows, char *err,
                      size_t errlen)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            snprintf(err, errlen, "cuTensorMapEncodeTiled entry point unavailable");
            return -1;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)D * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(reinterpret_cast<CUtensorMap *>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                          const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return -1;
    }
    return 0;
}

template <int KPL, bool A_RES>
static cudaError_t launch_pair_variant(const CUtensorMap &ta, const CUtensorMap &tb, const FusedArgs &args,
                                       const FusedPlan &plan, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(fused_score_topk_pair_kernel<KPL, A_RES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return e;
    fused_score_topk_pair_kernel<KPL, A_RES><<<plan.grid, kFusedThreads, plan.smem_bytes, st>>>(ta, tb, args);  // __cluster_dims__(2)
    return cudaGetLastError();
}

template <int KPL, bool A_RES>
static cudaError_t launch_mc_variant(const CUtensorMap &ta, const CUtensorMap &tb, const FusedArgs &args,
                                     const FusedPlan &plan, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(fused_score_topk_mc_kernel<KPL, A_RES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return e;
    fused_score_topk_mc_kernel<KPL, A_RES><<<plan.grid, kFusedThreads, plan.smem_bytes, st>>>(ta, tb, args);  // __cluster_dims__(2)
    return cudaGetLastError();
}

template <int KPL, bool A_RES>
static cudaError_t launch_variant(const CUtensorMap &ta, const CUtensorMap &tb, const FusedArgs &args,
                                  const FusedPlan &plan, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(fused_score_topk_kernel<KPL, A_RES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return e;
    fused_score_topk_kernel<KPL, A_RES><<<plan.grid, kFusedThreads, plan.smem_bytes, st>>>(ta, tb, args);
    return cudaGetLastError();
}

// `lists` are the buffers of this launch (a launch over a column group of the table passes its slice of a row's
// lists and the group's first column as col_base; fused_tc.cu numbers columns from the start of tensor map B)
cudaError_t launch_fused(const Side &img, const Side &chk, const FusedPlan &plan, const void *tmap_a,
                         const void *tmap_b, CandLists &lists, float *dump, cudaStream_t st, int64_t col_base)
{
    FusedArgs a = {};
    a.N = img.n; a.M = chk.n; a.num_kb = img.D / BK;
    a.n_row_blocks = plan.n_row_blocks;
    a.n_tiles = (chk.n + BN - 1) / BN;
    a.n_splits = plan.n_splits;
    a.tiles_per_split = plan.tiles_per_split;
    a.stages = plan.stages;
    a.keys = lists.keys; a.tau = lists.tau; a.count = lists.count;
    a.kprime = plan.kprime_list;
    a.dump = dump;
    a.tau_init = -INFINITY;
    a.skip_final = 0;
#ifdef MMALIGN_TUNING
    if (const char *e = getenv("MMALIGN_TAU_INIT")) a.tau_init = (float)atof(e);
    a.skip_final = getenv("MMALIGN_SKIP_FINAL") != nullptr;
    if (const char *e = getenv("MMALIGN_K1_DIAG")) a.diag = atoi(e);
    {
        const int hm = getenv("MMALIGN_K1_HIT") ? atoi(getenv("MMALIGN_K1_HIT")) : 0;
        cudaMemcpyToSymbolAsync(g_hit_mode, &hm, sizeof hm, 0, cudaMemcpyHostToDevice, st);
    }
#endif
    a.col_base = (uint32_t)col_base;
    a.epi_sleep_ns = (uint32_t)plan.epi_sleep_ns;
    a.compact_one = plan.compact_one;
    const CUtensorMap &ta = *reinterpret_cast<const CUtensorMap *>(tmap_a);
    const CUtensorMap &tb = *reinterpret_cast<const CUtensorMap *>(tmap_b);
    lists.cap = plan.cap; lists.n_splits = plan.n_splits; lists.n_row_blocks = plan.n_row_blocks;
    lists.kprime = plan.kprime;
    lists.kprime_list = plan.kprime_list;
#define VARIANT(KPL) (plan.mc    ? (plan.a_resident ? launch_mc_variant<KPL, true>(ta, tb, a, plan, st)        \
                                                    : launch_mc_variant<KPL, false>(ta, tb, a, plan, st))   \
                      : plan.pairs ? (plan.a_resident ? launch_pair_variant<KPL, true>(ta, tb, a, plan, st)    \
                                                    : launch_pair_variant<KPL, false>(ta, tb, a, plan, st)) \
                                  : (plan.a_resident ? launch_variant<KPL, true>(ta, tb, a, plan, st)           \
                                                    : launch_variant<KPL, false>(ta, tb, a, plan, st)))
    switch (plan.cap) {
    case 128: return VARIANT(4);
    case 256: return VARIANT(8);
    case 512: return VARIANT(16);
    }
#undef VARIANT
    return cudaErrorInvalidValue;
}

MMA_CHECK_READER(check_read_fused)

} // namespace mma
