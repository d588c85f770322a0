// prep.cu -- K0: operand preparation and the same-page pair index.
//
//  * prep_rows_kernel: per row, canonical sum of squares (the |a|^2 of pgvector's
//    cosine, src/evaluate_alignments.py:97), L2 normalisation
//    (src/insert_clip_embeddings.py:113-115), bf16 rounding for the tensor-core
//    operand, and the norm of the rounding error (certificate of rescore.cu).
//    HBM-bound: reads 4*D bytes, writes 2*D + 8 bytes per row.
//  * build_pair_index: the join of src/evaluate_alignments.py:57-63 as a CSR --
//    chunks radix-sorted by page key (stable, so ties keep index order), each
//    image binary-searches its page, an exclusive scan gives the pair offsets.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace mma {
MMA_CHECK_DECL

// SM count of the current device (grids are sized in multiples of it); cached per device
int sm_count()
{
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

__global__ void __launch_bounds__(256)
prep_rows_kernel(const float *__restrict__ emb, int64_t n, int D, __nv_bfloat16 *__restrict__ out,
                 float *__restrict__ norm2, float *__restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int d4 = D >> 2;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n;
         row += (int64_t)gridDim.x * 8) {
        const float4 *a = reinterpret_cast<const float4 *>(emb + row * D);
        const float n2 = warp_dot(a, a, d4, lane);
        const float inv = n2 > 0.f ? 1.0f / sqrtf(n2) : 0.f;
        float e2 = 0.f;
        uint2 *o = reinterpret_cast<uint2 *>(out + row * D);
        for (int c = lane; c < d4; c += 32) {
            const float4 x = a[c];
            const float y0 = x.x * inv, y1 = x.y * inv, y2 = x.z * inv, y3 = x.w * inv;
            const __nv_bfloat162 q01 = __floats2bfloat162_rn(y0, y1);
            const __nv_bfloat162 q23 = __floats2bfloat162_rn(y2, y3);
            const float d0 = y0 - __low2float(q01), d1 = y1 - __high2float(q01);
            const float d2 = y2 - __low2float(q23), d3 = y3 - __high2float(q23);
            e2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            uint2 w;
            w.x = *reinterpret_cast<const uint32_t *>(&q01);
            w.y = *reinterpret_cast<const uint32_t *>(&q23);
            o[c] = w;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) e2 += __shfl_xor_sync(0xFFFFFFFFu, e2, off);
        if (lane == 0) {
            norm2[row] = n2;
            // upper bound: fp32 summation slack + the <= 3 ulp of the fp32 normalisation itself
            err[row] = sqrtf(e2) * 1.0005f + 2e-6f;
        }
    }
}

// K0 of rows [row0, row0 + rows) of a side (a piece of an upload, or a query slab)
cudaError_t launch_prep(const Side &s, int64_t row0, int64_t rows, int sm_count, cudaStream_t st)
{
    if (rows <= 0) return cudaSuccess;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    prep_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(s.emb + row0 * s.D, rows, s.D, s.emb_bf16 + row0 * s.D,
                                                       s.norm2 + row0, s.err + row0);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
__global__ void max_float_kernel(const float *__restrict__ x, int64_t n, float *out)
{
    __shared__ float sm[32];
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, x[i]);
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
        // values are >= 0: the integer order of the bit patterns is the float order
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));
    }
}

cudaError_t reduce_max_float(const float *x, int64_t n, float *out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
    if (e != cudaSuccess || n == 0) return e;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
    max_float_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
__global__ void iota_kernel(int32_t *v, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        v[i] = (int32_t)i;
}

__global__ void page_range_kernel(const uint64_t *__restrict__ img_key, int64_t N,
                                  const uint64_t *__restrict__ sorted_key, int64_t M,
                                  int64_t *__restrict__ sp_start, int64_t *__restrict__ counts,
                                  unsigned long long *__restrict__ c_max)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= N;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i == N) { counts[N] = 0; continue; }
        const uint64_t k = img_key[i];
        int64_t lo = 0, hi = M;
        while (lo < hi) {  // lower bound
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_key[mid] < k) lo = mid + 1; else hi = mid;
        }
        const int64_t first = lo;
        hi = M;
        while (lo < hi) {  // upper bound
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_key[mid] <= k) lo = mid + 1; else hi = mid;
        }
        sp_start[i] = first;
        MMA_CHECK(first >= 0 && first <= lo && lo <= M);
        const int64_t cnt = (k == MMALIGN_NULL_KEY) ? 0 : lo - first;  // SQL NULL never joins
        counts[i] = cnt;
        if (cnt > 0) atomicMax(c_max, (unsigned long long)cnt);
    }
}

// scratch layout: sorted keys [M] u64 | iota [M] i32 | counts [N+2] i64 | CUB temp
size_t pair_index_scratch_bytes(int64_t N, int64_t M)
{
    const int64_t Mx = M > 0 ? M : 1;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int)Mx, 0, 64, 0);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int64_t *)nullptr, (int64_t *)nullptr, (int)(N + 1), 0);
    const size_t tmp = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return up(sizeof(uint64_t) * Mx) + up(sizeof(int32_t) * Mx) + up(sizeof(int64_t) * (N + 2)) + up(tmp) + 256;
}

cudaError_t build_pair_index(const Side &img, const Side &chk, PairIndex &px, void *scratch, size_t scratch_bytes,
                             cudaStream_t st)
{
    const int64_t N = img.n, M = chk.n;
    const int64_t Mx = M > 0 ? M : 1;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char *base = static_cast<char *>(scratch);
    uint64_t *sorted_key = reinterpret_cast<uint64_t *>(base); base += up(sizeof(uint64_t) * Mx);
    int32_t *iota = reinterpret_cast<int32_t *>(base); base += up(sizeof(int32_t) * Mx);
    int64_t *counts = reinterpret_cast<int64_t *>(base); base += up(sizeof(int64_t) * (N + 2));
    void *tmp = base;
    size_t tmp_bytes = scratch_bytes - (size_t)(base - static_cast<char *>(scratch));
    cudaError_t e;
#define CK(x) do { e = (x); if (e != cudaSuccess) return e; } while (0)
    CK(cudaMemsetAsync(counts + N + 1, 0, sizeof(int64_t), st));
    if (M > 0) {
        iota_kernel<<<(unsigned)((M + 255) / 256 > sm_count() * 8 ? sm_count() * 8 : (M + 255) / 256), 256, 0, st>>>(iota, M);
        CK(cudaGetLastError());
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, chk.key, sorted_key, iota, px.sorted_chunk, (int)M, 0, 64, st));
    }
    page_range_kernel<<<(unsigned)((N + 256) / 256 > sm_count() * 8 ? sm_count() * 8 : (N + 256) / 256), 256, 0, st>>>(
        img.key, N, sorted_key, M, px.sp_start, counts, reinterpret_cast<unsigned long long *>(counts + N + 1));
    CK(cudaGetLastError());
    CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, px.offsets, (int)(N + 1), st));
    CK(cudaMemcpyAsync(&px.P, px.offsets + N, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&px.c_max, counts + N + 1, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
#undef CK
    return cudaSuccess;
}

MMA_CHECK_READER(check_read_prep)

} // namespace mma
