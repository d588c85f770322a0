// gather_probe.cu -- what the memory system gives to the access pattern of the exact rescoring (K2): random gathers
// of whole embedding rows (D fp32 = 2 KB at D=512) out of a table far larger than L2, one warp per gather stream,
// R rows in flight per warp, W warps per SM.  The figure it prints is the ceiling K2's gathers are judged against
// (DESIGN.md section 4); the streaming-copy peak of MEASURED_PEAKS.json is not reachable by random 2 KB reads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu && tools/gather_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int R>
__global__ void gather_kernel(const float4 *__restrict__ table, const int32_t *__restrict__ idx, int64_t n_idx, int d4,
                              float *out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t e = warp * R; e + R <= n_idx; e += n_warps * R) {
        const float4 *row[R];
#pragma unroll
        for (int q = 0; q < R; ++q) row[q] = table + (int64_t)idx[e + q] * d4;
        for (int c = lane; c < d4; c += 128) {
            float4 y[R][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c + 32 * u < d4) {
#pragma unroll
                    for (int q = 0; q < R; ++q) y[q][u] = __ldg(row[q] + c + 32 * u);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c + 32 * u < d4) {
#pragma unroll
                    for (int q = 0; q < R; ++q) acc += y[q][u].x + y[q][u].y + y[q][u].z + y[q][u].w;
                }
        }
    }
    if (acc == 123.456f) out[0] = acc;  // (keeps the loads)
}

template <int R>
static int run(const float4 *table, const int32_t *idx, int64_t n_idx, int D, int warps_per_sm, int sms, float *out)
{
    const int threads = 128, ctas = sms * warps_per_sm / 4;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        gather_kernel<R><<<ctas, threads>>>(table, idx, n_idx, D / 4, out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    const double bytes = (double)n_idx * D * 4;
    printf("{\"D\": %d, \"rows_in_flight_per_warp\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"GBps\": %.1f}\n", D, R,
           warps_per_sm, best, bytes / best / 1e6);
    fflush(stdout);
    return 0;
}

int main(int argc, char **argv)
{
    const int64_t M = argc > 1 ? atoll(argv[1]) : 1000000;
    const int64_t n_idx = argc > 2 ? atoll(argv[2]) : 32000000;
    int dev_sms = 0;
    CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<int32_t> h((size_t)n_idx);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int64_t i = 0; i < n_idx; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        h[(size_t)i] = (int32_t)(s % (uint64_t)M);
    }
    int32_t *idx;
    float *out;
    CK(cudaMalloc(&idx, n_idx * 4));
    CK(cudaMalloc(&out, 4));
    CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
    for (int D : {512, 1024}) {
        float4 *table;
        CK(cudaMalloc(&table, (size_t)M * D * 4));
        CK(cudaMemset(table, 0, (size_t)M * D * 4));
        const int64_t n = D == 512 ? n_idx : n_idx / 2;
        for (int w : {16, 32, 64}) {
            if (run<2>(table, idx, n, D, w, dev_sms, out)) return 1;
            if (run<4>(table, idx, n, D, w, dev_sms, out)) return 1;
            if (w <= 32 && run<8>(table, idx, n, D, w, dev_sms, out)) return 1;
        }
        CK(cudaFree(table));
    }
    return 0;
}
