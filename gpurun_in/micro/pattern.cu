// DRAM access-pattern probe: what copy bandwidth does the step kernel's access pattern allow?
//   pattern 0: column records of 18 population rows of Hp reals (the engine's layout): a CTA moves 18 chunks of 1 KB, 8 KB apart
//   pattern 1: tiled layout [x][strip][pop][256]: a CTA moves one contiguous 18 KB block per column
#include <cstdio>
#include <cuda_runtime.h>
template <int PATTERN>
__global__ void __launch_bounds__(128, 3) k(const float2 *__restrict__ src, float2 *__restrict__ dst, int Hp, int W, int nyt, int chunk)
{
    const int yt = blockIdx.x % nyt, xs = (blockIdx.x / nyt) * chunk, xe = min(W, xs + chunk);
    const int t = threadIdx.x;
    for (int x = xs; x < xe; ++x) {
        float2 v[18];
#pragma unroll
        for (int p = 0; p < 18; ++p) {
            size_t i = PATTERN == 0 ? (((size_t)x * 18 + p) * Hp + yt * 256) / 2 + t : ((((size_t)x * nyt + yt) * 18 + p) * 256) / 2 + t;
            v[p] = src[i];
        }
#pragma unroll
        for (int p = 0; p < 18; ++p) {
            size_t i = PATTERN == 0 ? (((size_t)x * 18 + p) * Hp + yt * 256) / 2 + t : ((((size_t)x * nyt + yt) * 18 + p) * 256) / 2 + t;
            v[p].x += 1.0f;
            dst[i] = v[p];
        }
    }
}
int main()
{
    const int Hp = 2048, W = 8192, nyt = Hp / 256;
    const size_t n = (size_t)W * 18 * Hp;
    float2 *a, *b;
    cudaMalloc(&a, n * 4);
    cudaMalloc(&b, n * 4);
    cudaMemset(a, 0, n * 4);
    cudaMemset(b, 0, n * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int chunks : {55, 37}) {
        const int chunk = (W + chunks - 1) / chunks;
        for (int pat = 0; pat < 2; ++pat)
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                for (int it = 0; it < 20; ++it) {
                    if (pat == 0) k<0><<<nyt * chunks, 128>>>(a, b, Hp, W, nyt, chunk);
                    else k<1><<<nyt * chunks, 128>>>(a, b, Hp, W, nyt, chunk);
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                printf("chunks %d pattern %d: %.4f ms/iter, %.1f GB/s (%s)\n", chunks, pat, ms / 20, 2.0 * n * 4 / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
            }
    }
    // plain copy for reference
    cudaEventRecord(e0);
    for (int it = 0; it < 20; ++it) cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D: %.4f ms/iter, %.1f GB/s\n", ms / 20, 2.0 * n * 4 / (ms / 20 * 1e-3) / 1e9);
    return 0;
}
