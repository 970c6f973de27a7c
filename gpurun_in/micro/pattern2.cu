// DRAM access-pattern probe 2: which ingredient of the step kernel costs copy bandwidth?
#include <cstdio>
#include <cuda_runtime.h>
// MODE bits: 1 = __syncthreads per column, 2 = compute delay between loads and stores, 4 = rows shifted by one element
// (misaligned, scalar loads), 8 = L2 prefetch two columns ahead, 16 = half of the populations through cp.async + smem
template <int MODE>
__global__ void __launch_bounds__(128, 3) k(const float *__restrict__ src, float *__restrict__ dst, int Hp, int W, int nyt, int chunk, int spin)
{
    __shared__ __align__(16) float st[2][9][264];
    const int yt = blockIdx.x % nyt, xs = (blockIdx.x / nyt) * chunk, xe = min(W, xs + chunk);
    const int t = threadIdx.x;
    for (int x = xs; x < xe; ++x) {
        float2 v[18];
        const float *col = src + ((size_t)x * 18) * Hp + yt * 256 + 2 * t;
        if (MODE & 16) {
            if (t < 64)
                for (int p = 0; p < 9; ++p)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&st[x & 1][p][4 * t])), "l"(src + ((size_t)x * 18 + 9 + p) * Hp + yt * 256 + 4 * t));
            asm volatile("cp.async.commit_group;");
        }
        if (MODE & 8) {
            const int xp = min(x + 2, W - 1);
            if (t < 72) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ((size_t)xp * 18 + t / 8) * Hp + yt * 256 + (t % 8) * 32));
        }
#pragma unroll
        for (int p = 0; p < 18; ++p) {
            if ((MODE & 16) && p >= 9) continue;
            const float *a = col + (size_t)p * Hp;
            if ((MODE & 4) && (p % 3) != 0) {
                const int sh = (p % 3 == 1) ? -1 : 1;
                const int o = (yt * 256 + 2 * t + sh < 0 || yt * 256 + 2 * t + 1 + sh >= Hp) ? 0 : sh;
                v[p] = make_float2(a[o], a[o + 1]);
            } else
                v[p] = *reinterpret_cast<const float2 *>(a);
        }
        if (MODE & 16) {
            asm volatile("cp.async.wait_group 0;");
            __syncthreads();
#pragma unroll
            for (int p = 9; p < 18; ++p) v[p] = *reinterpret_cast<const float2 *>(&st[x & 1][p - 9][2 * t]);
        }
        if (MODE & 1) __syncthreads();
        if (MODE & 2) {
            long long t0 = clock64();
            while (clock64() - t0 < spin) {}
        }
        float *d = dst + ((size_t)x * 18) * Hp + yt * 256 + 2 * t;
#pragma unroll
        for (int p = 0; p < 18; ++p) {
            v[p].x += 1.0f;
            *reinterpret_cast<float2 *>(d + (size_t)p * Hp) = v[p];
        }
    }
}
template <int MODE>
void run(const float *a, float *b, int Hp, int W, int spin, const char *name)
{
    const int nyt = Hp / 256, chunks = 55, chunk = (W + chunks - 1) / chunks;
    const size_t n = (size_t)W * 18 * Hp;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int it = 0; it < 20; ++it) k<MODE><<<nyt * chunks, 128>>>(a, b, Hp, W, nyt, chunk, spin);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %.4f ms, %.1f GB/s (%s)\n", name, ms / 20, 2.0 * n * 4 / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
}
int main()
{
    const int Hp = 2048, W = 8192;
    const size_t n = (size_t)W * 18 * Hp;
    float *a, *b;
    cudaMalloc(&a, n * 4);
    cudaMalloc(&b, n * 4);
    cudaMemset(a, 0, n * 4);
    cudaMemset(b, 0, n * 4);
    run<0>(a, b, Hp, W, 0, "copy");
    run<1>(a, b, Hp, W, 0, "+barrier");
    run<3>(a, b, Hp, W, 1500, "+barrier +delay 1500 clk");
    run<3>(a, b, Hp, W, 3000, "+barrier +delay 3000 clk");
    run<4>(a, b, Hp, W, 0, "misaligned rows");
    run<8>(a, b, Hp, W, 0, "L2 prefetch 2 ahead");
    run<16>(a, b, Hp, W, 0, "g through cp.async");
    run<31>(a, b, Hp, W, 1500, "all, delay 1500");
    run<31>(a, b, Hp, W, 3000, "all, delay 3000");
    run<23>(a, b, Hp, W, 3000, "all but L2 prefetch, delay 3000");
    return 0;
}
