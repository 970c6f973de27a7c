// Skeleton of the fused fp32 step kernel: the same memory operations per column (g through a cp.async stage ring, f
// through registers, optional L2 prefetch, 18 x 8-byte stores), the arithmetic replaced by timed delays.  Which
// ordering / prefetch policy gives the memory system the best request stream?
#include <cstdio>
#include <cuda_runtime.h>
#define PT 264
__device__ __forceinline__ void spin(int clk)
{
    long long t0 = clock64();
    while (clock64() - t0 < clk) {}
}
// F_AHEAD: 0 = f loaded at the top of its own iteration (consumed after the psi delay), 1 = one iteration ahead
// L2A: L2 prefetch distance of f (0 = off); NS: g stages; GD: cp.async distance of g (columns ahead of x+2)
template <int F_AHEAD, int L2A, int NS, int GD, int G_REG>
__global__ void __launch_bounds__(128, 3) k(const float *__restrict__ src, float *__restrict__ dst, int Hp, int W, int nyt, int chunk, int d1, int d2)
{
    extern __shared__ __align__(16) float st[];  // [NS][9][PT]
    const int yt = blockIdx.x % nyt, xs = (blockIdx.x / nyt) * chunk, xe = min(W - 4 - GD, xs + chunk);
    const int t = threadIdx.x;
    const size_t S = (size_t)18 * Hp;
    auto fill = [&](int c) {  // g column c into its stage
        if (!G_REG) {
            if (t < 66 && c >= 0 && c < W) {
                const int y = yt * 256 - 4 + 4 * t;
                const int yy = y < 0 ? 0 : (y + 4 > Hp ? Hp - 4 : y);
#pragma unroll
                for (int p = 0; p < 9; ++p)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&st[((c % NS) * 9 + p) * PT + 4 * t])), "l"(src + (size_t)c * S + (size_t)(9 + p) * Hp + yy));
            }
            asm volatile("cp.async.commit_group;");
        }
    };
    auto load_f = [&](int c, float2 v[9]) {
        const float *col = src + (size_t)c * S + yt * 256 + 2 * t;
#pragma unroll
        for (int p = 0; p < 9; ++p) v[p] = *reinterpret_cast<const float2 *>(col + (size_t)p * Hp);
    };
    for (int c = xs; c < xs + 2 + GD; ++c) fill(c);
    float2 f[9], fn[9], g[9], gn[9];
    if (F_AHEAD) load_f(xs, f);
    if (G_REG) load_f(xs, gn);
    for (int x = xs; x < xe; ++x) {
        if (!G_REG) {
            if (GD == 1) asm volatile("cp.async.wait_group 0;");
            if (GD == 2) asm volatile("cp.async.wait_group 1;");
            if (GD == 3) asm volatile("cp.async.wait_group 2;");
        }
        __syncthreads();
        fill(x + 2 + GD);
        if (G_REG) {  // g by plain loads one column ahead, shared through smem by the thread itself
#pragma unroll
            for (int p = 0; p < 9; ++p) *reinterpret_cast<float2 *>(&st[(((x + 2) % NS) * 9 + p) * PT + 4 + 2 * t]) = gn[p];
            const float *col = src + (size_t)(x + 3) * S + yt * 256 + 2 * t;
#pragma unroll
            for (int p = 0; p < 9; ++p) gn[p] = *reinterpret_cast<const float2 *>(col + (size_t)(9 + p) * Hp);
        }
        if (F_AHEAD) load_f(x + 1, fn);
        else load_f(x, f);
        if (L2A > 0 && t >= 128 - 72) {
            const int tt = 127 - t, xp = min(x + F_AHEAD + L2A, W - 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)xp * S + (size_t)(tt / 8) * Hp + yt * 256 + (tt % 8) * 32));
        }
        // psi phase: read the stages of columns x .. x+2
#pragma unroll
        for (int p = 0; p < 9; ++p) g[p] = *reinterpret_cast<const float2 *>(&st[(((x + (p % 3)) % NS) * 9 + p) * PT + 4 + 2 * t]);
        spin(d1);
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < 9; ++p) acc += f[p].x + g[p].y;   // first use of f
        spin(d2);
        float *d = dst + (size_t)x * S + yt * 256 + 2 * t;
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            *reinterpret_cast<float2 *>(d + (size_t)p * Hp) = make_float2(f[p].x + acc, f[p].y);
            *reinterpret_cast<float2 *>(d + (size_t)(9 + p) * Hp) = make_float2(g[p].x, g[p].y + acc);
        }
        if (F_AHEAD) {
#pragma unroll
            for (int p = 0; p < 9; ++p) f[p] = fn[p];
        }
    }
    asm volatile("cp.async.wait_group 0;");
}
static int g_min_smem = 0;
template <int F_AHEAD, int L2A, int NS, int GD, int G_REG>
void run(const float *a, float *b, int Hp, int W, int d1, int d2, const char *name)
{
    const int nyt = Hp / 256, chunks = 55, chunk = (W + chunks - 1) / chunks;
    const size_t n = (size_t)W * 18 * Hp;
    auto kern = k<F_AHEAD, L2A, NS, GD, G_REG>;
    const int smem = max(NS * 9 * PT * 4, g_min_smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int it = 0; it < 20; ++it) kern<<<nyt * chunks, 128, smem>>>(a, b, Hp, W, nyt, chunk, d1, d2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-60s d=%d+%d occ=%d %.4f ms, %.1f GB/s (%s)\n", name, d1, d2, occ, ms / 20, 2.0 * n * 4 / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
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
    g_min_smem = 75 * 1024;  // 3 CTAs per SM, like the real kernels
    for (int d2 : {1200, 2400, 3600, 4800}) {
        const int d1 = d2 / 3;
        run<0, 2, 4, 1, 0>(a, b, Hp, W, d1, d2, "f at top, L2 prefetch 2, 4 stages (kernel as is)");
        run<0, 0, 4, 1, 0>(a, b, Hp, W, d1, d2, "f at top, no L2 prefetch");
        run<1, 0, 4, 1, 0>(a, b, Hp, W, d1, d2, "f one iteration ahead, no L2 prefetch");
        run<1, 2, 4, 1, 0>(a, b, Hp, W, d1, d2, "f one iteration ahead, L2 prefetch 2");
    }
    return 0;
}
