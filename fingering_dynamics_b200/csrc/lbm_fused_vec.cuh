// lbm_fused_vec.cuh -- the fused step kernel with VEC consecutive rows per thread (VEC = 2 for fp32).
//
// Same algorithm as k_fused (lbm_fused.cuh: march along x, g columns through a cp.async stage ring, f through
// registers with L2 prefetch, psi exchanged by warp shuffles, one barrier per column); a thread owns rows
// y0 + VEC*t .. y0 + VEC*t + VEC-1.  In fp32 the memory time per column is half of fp64's, so the kernel is
// bound by instruction issue; two rows per thread share the address arithmetic, flag decoding, shuffles,
// edge-lane work and loop overhead, and move 8 bytes per load/store instruction where rows are aligned pairs.
#pragma once
#include <type_traits>

#include <cstdlib>
#include <cstring>

#include "lbm_fused.cuh"

#ifndef FDLBM_FUSED_MINB32V
#define FDLBM_FUSED_MINB32V 3  // fp32, two rows per thread: the register budget of fp64 with one row
#endif
#ifndef FDLBM_F32_VEC
#define FDLBM_F32_VEC 2
#endif
#ifndef FDLBM_BULK_COPY_VEC
#define FDLBM_BULK_COPY_VEC 0
#endif

namespace fdlbm {

template <typename T, int NT, int VEC>
struct VecCfg {
    static constexpr int ROWS = NT * VEC;                 // rows per strip
    static constexpr int HALO = 16 / (int)sizeof(T);
    static constexpr int PT = ROWS + 2 * HALO;
    static constexpr int NS = 3 + FUSED_D;
    static constexpr int FAM = 9 * PT;
    static constexpr size_t SMEM = (size_t)(NS * FAM) * sizeof(T);
};

// columns [xs, xe) of strip yt (the body of k_fused_vec; also run by the face CTAs of k_fused_f32p)
template <typename T, int NT, int VEC>
__device__ __forceinline__ void fused_vec_strip(const LbmParams<T> &P, const int yt, const int xs, const int xe)
{
    using C = VecCfg<T, NT, VEC>;
    constexpr int D = FUSED_D, NS = C::NS, PT = C::PT, HALO = C::HALO, FAM = C::FAM, ROWS = C::ROWS;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert((NS & (NS - 1)) == 0, "stage ring must be a power of two");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *gst = reinterpret_cast<T *>(smem_raw);  // [NS][9][PT]
    __shared__ __align__(8) unsigned long long bars[NS];  // one mbarrier per g stage (bulk-copy path)
    const int t = threadIdx.x, lane = t & 31;
    const int H = P.H, Hp = P.Hp;

    const int y0 = yt * ROWS;
    const int yb = y0 + VEC * t;                 // first row of this thread
    const int ny = min(ROWS, H - y0);            // rows of this strip
    const int nv = max(0, min(VEC, ny - VEC * t));  // active rows of this thread
    const int jb = HALO + VEC * t;               // stage row of the thread's first cell

    // outer neighbours of the thread's rows: row yb-1 and row yb+nv.  Inside a warp they are the last row of
    // lane-1 and the first row of lane+1 (shuffles); the first and the last active lane evaluate theirs themselves.
    const bool has = nv > 0;
    const bool edge_lo = has && lane == 0;
    const bool edge_hi = has && (lane == 31 || VEC * (t + 1) >= ny);
    auto wrap_row = [&](int yy) {  // wrapped global row, or -1 for a ghost row of a y-wall variant
        if (yy < 0 || yy >= H) return P.y_wall ? -1 : (yy < 0 ? yy + H : yy - H);
        return yy;
    };
    const int ye_lo = wrap_row(yb - 1), ye_hi = wrap_row(yb + nv);

    auto slot = [](int c) { return c & (NS - 1); };
    auto in_domain = [&](int c) { return P.x_periodic || (P.gx0 + c >= 0 && P.gx0 + c < P.W); };

    // interior strips: bulk asynchronous copies completing on the stage's mbarrier (see k_fused)
    constexpr unsigned ROW_BYTES = (unsigned)(PT * sizeof(T));
    // (measured for fp32 / two rows per thread: 26.7 GLUPS with bulk copies vs 30.2 with per-thread cp.async: off)
    const bool bulk = FDLBM_BULK_COPY_VEC && y0 - HALO >= 0 && y0 + ROWS + HALO <= H;  // CTA-uniform
    if (t == 0) {
#pragma unroll
        for (int s_ = 0; s_ < NS; ++s_) mbar_init(&bars[s_], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // one pipeline step: g column v+2+D, rows [y0-HALO, y0+ny+HALO) with the y wrap applied
    auto prefetch = [&](int v) {
        const int cg = v + 2 + D;
        if (cg >= xs - 2 && cg <= xe + 1) {
            T *stage = gst + slot(cg) * FAM;
            const T *col = P.src + lat_idx(Hp, cg, 9, 0);
            if (bulk) {
                if (t == 0) {
                    mbar_expect_tx(&bars[slot(cg)], 9u * ROW_BYTES);
#pragma unroll
                    for (int pop = 0; pop < 9; ++pop)
                        bulk_g2s(stage + pop * PT, col + (size_t)pop * Hp + (y0 - HALO), ROW_BYTES, &bars[slot(cg)]);
                }
            } else {
                stage_fill<T, NT, PT, HALO>(stage, col, Hp, H, y0, ny);
            }
        }
        cp_async_commit();
    };
    auto landed = [&](int c) {
        if (bulk) mbar_wait(&bars[slot(c)], (unsigned)(((c - (xs - 2)) / NS) & 1));
    };
    // raw flags of the thread's rows (reflect bytes packed little-endian, solid-mask word) and of one edge row
    auto load_flags = [&](int c, int yy, int n) -> RawFlags {
        RawFlags r{0u, 0u};
        if (c > xe + 1 || yy < 0 || n <= 0 || !in_domain(c)) return r;
        const uint8_t *p = P.reflect + cell_idx(Hp, c, yy);
        if (VEC == 2 && n == 2)
            r.refl = *reinterpret_cast<const unsigned short *>(p);  // yy is even: 2-byte aligned
        else
            r.refl = p[0];
        r.word = P.solid[(size_t)(c + G) * (Hp >> 5) + (yy >> 5)];
        return r;
    };
    auto decode = [](RawFlags r, int yy, int v) -> unsigned {  // reflect bits | solid << 8 of row yy + v
        return ((r.refl >> (8 * v)) & 0xffu) | (((r.word >> ((yy + v) & 31)) & 1u) << 8);
    };
    // psi_new of the cell (column c, wrapped global row yy >= 0, stage row jj) from the g stages
    auto psi_staged = [&](int c, int yy, int jj, unsigned flags, T g[9]) -> T {
        if (yy < 0) return P.psi_wall;
        const int gx = P.gx0 + c;
        if (!P.x_periodic) {
            if (gx < 0) return P.psi_left;
            if (gx >= P.W) return P.psi_right;
        }
        pull_staged<T, PT>(gst + slot(c - 1) * FAM, gst + slot(c) * FAM, gst + slot(c + 1) * FAM, jj, flags & 0xffu, g);
        if (P.zou_he && (gx == 0 || gx == P.W - 1)) zou_he_g(P, gx, yy, g);
        if (flags & 0x100u) return P.psi_wall;
        return (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) + g[8];
    };
    // psi_new of column c on the thread's rows (q[0..VEC-1]) and its outer neighbours (q_lo, q_hi); pulled g returned
    // (flags arrive DECODED: fl[] own rows, fl_lo / fl_hi the outer neighbours)
    auto psi_column = [&](int c, unsigned fl_lo, unsigned fl_hi, const unsigned fl[VEC], T g[VEC][9], T q[VEC], T &q_lo,
                          T &q_hi) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            q[v] = T(0);
            if (v < nv) q[v] = psi_staged(c, yb + v, jb + v, fl[v], g[v]);
        }
        T e_lo = T(0), e_hi = T(0);
        if (edge_lo || edge_hi) {
            T gh[9];
            if (edge_lo) e_lo = psi_staged(c, ye_lo, jb - 1, fl_lo, gh);
            if (edge_hi) e_hi = psi_staged(c, ye_hi, jb + nv, fl_hi, gh);
        }
        const T dn = __shfl_up_sync(FULL, q[VEC - 1], 1), up = __shfl_down_sync(FULL, q[0], 1);
        q_lo = edge_lo ? e_lo : dn;
        q_hi = edge_hi ? e_hi : up;
        if (VEC == 2 && nv == 1) q[VEC - 1] = q_hi;  // strip tail: the row above the last active row
    };

    T g_cur[VEC][9], g_nxt[VEC][9];
    T pm[VEC], p0[VEC], pp[VEC], pm_lo, pm_hi, p0_lo, p0_hi, pp_lo, pp_hi;  // psi_new: columns x-1, x, x+1
    unsigned fl_cur[VEC], fl_nxt[VEC];
    const RawFlags z{0u, 0u};
    RawFlags fq0 = z, fq1 = z, lq0 = z, lq1 = z, hq0 = z, hq1 = z;  // look-ahead queues: own rows / low edge / high edge

    for (int v = xs - 4 - D; v < xs - 1; ++v) prefetch(v);
    const RawFlags rf_m1 = load_flags(xs - 1, yb, nv), rl_m1 = edge_lo ? load_flags(xs - 1, ye_lo, 1) : z,
                   rh_m1 = edge_hi ? load_flags(xs - 1, ye_hi, 1) : z;
    const RawFlags rf_0 = load_flags(xs, yb, nv), rl_0 = edge_lo ? load_flags(xs, ye_lo, 1) : z,
                   rh_0 = edge_hi ? load_flags(xs, ye_hi, 1) : z;
    fq0 = load_flags(xs + 1, yb, nv);
    fq1 = load_flags(xs + 2, yb, nv);
    if (edge_lo) {
        lq0 = load_flags(xs + 1, ye_lo, 1);
        lq1 = load_flags(xs + 2, ye_lo, 1);
    }
    if (edge_hi) {
        hq0 = load_flags(xs + 1, ye_hi, 1);
        hq1 = load_flags(xs + 2, ye_hi, 1);
    }
    cp_async_wait<D>();
    landed(xs - 2);
    landed(xs - 1);
    landed(xs);
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) fl_nxt[v] = decode(rf_m1, yb, v);
    psi_column(xs - 1, decode(rl_m1, ye_lo, 0), decode(rh_m1, ye_hi, 0), fl_nxt, g_nxt, pm, pm_lo, pm_hi);
    __syncthreads();
    prefetch(xs - 1);
    cp_async_wait<D>();
    landed(xs + 1);
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) fl_cur[v] = decode(rf_0, yb, v);
    psi_column(xs, decode(rl_0, ye_lo, 0), decode(rh_0, ye_hi, 0), fl_cur, g_cur, p0, p0_lo, p0_hi);

    for (int x = xs; x < xe; ++x) {
        cp_async_wait<D - 1>();  // g column x+2 has landed
        landed(x + 2);
        __syncthreads();         // ... for every thread; and everybody is done with iteration x-1
        // decode the flags of column x+1 before any new global load is issued (see k_fused)
        unsigned fl_lo_n = decode(lq0, ye_lo, 0), fl_hi_n = decode(hq0, ye_hi, 0);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            fl_nxt[v] = decode(fq0, yb, v);
            asm volatile("" : "+r"(fl_nxt[v]));
        }
        asm volatile("" : "+r"(fl_lo_n), "+r"(fl_hi_n)::"memory");
        prefetch(x);
        T f[VEC][9];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (v < nv) pull(P, x, yb + v, 0, fl_cur[v] & 0xffu, f[v]);
        {
            constexpr int LPP = (ROWS * (int)sizeof(T) + 127) / 128;  // 128-byte lines per population row of the strip
            const int cf = x + FUSED_L2_AHEAD;
            const int tt = NT - 1 - t;
            if (tt < 9 * LPP && cf <= xe + 1) {
                const int pop = tt / LPP, ln = tt - pop * LPP;
                const int yy = y0 + ln * (128 / (int)sizeof(T));
                if (yy < H) prefetch_l2(P.src + lat_idx(Hp, cf, pop, yy));
            }
        }
        const RawFlags fq2 = load_flags(x + 3, yb, nv);
        const RawFlags lq2 = edge_lo ? load_flags(x + 3, ye_lo, 1) : z;
        const RawFlags hq2 = edge_hi ? load_flags(x + 3, ye_hi, 1) : z;
        psi_column(x + 1, fl_lo_n, fl_hi_n, fl_nxt, g_nxt, pp, pp_lo, pp_hi);

        const int gx_ = P.gx0 + x;
        const bool face = P.zou_he && (gx_ == 0 || gx_ == P.W - 1);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (v < nv) {
                if (face) zou_he_f(P, x, gx_, yb + v, f[v], PullRow<T>());
                if (!(fl_cur[v] & 0x100u)) {
                    // 3x3 psi neighbourhood of row v: rows v-1 / v+1 are own rows or the outer neighbours
                    const T S0 = v > 0 ? p0[v > 0 ? v - 1 : 0] : p0_lo, N0 = v < VEC - 1 ? p0[v < VEC - 1 ? v + 1 : 0] : p0_hi;
                    const T Sm = v > 0 ? pm[v > 0 ? v - 1 : 0] : pm_lo, Nm = v < VEC - 1 ? pm[v < VEC - 1 ? v + 1 : 0] : pm_hi;
                    const T Sp = v > 0 ? pp[v > 0 ? v - 1 : 0] : pp_lo, Np = v < VEC - 1 ? pp[v < VEC - 1 ? v + 1 : 0] : pp_hi;
                    T gx, gy, lap;
                    //        C      E      W      N   S   NE  NW  SW  SE
                    stencil9(p0[v], pp[v], pm[v], N0, S0, Np, Nm, Sm, Sp, gx, gy, lap);
                    Macro<T> m;
                    moments(P, f[v], p0[v], gx, gy, lap, m);
                    collide(P, m, f[v], g_cur[v]);
                }
            }
        }
        // store: aligned row pairs go out as one 8/16-byte store per population
        if (VEC == 2 && nv == 2) {
            typedef typename std::conditional<sizeof(T) == 4, float2, double2>::type T2;
            auto put2 = [&](T *lat, int xl) {
                T *o = lat + lat_idx(Hp, xl, 0, yb);
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    T2 a, b;
                    a.x = f[0][i], a.y = f[VEC - 1][i], b.x = g_cur[0][i], b.y = g_cur[VEC - 1][i];
                    *reinterpret_cast<T2 *>(o + (size_t)i * Hp) = a;
                    *reinterpret_cast<T2 *>(o + (size_t)(9 + i) * Hp) = b;
                }
            };
            put2(P.dst, x);
            if (P.peer_lo && x < G) put2(P.peer_lo, P.peer_lo_Wl + x);
            if (P.peer_hi && x >= P.Wl - G) put2(P.peer_hi, x - P.Wl);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (v < nv) store_cell(P, x, yb + v, f[v], g_cur[v]);
        }
        if (P.zou_he && (gx_ < 2 || gx_ >= P.W - 2)) {  // the next step's Zou-He needs grad psi and mu at the faces
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (v < nv) P.psi_new[cell_idx(Hp, x, yb + v)] = p0[v];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
#pragma unroll
            for (int i = 0; i < 9; ++i) g_cur[v][i] = g_nxt[v][i];
            pm[v] = p0[v], p0[v] = pp[v];
            fl_cur[v] = fl_nxt[v];
        }
        pm_lo = p0_lo, pm_hi = p0_hi, p0_lo = pp_lo, p0_hi = pp_hi;
        fq0 = fq1, fq1 = fq2;
        lq0 = lq1, lq1 = lq2;
        hq0 = hq1, hq1 = hq2;
    }
    cp_async_wait<0>();
}

template <typename T, int NT, int VEC>
__global__ void __launch_bounds__(NT, sizeof(T) == 8 ? FDLBM_FUSED_MINB64 : FDLBM_FUSED_MINB32V)
    k_fused_vec(const __grid_constant__ LbmParams<T> P, int nyt, int chunk)
{
    const int xs = (blockIdx.x / nyt) * chunk;
    fused_vec_strip<T, NT, VEC>(P, blockIdx.x % nyt, xs, min(P.Wl, xs + chunk));
}

template <typename T, int NT, int VEC>
int launch_fused_vec(const LbmParams<T> &P, cudaStream_t stream)
{
    using C = VecCfg<T, NT, VEC>;
    auto kern = k_fused_vec<T, NT, VEC>;
    // resident CTA slots, cached per device (the shared-memory attribute is a per-device setting too)
    static int n_cta_of[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaErrorInvalidDevice;
    int &n_cta = n_cta_of[dev & 63];
    if (n_cta == 0) {
        int sms = 0, occ = 0;
        cudaError_t e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) occ = 1;
        n_cta = sms * occ;
    }
    const int nyt = (P.H + C::ROWS - 1) / C::ROWS;
    const int chunk = fused_chunk(nyt, n_cta, P.Wl);
    const int nchunks = (P.Wl + chunk - 1) / chunk;
    kern<<<nyt * nchunks, NT, C::SMEM, stream>>>(P, nyt, chunk);
    return 0;
}

}  // namespace fdlbm

#include "lbm_fused_f32.cuh"

namespace fdlbm {

// the step kernel used for each storage type
template <typename T>
int launch_fused_auto(const LbmParams<T> &P, cudaStream_t stream);
template <>
inline int launch_fused_auto<double>(const LbmParams<double> &P, cudaStream_t stream)
{
#ifdef FDLBM_F64_VEC1
    return launch_fused_vec<double, FUSED_TY, 1>(P, stream);
#else
    return launch_fused<double>(P, stream);
#endif
}
template <>
inline int launch_fused_auto<float>(const LbmParams<float> &P, cudaStream_t stream)
{
#if FDLBM_F32_VEC == 2
    // even heights: packed two-row kernel (lbm_fused_f32.cuh); FDLBM_F32_KERNEL=vec selects the scalar two-row kernel
    static const bool force_vec = [] {
        const char *v = getenv("FDLBM_F32_KERNEL");
        return v && strcmp(v, "vec") == 0;
    }();
    if (!force_vec && f32p::applicable(P)) return f32p::launch(P, stream);
    return launch_fused_vec<float, FUSED_TY, 2>(P, stream);
#else
    return launch_fused<float>(P, stream, B);
#endif
}

}  // namespace fdlbm
