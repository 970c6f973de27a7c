// lbm_fused.cuh -- the one-pass step kernel: stream + bounce-back + Zou-He + moments + interaction
// force (psi stencils) + BGK collision, 18 populations read once and written once per lattice update.
//
// The collision of cell x needs grad/lap of the NEW psi at x, i.e. psi_new on the 8 neighbours, each a
// sum of g pulled from ITS neighbours (two-ring dependency, SURVEY.md finding 5).  Instead of staging
// psi_new through HBM (+30% traffic), a CTA owns a strip of TY consecutive y and MARCHES along x:
//
//   * column records are streamed from HBM into shared-memory stage rings with cp.async (16-byte
//     chunks, D columns ahead of the compute front), so the loads of the next columns are in flight
//     while the current column is being collided and no registers are tied up by them;
//   * iteration x:  wait for column x+2 -> pull g of column x+1 from the stages (+ bounce-back, Zou-He)
//     -> psi_new(x+1, .) into a 4-slot psi ring (plus the strip's two halo cells), barrier,
//     pull f of column x, moments, stencils from psi ring rows x-1, x, x+1, collide with the g pulled
//     one iteration earlier (kept in registers), store the 18 populations of column x (coalesced).
//
// Work is the linearised list of (y strip, column) pairs cut into equal contiguous chunks, one per CTA,
// with exactly one resident wave (grid = SMs x CTAs/SM), so there is no tail and the pipeline warm-up
// (a few extra column loads, L2 hits) is paid once per chunk.  The y wrap is resolved when a stage is
// filled; the x wrap / slab halo comes from the two ghost columns.
#pragma once
#include "lbm_device.cuh"

namespace fdlbm {

constexpr int FUSED_TY = 128;  // rows per strip = threads per CTA
constexpr int FUSED_D = 2;     // prefetch distance in columns

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
template <int N>
__device__ __forceinline__ void cp_async_small(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem),
                 "n"(N));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, int TY>
struct FusedCfg {
    static constexpr int HALO = 16 / (int)sizeof(T);      // rows of apron per side: keeps 16-byte chunks aligned
    static constexpr int PT = TY + 2 * HALO;              // stage row pitch (elements)
    static constexpr int NS = 3 + FUSED_D;                // stages per ring (f: x-1..x+1+D, g: x..x+2+D)
    static constexpr int FAM = 9 * PT;                    // elements of one stage (one family of one column)
    static constexpr int RING = TY + 2;                   // psi ring row: y0-1 .. y0+TY
    static constexpr size_t SMEM = (size_t)(2 * NS * FAM + 4 * RING) * sizeof(T);
};

// Fill one stage: family `base` (0: f, 9: g) of local column c, rows [y0-HALO, y0+rows+HALO) with the y wrap
// applied.  16-byte cp.async for chunks that are contiguous in memory, element-wise otherwise.
template <typename T, int TY>
__device__ __forceinline__ void stage_fill(const LbmParams<T> &P, T *stage, int c, int base, int y0, int rows)
{
    using C = FusedCfg<T, TY>;
    constexpr int EPC = 16 / (int)sizeof(T);  // elements per chunk
    const int nch = (rows + 2 * C::HALO + EPC - 1) / EPC;
    const T *col = P.src + lat_idx(P.Hp, c, base, 0);
    for (int ch = threadIdx.x; ch < nch; ch += TY) {
        const int r0 = ch * EPC;
        const int y = y0 - C::HALO + r0;
        if (y >= 0 && y + EPC <= P.H) {
#pragma unroll
            for (int pop = 0; pop < 9; ++pop) cp_async16(stage + pop * C::PT + r0, col + (size_t)pop * P.Hp + y);
        } else {
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
                int yy = (y + e) % P.H;
                if (yy < 0) yy += P.H;
#pragma unroll
                for (int pop = 0; pop < 9; ++pop)
                    cp_async_small<(int)sizeof(T)>(stage + pop * C::PT + r0 + e, col + (size_t)pop * P.Hp + yy);
            }
        }
    }
}

// pull-stream + bounce-back from the stage rings: v_i <- stage_i(column c - e_x)[row j - e_y]
template <typename T, int PT>
__device__ __forceinline__ void pull_staged(const T *sm, const T *s0, const T *sp, int j, unsigned bits, T v[9])
{
    v[0] = s0[j];
    v[1] = sm[1 * PT + j];
    v[2] = s0[2 * PT + j - 1];
    v[3] = sp[3 * PT + j];
    v[4] = s0[4 * PT + j + 1];
    v[5] = sm[5 * PT + j - 1];
    v[6] = sp[6 * PT + j - 1];
    v[7] = sp[7 * PT + j + 1];
    v[8] = sm[8 * PT + j + 1];
    if (bits) {
#pragma unroll
        for (int i = 1; i < 9; ++i)
            if ((bits >> (i - 1)) & 1u) v[i] = s0[opp(i) * PT + j];
    }
}

template <typename T, int TY>
__global__ void __launch_bounds__(TY) k_fused(const __grid_constant__ LbmParams<T> P, int nyt, int chunk)
{
    using C = FusedCfg<T, TY>;
    constexpr int D = FUSED_D, NS = C::NS, PT = C::PT, HALO = C::HALO;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *fst = reinterpret_cast<T *>(smem_raw);   // [NS][9][PT]
    T *gst = fst + NS * C::FAM;                 // [NS][9][PT]
    T *ring = gst + NS * C::FAM;                // [4][RING]
    const int t = threadIdx.x;
    auto slot = [](int c) { return ((c % NS) + NS) % NS; };

    // CTA b marches over columns [xs, xe) of strip yt.  Strips are the FAST index: the nyt CTAs that work
    // on the same column range start together and advance in step, so the apron rows a strip reads
    // (HALO rows of its neighbours) are L2 hits on lines the neighbour strip is streaming anyway.
    {
        const int yt = blockIdx.x % nyt;
        const int xs = (blockIdx.x / nyt) * chunk;
        const int xe = min(P.Wl, xs + chunk);

        const int y0 = yt * TY;
        const int y = y0 + t;
        const int ny = min(TY, P.H - y0);
        const bool active = t < ny;
        const int j = t + HALO;                          // stage row of this thread's cell
        const int y_halo = t == 0 ? y0 - 1 : y0 + ny;    // lanes 0 and 1 also serve the strip's halo cells
        const int j_halo = t == 0 ? HALO - 1 : HALO + ny;
        const int s_halo = t == 0 ? 0 : ny + 1;

        // one pipeline step: g column v+2+D and f column v+1+D (only the columns this run will read)
        auto prefetch = [&](int v) {
            const int cg = v + 2 + D, cf = v + 1 + D;
            if (cg >= xs - 2 && cg <= xe + 1) stage_fill<T, TY>(P, gst + slot(cg) * C::FAM, cg, 9, y0, ny);
            if (cf >= xs - 1 && cf <= xe) stage_fill<T, TY>(P, fst + slot(cf) * C::FAM, cf, 0, y0, ny);
            cp_async_commit();
        };
        // Per-cell flags are plain global loads issued two columns ahead of their use and carried RAW in
        // registers (reflect byte, solid-mask word): nothing depends on them until they are decoded two
        // iterations later, so their latency never sits on the critical path.
        struct RawFlags {
            unsigned refl, word;
        };
        auto load_flags = [&](int c, int yy) -> RawFlags {
            RawFlags r{0u, 0u};
            if (c > xe + 1) return r;
            if (yy < 0 || yy >= P.H) {
                if (P.y_wall) return r;
                yy = yy < 0 ? yy + P.H : yy - P.H;
            }
            const int gx = P.gx0 + c;
            if (!P.x_periodic && (gx < 0 || gx >= P.W)) return r;
            r.refl = P.reflect[cell_idx(P.Hp, c, yy)];
            r.word = P.solid[(size_t)(c + G) * (P.Hp >> 5) + (yy >> 5)];
            return r;
        };
        // decode for the cell in global row yy: reflect bits | solid << 8
        auto decode = [&](RawFlags r, int yy) -> unsigned {
            if (yy < 0) yy += P.H;
            if (yy >= P.H) yy -= P.H;
            return r.refl | (((r.word >> (yy & 31)) & 1u) << 8);
        };
        // psi_new of (column c, stage row jj / global row yy) from the g stages
        auto psi_staged = [&](int c, int yy, int jj, unsigned flags, T g[9]) -> T {
            if (yy < 0 || yy >= P.H) {
                if (P.y_wall) return P.psi_wall;
                yy = yy < 0 ? yy + P.H : yy - P.H;
            }
            const int gx = P.gx0 + c;
            if (!P.x_periodic) {
                if (gx < 0) return P.psi_left;
                if (gx >= P.W) return P.psi_right;
            }
            pull_staged<T, PT>(gst + slot(c - 1) * C::FAM, gst + slot(c) * C::FAM, gst + slot(c + 1) * C::FAM, jj,
                               flags & 0xffu, g);
            if (P.zou_he && (gx == 0 || gx == P.W - 1)) zou_he_g(P, gx, yy, g);
            if (flags & 0x100u) return P.psi_wall;
            return (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) + g[8];
        };

        T g_cur[9], g_nxt[9], psi_cur = T(0), psi_nxt = T(0);
        unsigned fl_cur = 0, fl_nxt = 0;             // flags of this thread's cell in columns x, x+1
        RawFlags fq0{0, 0}, fq1{0, 0}, hq0{0, 0}, hq1{0, 0};  // look-ahead queues: own / halo cell, columns x+1, x+2

        // pipeline warm-up.  Steps v = xs-4-D .. xs-2 bring in g columns xs-2 .. xs+D (all NS slots) and
        // f columns xs-1 .. xs-1+D; psi(xs-1) needs g columns xs-2..xs, then g column xs-2 makes room.
        for (int v = xs - 4 - D; v < xs - 1; ++v) prefetch(v);
        const RawFlags z{0, 0};
        const RawFlags rf_m1 = active ? load_flags(xs - 1, y) : z, rh_m1 = t < 2 ? load_flags(xs - 1, y_halo) : z;
        const RawFlags rf_0 = active ? load_flags(xs, y) : z, rh_0 = t < 2 ? load_flags(xs, y_halo) : z;
        if (active) {
            fq0 = load_flags(xs + 1, y);
            fq1 = load_flags(xs + 2, y);
        }
        if (t < 2) {
            hq0 = load_flags(xs + 1, y_halo);
            hq1 = load_flags(xs + 2, y_halo);
        }
        cp_async_wait<D>();  // g columns <= xs have landed (this thread's share)
        __syncthreads();
        {
            T gh[9];
            T *row = ring + ((xs - 1 + 4) & 3) * C::RING;
            if (active) row[t + 1] = psi_staged(xs - 1, y, j, decode(rf_m1, y), gh);
            if (t < 2) row[s_halo] = psi_staged(xs - 1, y_halo, j_halo, decode(rh_m1, y_halo), gh);
        }
        __syncthreads();
        prefetch(xs - 1);
        cp_async_wait<D>();  // g column xs+1 has landed
        __syncthreads();
        {
            T gh[9];
            T *row = ring + (xs & 3) * C::RING;
            fl_cur = decode(rf_0, y);
            if (active) row[t + 1] = psi_cur = psi_staged(xs, y, j, fl_cur, g_cur);
            if (t < 2) row[s_halo] = psi_staged(xs, y_halo, j_halo, decode(rh_0, y_halo), gh);
        }

        for (int x = xs; x < xe; ++x) {
            cp_async_wait<D - 1>();  // g column x+2 and f column x+1 have landed
            __syncthreads();         // ... for every thread; and everybody is done with iteration x-1
            prefetch(x);             // overwrites the stages of g column x-1 and f column x-2: no longer read
            const RawFlags fq2 = active ? load_flags(x + 3, y) : z;       // decoded two iterations from now
            const RawFlags hq2 = t < 2 ? load_flags(x + 3, y_halo) : z;
            {
                T gh[9];
                T *row = ring + ((x + 1) & 3) * C::RING;
                fl_nxt = decode(fq0, y);
                if (active) row[t + 1] = psi_nxt = psi_staged(x + 1, y, j, fl_nxt, g_nxt);
                if (t < 2) row[s_halo] = psi_staged(x + 1, y_halo, j_halo, decode(hq0, y_halo), gh);
            }
            __syncthreads();
            if (active) {
                T f[9];
                pull_staged<T, PT>(fst + slot(x - 1) * C::FAM, fst + slot(x) * C::FAM, fst + slot(x + 1) * C::FAM, j,
                                   fl_cur & 0xffu, f);
                if (P.zou_he) {
                    const int gx_ = P.gx0 + x;
                    if (gx_ == 0 || gx_ == P.W - 1) zou_he_f(P, x, gx_, y, f, PullRow<T>());
                }
                if (!(fl_cur & 0x100u)) {
                    const T *rm = ring + ((x - 1 + 4) & 3) * C::RING + t + 1, *r0 = ring + (x & 3) * C::RING + t + 1,
                            *rp = ring + ((x + 1) & 3) * C::RING + t + 1;
                    T gx, gy, lap;
                    //        C      E      W      N      S      NE     NW     SW      SE
                    stencil9(psi_cur, rp[0], rm[0], r0[1], r0[-1], rp[1], rm[1], rm[-1], rp[-1], gx, gy, lap);
                    Macro<T> m;
                    moments(P, f, psi_cur, gx, gy, lap, m);
                    collide(P, m, f, g_cur);
                }
                store_cell(P, x, y, f, g_cur);
                if (P.zou_he) {  // the next step's Zou-He needs grad psi and mu at the face columns
                    const int gx_ = P.gx0 + x;
                    if (gx_ < 2 || gx_ >= P.W - 2) P.psi_new[cell_idx(P.Hp, x, y)] = psi_cur;
                }
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) g_cur[i] = g_nxt[i];
            psi_cur = psi_nxt;
            fl_cur = fl_nxt;
            fq0 = fq1;
            fq1 = fq2;
            hq0 = hq1;
            hq1 = hq2;
        }
        cp_async_wait<0>();
    }
}

// returns 0 or a cudaError_t
template <typename T>
int launch_fused(const LbmParams<T> &P, cudaStream_t stream)
{
    using C = FusedCfg<T, FUSED_TY>;
    static int n_cta = 0;  // per instantiation; device properties do not change within a process
    if (n_cta == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k_fused<T, FUSED_TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused<T, FUSED_TY>, FUSED_TY, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) occ = 1;
        n_cta = sms * occ;
    }
    // one resident wave: (CTA slots / strips) column chunks per strip, all strips in step
    const int nyt = (P.H + FUSED_TY - 1) / FUSED_TY;
    int nchunks = n_cta / nyt;
    if (nchunks < 1) nchunks = 1;
    int chunk = (P.Wl + nchunks - 1) / nchunks;
    if (chunk < 8) chunk = 8;
    nchunks = (P.Wl + chunk - 1) / chunk;
    k_fused<T, FUSED_TY><<<nyt * nchunks, FUSED_TY, C::SMEM, stream>>>(P, nyt, chunk);
    return 0;
}

}  // namespace fdlbm
