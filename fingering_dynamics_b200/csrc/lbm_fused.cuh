// lbm_fused.cuh -- the one-pass step kernel: stream + bounce-back + Zou-He + moments + interaction
// force (psi stencils) + BGK collision, 18 populations read once and written once per lattice update.
//
// The collision of cell x needs grad/lap of the NEW psi at x, i.e. psi_new on the 8 neighbours, each a
// sum of g pulled from ITS neighbours (two-ring dependency, SURVEY.md finding 5).  Instead of staging
// psi_new through HBM (+30% traffic), a CTA owns a strip of TY consecutive y and MARCHES along x:
//
//   iteration x:  pull g of column x+1 (+ boundary rules) -> psi_new(x+1, .) into a 4-slot shared-
//                 memory ring (plus the two halo cells y0-1, y0+ny done by lanes 0/1 of warp 0);
//                 keep the pulled g in registers for the next iteration
//                 __syncthreads()
//                 pull f of column x, moments, stencils from ring rows x-1, x, x+1, collide with the
//                 g kept from the previous iteration, store the 18 populations of column x.
//
// The 4-slot ring makes one barrier per column sufficient: the slot written in iteration x+1 is not
// read in iteration x.  Work is the linearised list of (y strip, column) pairs cut into equal
// contiguous chunks, one per CTA, with exactly one resident wave (grid = SMs x occupancy), so there is
// no tail and the ring warm-up (2 extra g-column pulls, L2 hits) is paid once per chunk.
#pragma once
#include "lbm_device.cuh"

namespace fdlbm {

constexpr int FUSED_TY = 128;

// psi seen by the stencil at local column xl, row yy (yy may be -1 or H): ghost rows/columns resolved
// like psi_fetch, everything else computed from the populations.
template <typename T>
__device__ __forceinline__ T psi_new_at(const LbmParams<T> &P, int xl, int yy, T g[9], unsigned &bits)
{
    if (yy < 0 || yy >= P.H) {
        if (P.y_wall) return P.psi_wall;
        yy = yy < 0 ? yy + P.H : yy - P.H;
    }
    const int gx = P.gx0 + xl;
    if (!P.x_periodic) {
        if (gx < 0) return P.psi_left;
        if (gx >= P.W) return P.psi_right;
    }
    bits = P.reflect[cell_idx(P.Hp, xl, yy)];
    return stream_bc_g(P, xl, yy, bits, g);
}

template <typename T, int TY>
__global__ void __launch_bounds__(TY) k_fused(const __grid_constant__ LbmParams<T> P, int total_cols, int cols_per_cta)
{
    __shared__ T ring[4][TY + 2];
    const int t = threadIdx.x;
    int c = blockIdx.x * cols_per_cta;
    const int c_end = min(c + cols_per_cta, total_cols);
    while (c < c_end) {
        const int yt = c / P.Wl;
        const int xs = c - yt * P.Wl;
        const int xe = min(P.Wl, xs + (c_end - c));
        c += xe - xs;

        const int y0 = yt * TY;
        const int y = y0 + t;
        const int ny = min(TY, P.H - y0);
        const bool active = t < ny;
        const int y_halo = t == 0 ? y0 - 1 : y0 + ny;  // lanes 0 and 1 also serve the strip's halo cells
        const int s_halo = t == 0 ? 0 : ny + 1;

        T g_cur[9], g_nxt[9], psi_cur = T(0), psi_nxt = T(0);
        unsigned bits_cur = 0, bits_nxt = 0;

        // ring warm-up: columns xs-1 and xs
        {
            T gh[9];
            unsigned bh;
            T *row = ring[(xs - 1 + 4) & 3];
            if (active) row[t + 1] = psi_new_at(P, xs - 1, y, gh, bh);
            if (t < 2) row[s_halo] = psi_new_at(P, xs - 1, y_halo, gh, bh);
            row = ring[xs & 3];
            if (active) row[t + 1] = psi_cur = psi_new_at(P, xs, y, g_cur, bits_cur);
            if (t < 2) row[s_halo] = psi_new_at(P, xs, y_halo, gh, bh);
        }

        for (int x = xs; x < xe; ++x) {
            {
                T gh[9];
                unsigned bh;
                T *row = ring[(x + 1) & 3];
                if (active) row[t + 1] = psi_nxt = psi_new_at(P, x + 1, y, g_nxt, bits_nxt);
                if (t < 2) row[s_halo] = psi_new_at(P, x + 1, y_halo, gh, bh);
            }
            __syncthreads();
            if (active) {
                T f[9];
                stream_bc_f(P, x, y, bits_cur, f);
                if (!is_solid(P, x, y)) {
                    const T *rm = ring[(x - 1 + 4) & 3] + t + 1, *r0 = ring[x & 3] + t + 1, *rp = ring[(x + 1) & 3] + t + 1;
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
            bits_cur = bits_nxt;
        }
        __syncthreads();  // the ring is reused by the next run of this CTA
    }
}

// returns 0 or a cudaError_t
template <typename T>
int launch_fused(const LbmParams<T> &P, cudaStream_t stream)
{
    static int n_cta = 0;  // per instantiation; device properties do not change within a process
    if (n_cta == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused<T, FUSED_TY>, FUSED_TY, 0);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) occ = 1;
        n_cta = sms * occ;
    }
    const int nyt = (P.H + FUSED_TY - 1) / FUSED_TY;
    const int total = nyt * P.Wl;
    int cpc = (total + n_cta - 1) / n_cta;
    if (cpc < 8) cpc = 8;
    const int grid = (total + cpc - 1) / cpc;
    k_fused<T, FUSED_TY><<<grid, FUSED_TY, 0, stream>>>(P, total, cpc);
    return 0;
}

}  // namespace fdlbm
