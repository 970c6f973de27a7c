// lbm_kernels.cuh -- __global__ kernels of the engine (sm_100a).  See lbm_device.cuh for the layout.
#pragma once
#include "lbm_device.cuh"

namespace fdlbm {

constexpr int TPB = 128;  // threads per block along y (the contiguous axis)

// One CTA per (y tile, column); the launch grid is 1-D (y tiles fastest) so that the column count is not bound by
// the 65535 limit of gridDim.y.  Returns the cell of this thread: row y (may be >= H), column index col.
__device__ __forceinline__ void block_cell(int H, int &y, int &col)
{
    const unsigned nyt = (unsigned)(H + TPB - 1) / TPB;
    y = (int)(blockIdx.x % nyt) * TPB + threadIdx.x;
    col = (int)(blockIdx.x / nyt);
}

// ---------------------------------------------------------------------------------------------
// two-pass path, pass 1: psi_new for columns [xl_begin, xl_begin + gridDim.y)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TPB) k_psi(const __grid_constant__ LbmParams<T> P, int xl_begin)
{
    int y, xl;
    block_cell(P.H, y, xl);
    xl += xl_begin;
    if (y >= P.H) return;
    T g[9];
    const unsigned bits = P.reflect[cell_idx(P.Hp, xl, y)];
    P.psi_new[cell_idx(P.Hp, xl, y)] = stream_bc_g(P, xl, y, bits, g);
}

// ---------------------------------------------------------------------------------------------
// two-pass path, pass 2: stream + boundaries + moments (+ collide | + write-out)
//   FINALIZE = false: write post-collision populations of the next state to dst
//   FINALIZE = true : write PRE-collision populations to dst and every macroscopic field to `out`
//                     (what the reference holds after the iteration, fingering_periodic.py:470-479)
// ---------------------------------------------------------------------------------------------
template <typename T, bool FINALIZE>
__global__ void __launch_bounds__(TPB) k_step_twopass(const __grid_constant__ LbmParams<T> P, FieldPtrs<T> out)
{
    int y, xl;
    block_cell(P.H, y, xl);
    if (y >= P.H) return;
    const size_t c = cell_idx(P.Hp, xl, y);
    const unsigned bits = P.reflect[c];
    T f[9], g[9];
    const T psi = stream_bc_g(P, xl, y, bits, g);
    stream_bc_f(P, xl, y, bits, f);
    const bool solid = is_solid(P, xl, y);
    T gx, gy, lap;
    if (!solid || FINALIZE) stencil_from_array(P, P.psi_new, xl, y, gx, gy, lap);
    if (FINALIZE) {
        store_cell(P, xl, y, f, g);
        out.gx[c] = gx;
        out.gy[c] = gy;
        out.lap[c] = lap;
        if (solid) {
            out.rho[c] = out.ux[c] = out.uy[c] = out.p[c] = out.mu[c] = out.mix_tau[c] = T(0);
        } else {
            Macro<T> m;
            moments(P, f, psi, gx, gy, lap, m);
            out.rho[c] = m.rho;
            out.ux[c] = m.ux;
            out.uy[c] = m.uy;
            out.p[c] = m.p;
            out.mu[c] = m.mu;
            const T D = (T(1) - psi) + P.M * (T(1) + psi);
            out.mix_tau[c] = P.eta6m / (m.rho * D) + T(0.5);
        }
    } else {
        if (!solid) {
            Macro<T> m;
            moments(P, f, psi, gx, gy, lap, m);
            collide(P, m, f, g);
        }
        store_cell(P, xl, y, f, g);
    }
}

// ---------------------------------------------------------------------------------------------
// first collision from caller-supplied macroscopic arrays (Compute.__init__ leaves them mutually
// inconsistent: fingering_periodic.py:111, fingering.py:121-123), in place on the lattice in P.dst
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(TPB) k_collide_first(const __grid_constant__ LbmParams<T> P, FieldPtrs<T> in,
                                                       const T *psi)
{
    int y, xl;
    block_cell(P.H, y, xl);
    if (y >= P.H) return;
    // Solid cells are never collided, but what has streamed INTO them is part of the state (a restart from a
    // step-k get_state has non-zero populations there): on the two edge columns of a peer-attached slab they must
    // still reach the neighbour's ghost columns, unchanged.
    const bool solid = is_solid(P, xl, y);
    if (solid && !((P.peer_lo && xl < G) || (P.peer_hi && xl >= P.Wl - G))) return;
    const size_t c = cell_idx(P.Hp, xl, y);
    T f[9], g[9];
    const T *s = P.dst + lat_idx(P.Hp, xl, 0, y);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        f[i] = s[(size_t)i * P.Hp];
        g[i] = s[(size_t)(9 + i) * P.Hp];
    }
    if (solid) {
        store_cell(P, xl, y, f, g);
        return;
    }
    Macro<T> m;
    m.rho = in.rho[c];
    m.ux = in.ux[c];
    m.uy = in.uy[c];
    m.p = in.p[c];
    m.mu = in.mu[c];
    m.inv_mt = T(1) / in.mix_tau[c];
    m.psi = psi[c];
    m.gx = in.gx[c];
    m.gy = in.gy[c];
    collide(P, m, f, g);
    store_cell(P, xl, y, f, g);
}

// ---------------------------------------------------------------------------------------------
// layout conversion between the reference's host layout and the device layout
// ---------------------------------------------------------------------------------------------
// host plane: (H, ncw) row-major, holding global columns [col0, col0+ncw)
// device plane: base[(xl+G) * xstride + y], xl in [xl_lo, xl_hi); global column of xl is gx0 + xl,
// wrapped modulo W when `wrap` (periodic ghost columns).  Columns outside the host window are skipped.
template <typename TS, typename TD>
__global__ void k_transpose_in(const TS *__restrict__ src, int H, int ncw, int col0, TD *__restrict__ base,
                               size_t xstride, int xl_lo, int xl_hi, int gx0, int W, int wrap)
{
    __shared__ TS tile[32][33];
    const int xl_t = xl_lo + blockIdx.x * 32, y_t = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {  // coalesced read along x
        const int y = y_t + j, xl = xl_t + threadIdx.x;
        int gx = gx0 + xl;
        if (wrap) gx = ((gx % W) + W) % W;
        const int cw = gx - col0;
        if (y < H && xl < xl_hi && gx >= 0 && gx < W && cw >= 0 && cw < ncw)
            tile[j][threadIdx.x] = src[(size_t)y * ncw + cw];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {  // coalesced write along y
        const int xl = xl_t + j, y = y_t + threadIdx.x;
        int gx = gx0 + xl;
        if (wrap) gx = ((gx % W) + W) % W;
        const int cw = gx - col0;
        if (y < H && xl < xl_hi && gx >= 0 && gx < W && cw >= 0 && cw < ncw)
            base[(size_t)(xl + G) * xstride + y] = (TD)tile[threadIdx.x][j];
    }
}

template <typename TS, typename TD>
__global__ void k_transpose_out(const TS *__restrict__ base, size_t xstride, int xl_lo, int xl_hi, int gx0,
                                TD *__restrict__ dst, int H, int ncw, int col0)
{
    __shared__ TS tile[32][33];
    const int xl_t = xl_lo + blockIdx.x * 32, y_t = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int xl = xl_t + j, y = y_t + threadIdx.x;
        if (y < H && xl < xl_hi) tile[j][threadIdx.x] = base[(size_t)(xl + G) * xstride + y];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int y = y_t + j, xl = xl_t + threadIdx.x;
        const int cw = gx0 + xl - col0;
        if (y < H && xl < xl_hi && cw >= 0 && cw < ncw) dst[(size_t)y * ncw + cw] = (TD)tile[threadIdx.x][j];
    }
}

// count non-finite populations of the owned columns (watchdog standing in for np.seterr(all='raise'))
template <typename T>
__global__ void __launch_bounds__(TPB) k_count_nonfinite(const T *__restrict__ lat, int H, int Hp, unsigned long long *n_bad)
{
    int y, xl;
    block_cell(H, y, xl);
    unsigned bad = 0;
    if (y < H) {
        const T *s = lat + lat_idx(Hp, xl, 0, y);
#pragma unroll
        for (int i = 0; i < NPOP; ++i) bad += !isfinite(s[(size_t)i * Hp]);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, (unsigned long long)bad);
}

// solid bytes [(xl+G)*Hp + y] -> bitfield words [(xl+G)*(Hp/32) + y/32]; one warp ballot per word
__global__ void k_pack_solid(const uint8_t *__restrict__ bytes, uint32_t *__restrict__ bits, size_t n_cells)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, i < n_cells && bytes[i] != 0);
    if ((threadIdx.x & 31) == 0 && i < n_cells) bits[i >> 5] = b;
}

}  // namespace fdlbm
