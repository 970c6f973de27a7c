// lbm_init.cuh -- Compute.__init__ on the device: the initial state of
// fingering_periodic.py:90-121 and fingering.py:95-127 is analytic in (column, solid flag), so a large grid
// needs no host arrays at all: 2 bytes of geometry per cell go up, nothing else.
//
// The arithmetic is done in double in the REFERENCE'S operation order with the rounding of every operation
// pinned (__dmul_rn / __dadd_rn / __ddiv_rn are never contracted into FMAs): for an fp64 engine the populations
// and macroscopic arrays are bit-identical to what the NumPy code builds; an fp32 engine stores them rounded,
// exactly like fdlbm_set_state does with the caller's float64 arrays.
#pragma once
#include "lbm_device.cuh"
#include "lbm_kernels.cuh"

namespace fdlbm {

struct InitParams {
    int variant;               // FDLBM_INIT_FP = 1, FDLBM_INIT_FG = 2
    int n_inject;              // global columns [0, n_inject) hold psi_inject
    int have_rho;              // 1: rho comes from the rho plane of the field block (uploaded by the caller)
    double psi_inject, psi_rest, rho0;
    double gamma, a, kappa, Eta_n, M, psi_wall, psi_left, psi_right;
};

namespace ex {  // exactly rounded, never fused
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
}  // namespace ex

// psi of Compute.__init__ as the stencils see it (fingering_periodic.py:90-95,216-218; fingering.py:221-224)
template <typename T>
__device__ __forceinline__ double init_psi(const LbmParams<T> &P, const InitParams &I, int xl, int y)
{
    if (y < 0 || y >= P.H) {
        if (P.y_wall) return I.psi_wall;
        y = y < 0 ? y + P.H : y - P.H;
    }
    int gx = P.gx0 + xl;
    if (P.x_periodic) {
        gx = gx < 0 ? gx + P.W : (gx >= P.W ? gx - P.W : gx);
    } else {
        if (gx < 0) return I.psi_left;
        if (gx >= P.W) return I.psi_right;
    }
    if (is_solid(P, xl, y)) return I.psi_wall;
    return gx < I.n_inject ? I.psi_inject : I.psi_rest;
}

// psi plane of all local columns including the ghosts (the first step's Zou-He and k_collide_first read it)
template <typename T>
__global__ void __launch_bounds__(TPB) k_init_psi(const __grid_constant__ LbmParams<T> P, const InitParams I, T *psi)
{
    int y, col;
    block_cell(P.H, y, col);
    const int xl = col - G;
    if (y >= P.H) return;
    const int gx = P.gx0 + xl;
    if (!P.x_periodic && (gx < 0 || gx >= P.W)) return;
    psi[cell_idx(P.Hp, xl, y)] = (T)init_psi(P, I, xl, y);
}

// macroscopic arrays + equilibrium populations of the owned columns
template <typename T>
__global__ void __launch_bounds__(TPB) k_init_cells(const __grid_constant__ LbmParams<T> P, const InitParams I, FieldPtrs<T> out)
{
    using namespace ex;
    int y, xl;
    block_cell(P.H, y, xl);
    if (y >= P.H) return;
    const size_t c = cell_idx(P.Hp, xl, y);
    // stencils of the initial psi in the reference's accumulation order (fingering_periodic.py:214-256)
    const double C = init_psi(P, I, xl, y);
    const double E = init_psi(P, I, xl + 1, y), Wv = init_psi(P, I, xl - 1, y);
    const double N = init_psi(P, I, xl, y + 1), S = init_psi(P, I, xl, y - 1);
    const double NE = init_psi(P, I, xl + 1, y + 1), NW = init_psi(P, I, xl - 1, y + 1);
    const double SW = init_psi(P, I, xl - 1, y - 1), SE = init_psi(P, I, xl + 1, y - 1);
    const double gx = div(add(add(add(add(add(add(0.0, mul(4, E)), mul(-4, Wv)), NE), -NW), -SW), SE), 12);
    const double gy = div(add(add(add(add(add(add(0.0, mul(4, N)), mul(-4, S)), NE), NW), -SW), -SE), 12);
    const double lap =
        div(add(add(add(add(add(add(add(add(add(0.0, mul(-20, C)), mul(4, N)), mul(4, E)), mul(4, Wv)), mul(4, S)), NE), NW), SW), SE), 6);
    out.gx[c] = (T)gx;
    out.gy[c] = (T)gy;
    out.lap[c] = (T)lap;
    T *o = P.dst + lat_idx(P.Hp, xl, 0, y);
    if (is_solid(P, xl, y)) {  // the reference's masked arrays have no entry here; populations are zero (fingering_periodic.py:108-109)
        out.rho[c] = out.ux[c] = out.uy[c] = out.p[c] = out.mu[c] = out.mix_tau[c] = T(0);
#pragma unroll
        for (int i = 0; i < NPOP; ++i) o[(size_t)i * P.Hp] = T(0);
        return;
    }
    const double psi = C;
    const double rho = I.have_rho ? (double)out.rho[c] : I.rho0;
    // fingering_periodic.py:111,117 -- mu is still zero when p (and everything after it) is computed;
    // fingering.py:121-123 -- p with mu = 0, THEN mu, THEN uy from mu * nabla_psiy (ux stays 0)
    const double p = add(mul(1.0 / 3, rho), mul(psi, 0.0));
    double mu = 0.0, uy = 0.0;
    if (I.variant == 2) {
        mu = sub(mul(mul(I.a, psi), sub(1.0, mul(psi, psi))), mul(I.kappa, lap));
        uy = div(add(0.0, div(mul(mu, gy), 2)), rho);
    }
    // tau_mix, fingering_periodic.py:201-208
    const double v1 = div(I.Eta_n, rho), v2 = div(mul(I.Eta_n, I.M), rho);
    const double mix_v = div(mul(mul(2, v1), v2), add(mul(v1, sub(1.0, psi)), mul(v2, add(1.0, psi))));
    const double mix_tau = add(mul(3, mix_v), 0.5);
    out.rho[c] = (T)rho;
    out.ux[c] = T(0);
    out.uy[c] = (T)uy;
    out.p[c] = (T)p;
    out.mu[c] = (T)mu;
    out.mix_tau[c] = (T)mix_tau;
    // f = f_eq, g = g_eq (fingering_periodic.py:119-121 with 155-192)
    const double w0 = 4.0 / 9;
    const double c0 = mul(3.0, sub(1.0, w0));
    const double A0 = div(sub(rho, mul(c0, p)), w0), A18 = mul(3, p);
    const double B0 = div(sub(psi, mul(mul(c0, I.gamma), mu)), w0), B18 = mul(mul(3, I.gamma), mu);
    const double usq = add(mul(0.0, 0.0), mul(uy, uy));
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int ey = (i == 2 || i == 5 || i == 6) ? 1 : ((i == 4 || i == 7 || i == 8) ? -1 : 0);
        const double w = i == 0 ? w0 : (i < 5 ? 1.0 / 9 : 1.0 / 36);
        const double eu = add(0.0, mul((double)ey, uy));  // e_x * ux = 0 exactly
        const double poly = sub(add(mul(3, eu), mul(4.5, mul(eu, eu))), mul(1.5, usq));
        o[(size_t)i * P.Hp] = (T)mul(w, add(i == 0 ? A0 : A18, mul(rho, poly)));
        o[(size_t)(9 + i) * P.Hp] = (T)mul(w, add(i == 0 ? B0 : B18, mul(psi, poly)));
    }
}

}  // namespace fdlbm
