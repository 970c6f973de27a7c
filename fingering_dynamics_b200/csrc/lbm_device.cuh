// lbm_device.cuh -- device-side building blocks of the two-phase D2Q9 step (sm_100a).
//
// Device layout (DESIGN.md "HBM layout"): one "column record" per lattice column x (the flow axis is
// the SLOW axis), holding the 18 populations f0..f8,g0..g8 as 18 runs of Hp reals along y:
//     lattice[((xl + G) * 18 + pop) * Hp + y],   xl in [-G, Wl+G),  G = 2 ghost columns per side.
// y is contiguous, so a warp reads 32 consecutive y of one population: fully coalesced; the
// y-periodic wrap lives on the fast axis and the slab halo (2 columns) is ONE contiguous block.
// Per-cell flags: reflect bits (uint8, bit i-1 = "direction i is bounced back") and the solid-mask
// bitfield (1 bit per cell, 32 cells along y per word).
//
// The state kept between steps is the POST-collision populations f*, g*.  One step is
//   pull (stream, periodic) -> reflect bits (half-way bounce-back) -> Zou-He faces -> moments
//   -> psi stencils -> collide,
// i.e. the reference iteration (fingering_periodic.py:455-479) rotated by the collision.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fdlbm {

constexpr int G = 2;      // ghost columns per side
constexpr int NPOP = 18;  // f0..f8, g0..g8

template <typename T>
struct LbmParams {
    const T *src;             // lattice read (post-collision state)
    T *dst;                   // lattice written
    const uint8_t *reflect;   // [(xl+G)*Hp + y]
    const uint32_t *solid;    // [(xl+G)*(Hp/32) + y/32], bit y%32
    const T *psi_old;         // psi of the state in src (Zou-He needs its gradient), [(xl+G)*Hp + y]
    T *psi_new;               // psi after this step's streaming
    const T *inlet_ux, *outlet_ux;
    int H, Hp, Wl, gx0, W;    // gx0 = global column of local column 0
    int y_wall, x_periodic, zou_he;
    T inv_tau, gamma, a, kappa, eta6m, M, psi_wall, psi_left, psi_right, f3coef;
    // slab neighbours' dst lattices, mapped over NVLink (peer memory), or nullptr: the step kernel stores its
    // two edge columns straight into their ghost columns (halo exchange fused into the step)
    T *peer_lo, *peer_hi;
    int peer_lo_Wl;           // owned columns of the low-side neighbour (its high ghosts are columns Wl, Wl+1)
};

// macroscopic outputs of the finalize pass / inputs of the first collision, each [(xl+G)*Hp + y]
template <typename T>
struct FieldPtrs {
    T *rho, *ux, *uy, *p, *mu, *mix_tau, *gx, *gy, *lap;
};

__device__ __forceinline__ size_t lat_idx(int Hp, int xl, int pop, int y)
{
    return ((size_t)(xl + G) * NPOP + pop) * (size_t)Hp + y;
}
__device__ __forceinline__ size_t cell_idx(int Hp, int xl, int y) { return (size_t)(xl + G) * Hp + y; }

template <typename T>
__device__ __forceinline__ bool is_solid(const LbmParams<T> &P, int xl, int y)
{
    return (P.solid[(size_t)(xl + G) * (P.Hp >> 5) + (y >> 5)] >> (y & 31)) & 1u;
}

// D2Q9 numbering of the reference (fingering_periodic.py:53-89):
// e0=(0,0) e1=(1,0) e2=(0,1) e3=(-1,0) e4=(0,-1) e5=(1,1) e6=(-1,1) e7=(-1,-1) e8=(1,-1)
__device__ __forceinline__ constexpr int opp(int i)
{
    return i == 0 ? 0 : (i <= 4 ? ((i + 1) & 3) + 1 : ((i - 3) & 3) + 5);
}
static_assert(opp(1) == 3 && opp(2) == 4 && opp(3) == 1 && opp(4) == 2, "axis opposites");
static_assert(opp(5) == 7 && opp(6) == 8 && opp(7) == 5 && opp(8) == 6, "diagonal opposites");

// Pull-stream + half-way bounce-back of one population family (base = 0: f, 9: g) at (xl, y):
// v_i <- src_i(x - e_i) with wrap in y (x wraps through the ghost columns), then for every set
// reflect bit v_i <- src_opp(i)(x)  [fingering_periodic.py:327-343 + bounce_back.py:89-167].
template <typename T>
__device__ __forceinline__ void pull_hp(const LbmParams<T> &P, int Hp, int xl, int y, int base, unsigned bits, T v[9])
{
    // Hp is passed separately: when the caller knows it at compile time every population offset below
    // folds into the load's immediate field
    const int H = P.H;
    int ym = y - 1;
    if (ym < 0) ym += H;
    int yp = y + 1;
    if (yp >= H) yp -= H;
    const T *c0 = P.src + lat_idx(Hp, xl, base, 0);
    const T *cm = c0 - (size_t)NPOP * Hp;  // column x-1: source of e_x = +1
    const T *cp = c0 + (size_t)NPOP * Hp;  // column x+1: source of e_x = -1
    // streamed source of each direction
    const T *s1 = cm + 1 * Hp + y, *s2 = c0 + 2 * Hp + ym, *s3 = cp + 3 * Hp + y, *s4 = c0 + 4 * Hp + yp;
    const T *s5 = cm + 5 * Hp + ym, *s6 = cp + 6 * Hp + ym, *s7 = cp + 7 * Hp + yp, *s8 = cm + 8 * Hp + yp;
    if (bits) {
        // bounced-back directions read the opposite population of the cell itself: select the ADDRESS, so
        // that every register is the target of exactly one load (no write-after-write wait on a load in flight)
        const T *o = c0 + y;
        if (bits & 0x01u) s1 = o + 3 * Hp;
        if (bits & 0x02u) s2 = o + 4 * Hp;
        if (bits & 0x04u) s3 = o + 1 * Hp;
        if (bits & 0x08u) s4 = o + 2 * Hp;
        if (bits & 0x10u) s5 = o + 7 * Hp;
        if (bits & 0x20u) s6 = o + 8 * Hp;
        if (bits & 0x40u) s7 = o + 5 * Hp;
        if (bits & 0x80u) s8 = o + 6 * Hp;
    }
#ifdef FDLBM_LDCS   // streaming (evict-first) loads: every population is read exactly once per step
    v[0] = __ldcs(c0 + y);
    v[1] = __ldcs(s1);
    v[2] = __ldcs(s2);
    v[3] = __ldcs(s3);
    v[4] = __ldcs(s4);
    v[5] = __ldcs(s5);
    v[6] = __ldcs(s6);
    v[7] = __ldcs(s7);
    v[8] = __ldcs(s8);
#else
    v[0] = c0[y];
    v[1] = *s1;
    v[2] = *s2;
    v[3] = *s3;
    v[4] = *s4;
    v[5] = *s5;
    v[6] = *s6;
    v[7] = *s7;
    v[8] = *s8;
#endif
}

template <typename T>
__device__ __forceinline__ void pull(const LbmParams<T> &P, int xl, int y, int base, unsigned bits, T v[9])
{
    pull_hp(P, P.Hp, xl, y, base, bits, v);
}

// Zou-He rules for g on the faces (fingering_periodic.py:278-324, fingering.py:305-390).
template <typename T>
__device__ __forceinline__ void zou_he_g(const LbmParams<T> &P, int gx, int y, T g[9])
{
    // weight ratios of the reference's rules: w1/(w1+w5+w8) = 2/3, w5/(w1+w5+w8) = 1/6, w6/(w6+w8) = 1/2
    const T r1 = T(2) / T(3), r5 = T(1) / T(6);
    if (gx == 0) {
        const bool fg = P.zou_he == 2;
        if (fg && y == 0) {  // fingering.py:336-348
            g[1] = g[3];
            g[2] = g[4];
            g[5] = g[7];
            g[6] = T(0.5) * (T(1) - (g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[7]));
            g[8] = g[6];
        } else if (fg && y == P.H - 1) {  // fingering.py:352-364
            g[1] = g[3];
            g[4] = g[2];
            g[8] = g[6];
            g[5] = T(0.5) * (T(1) - (g[0] + g[1] + g[2] + g[3] + g[4] + g[6] + g[8]));
            g[7] = g[5];
        } else {
            const T psi_in = P.psi_left - (g[0] + g[2] + g[3] + g[4] + g[6] + g[7]);
            g[1] = r1 * psi_in;
            g[5] = r5 * psi_in;
            g[8] = g[5];
        }
    }
    if (gx == P.W - 1) {
        const T psi_out = P.psi_right - (g[0] + g[1] + g[2] + g[4] + g[5] + g[8]);
        g[3] = r1 * psi_out;
        g[6] = r5 * psi_out;
        g[7] = g[6];
        if (P.zou_he == 2) {  // fingering.py:387-390
            if (y == 0) g[2] = g[4];
            if (y == P.H - 1) g[4] = g[2];
        }
    }
}

// psi as the stencils see it, including ghost rows / columns
// (fingering_periodic.py:216-218, fingering.py:221-224, validation.py:246).
template <typename T>
__device__ __forceinline__ T psi_fetch(const LbmParams<T> &P, const T *psi, int xl, int y)
{
    if (y < 0 || y >= P.H) {
        if (P.y_wall) return P.psi_wall;
        y = y < 0 ? y + P.H : y - P.H;
    }
    if (!P.x_periodic) {
        const int gx = P.gx0 + xl;
        if (gx < 0) return P.psi_left;
        if (gx >= P.W) return P.psi_right;
    }
    return psi[cell_idx(P.Hp, xl, y)];
}

// isotropic 9-point stencils (fingering_periodic.py:214-256)
template <typename T>
__device__ __forceinline__ void stencil9(T C, T E, T Wv, T N, T S, T NE, T NW, T SW, T SE, T &gx, T &gy, T &lap)
{
    gx = (T(4) * (E - Wv) + ((NE - NW) + (SE - SW))) * (T(1) / T(12));
    gy = (T(4) * (N - S) + ((NE - SE) + (NW - SW))) * (T(1) / T(12));
    lap = (T(-20) * C + T(4) * ((N + S) + (E + Wv)) + ((NE + NW) + (SW + SE))) * (T(1) / T(6));
}

template <typename T>
__device__ __forceinline__ void stencil_from_array(const LbmParams<T> &P, const T *psi, int xl, int y, T &gx,
                                                   T &gy, T &lap)
{
    const T C = psi_fetch(P, psi, xl, y);
    const T E = psi_fetch(P, psi, xl + 1, y), Wv = psi_fetch(P, psi, xl - 1, y);
    const T N = psi_fetch(P, psi, xl, y + 1), S = psi_fetch(P, psi, xl, y - 1);
    const T NE = psi_fetch(P, psi, xl + 1, y + 1), NW = psi_fetch(P, psi, xl - 1, y + 1);
    const T SW = psi_fetch(P, psi, xl - 1, y - 1), SE = psi_fetch(P, psi, xl + 1, y - 1);
    stencil9(C, E, Wv, N, S, NE, NW, SW, SE, gx, gy, lap);
}

template <typename T>
__device__ __forceinline__ T chem_potential(const LbmParams<T> &P, T psi, T lap)
{
    return P.a * psi * (T(1) - psi * psi) - P.kappa * lap;  // fingering_periodic.py:141-149
}

// rho on the inlet face (fingering_periodic.py:276-277), from post-stream/post-bounce-back f of the
// face cell and the PREVIOUS psi's gradient and chemical potential.
template <typename T>
__device__ __forceinline__ T inlet_rho(const T f[9], T ux, T psx, T mu)
{
    return (f[0] + f[2] + f[4] + T(2) * (f[3] + f[6] + f[7]) - psx * mu * T(0.5)) / (T(1) - ux);
}

// rho_inlet of ANOTHER row of the inlet column (needed by the two FG corner nodes, fingering.py:341,357).
// PullRow: populations of that row come from streaming (the step kernels).
template <typename T>
struct PullRow {
    __device__ __noinline__ T operator()(const LbmParams<T> &P, int xl, int y) const
    {
        T f[9];
        pull(P, xl, y, 0, P.reflect[cell_idx(P.Hp, xl, y)], f);
        T psx, psy, lap;
        stencil_from_array(P, P.psi_old, xl, y, psx, psy, lap);
        const T mu = chem_potential(P, P.psi_old[cell_idx(P.Hp, xl, y)], lap);
        return inlet_rho(f, P.inlet_ux[y], psx, mu);
    }
};

// Zou-He rules for f on the faces (fingering_periodic.py:268-324, fingering.py:298-390).
template <typename T, typename RhoRow>
__device__ __forceinline__ void zou_he_f(const LbmParams<T> &P, int xl, int gx, int y, T f[9], RhoRow rho_row)
{
    T psx, psy, lap;
    stencil_from_array(P, P.psi_old, xl, y, psx, psy, lap);
    const T mu = chem_potential(P, P.psi_old[cell_idx(P.Hp, xl, y)], lap);
    const T sixth = T(1) / T(6);
    if (gx == 0) {
        const bool fg = P.zou_he == 2;
        if (fg && y == 0) {  // fingering.py:335-347
            const T rho1 = rho_row(P, xl, 1);
            f[1] = f[3];
            f[2] = f[4];
            f[5] = f[7];
            f[6] = T(0.5) * (rho1 - (f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[7]));
            f[8] = f[6];
        } else if (fg && y == P.H - 1) {  // fingering.py:351-363
            const T rho2 = rho_row(P, xl, P.H - 2);
            f[1] = f[3];
            f[4] = f[2];
            f[8] = f[6];
            f[5] = T(0.5) * (rho2 - (f[0] + f[1] + f[2] + f[3] + f[4] + f[6] + f[8]));
            f[7] = f[5];
        } else {
            const T ux = P.inlet_ux[y];
            const T rho_in = inlet_rho(f, ux, psx, mu);
            const T half_d = T(0.5) * (f[2] - f[4]);
            const T ur = ux * rho_in;
            f[1] = f[3] + (T(2) / T(3)) * ur - psx * mu * sixth;
            f[5] = f[7] - half_d + sixth * ur - psx * mu * sixth - psy * mu * T(0.25);
            f[8] = f[6] + half_d + sixth * ur - psx * mu * sixth + psy * mu * T(0.25);
        }
    }
    if (gx == P.W - 1) {
        const T ux = P.outlet_ux[y];
        const T rho_out = (f[0] + f[2] + f[4] + T(2) * (f[1] + f[5] + f[8]) + psx * mu * T(0.5)) / (T(1) + ux);
        const T half_d = T(0.5) * (f[2] - f[4]);
        const T ur = ux * rho_out;
        f[3] = f[1] - P.f3coef * ur + psx * mu * sixth;
        f[6] = f[8] - half_d - sixth * ur + psy * mu * T(0.25) + psx * mu * sixth;
        f[7] = f[5] + half_d - sixth * ur - psy * mu * T(0.25) + psx * mu * sixth;
        if (P.zou_he == 2) {  // fingering.py:387-390
            if (y == 0) f[2] = f[4];
            if (y == P.H - 1) f[4] = f[2];
        }
    }
}

// Post-stream, post-boundary g of cell (xl,y) and the new order parameter
// psi = sum g on fluid, psi_wall on solids (fingering_periodic.py:210-212).
template <typename T>
__device__ __forceinline__ T stream_bc_g(const LbmParams<T> &P, int xl, int y, unsigned bits, T g[9])
{
    pull(P, xl, y, 9, bits, g);
    if (P.zou_he) {
        const int gx = P.gx0 + xl;
        if (gx == 0 || gx == P.W - 1) zou_he_g(P, gx, y, g);
    }
    if (is_solid(P, xl, y)) return P.psi_wall;
    return (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) + g[8];
}

template <typename T>
__device__ __forceinline__ void stream_bc_f(const LbmParams<T> &P, int xl, int y, unsigned bits, T f[9])
{
    pull(P, xl, y, 0, bits, f);
    if (P.zou_he) {
        const int gx = P.gx0 + xl;
        if (gx == 0 || gx == P.W - 1) zou_he_f(P, xl, gx, y, f, PullRow<T>());
    }
}

__device__ __forceinline__ double rcp_exact(double x) { return 1.0 / x; }
#ifdef FDLBM_F32_IEEE_DIV
__device__ __forceinline__ float rcp_exact(float x) { return 1.0f / x; }
#else
__device__ __forceinline__ float rcp_exact(float x) { return __frcp_rn(x); }
#endif

// Macroscopic moments of one fluid cell (fingering_periodic.py:123-152, 201-208).
template <typename T>
struct Macro {
    T rho, ux, uy, p, mu, inv_mt, psi, gx, gy;
};

template <typename T>
__device__ __forceinline__ void moments(const LbmParams<T> &P, const T f[9], T psi, T gx, T gy, T lap, Macro<T> &m)
{
    m.psi = psi;
    m.gx = gx;
    m.gy = gy;
    m.rho = (((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]))) + f[8];
    m.mu = chem_potential(P, psi, lap);
    const T jx = (f[1] - f[3]) + ((f[5] - f[6]) + (f[8] - f[7]));
    const T jy = (f[2] - f[4]) + ((f[5] - f[8]) + (f[6] - f[7]));
    // tau_mix = 6 Eta_n M / (rho D) + 1/2, D = (1-psi) + M (1+psi): fingering_periodic.py:201-208 with
    // v1, v2 substituted.  One division serves 1/rho and 1/tau_mix:  r = 1 / (rho * X),
    // X = eta6m + rho D / 2  =>  1/rho = r X,  1/tau_mix = rho D / X = rho^2 D r.
    const T D = (T(1) - psi) + P.M * (T(1) + psi);
    const T rD = m.rho * D;
    const T X = P.eta6m + T(0.5) * rD;
    // fp64: IEEE division (a MUFU.RCP64H seed + Newton steps by hand measured 2-12 % SLOWER on B200, A/B in one
    // run); fp32: correctly rounded reciprocal without the division's slow-path call
    const T r = rcp_exact(m.rho * X);
    const T inv_rho = r * X;
    m.inv_mt = rD * m.rho * r;
    m.ux = (jx + T(0.5) * m.mu * gx) * inv_rho;
    m.uy = (jy + T(0.5) * m.mu * gy) * inv_rho;
    m.p = m.rho * (T(1) / T(3)) + psi * m.mu;
}

// BGK collision with the Guo-type forcing term (fingering_periodic.py:155-199, 258-264), in place.
template <typename T>
__device__ __forceinline__ void collide(const LbmParams<T> &P, const Macro<T> &m, T f[9], T g[9])
{
    const T w0 = T(4) / T(9), w1 = T(1) / T(9), w5 = T(1) / T(36);
    const T c0 = T(5) / T(3);  // 3 (1 - w0)
    const T usq15 = T(1.5) * (m.ux * m.ux + m.uy * m.uy);
    const T pref = T(1) - T(0.5) * m.inv_mt;
    const T Fx = m.mu * pref * m.gx, Fy = m.mu * pref * m.gy;  // force x (1 - 1/(2 tau_mix))
    const T uF = m.ux * Fx + m.uy * Fy;
    const T om_f = m.inv_mt, om_g = P.inv_tau;
    const T p3 = T(3) * m.p, gm3 = T(3) * P.gamma * m.mu;
    {  // i = 0
        const T feq = m.rho - c0 * m.p - w0 * m.rho * usq15;
        const T geq = m.psi - c0 * P.gamma * m.mu - w0 * m.psi * usq15;
        const T Fi = w0 * (T(-3) * uF);
        f[0] = f[0] - om_f * (f[0] - feq) + Fi;
        g[0] = g[0] - om_g * (g[0] - geq);
    }
#pragma unroll
    for (int i = 1; i < 9; ++i) {
        const int ex = (i == 1 || i == 5 || i == 8) ? 1 : ((i == 3 || i == 6 || i == 7) ? -1 : 0);
        const int ey = (i == 2 || i == 5 || i == 6) ? 1 : ((i == 4 || i == 7 || i == 8) ? -1 : 0);
        const T w = i < 5 ? w1 : w5;
        const T eu = T(ex) * m.ux + T(ey) * m.uy;
        const T eF = T(ex) * Fx + T(ey) * Fy;
        const T poly = T(3) * eu + T(4.5) * eu * eu - usq15;
        const T feq = w * (p3 + m.rho * poly);
        const T geq = w * (gm3 + m.psi * poly);
        const T Fi = w * (T(3) * (eF - uF) + T(9) * eu * eF);
        f[i] = f[i] - om_f * (f[i] - feq) + Fi;
        g[i] = g[i] - om_g * (g[i] - geq);
    }
}

template <typename T>
__device__ __forceinline__ void store_cell_at(T *lat, int Hp, int xl, int y, const T f[9], const T g[9])
{
    T *o = lat + lat_idx(Hp, xl, 0, y);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#ifdef FDLBM_STCS   // streaming stores: the written lattice is not read again before the next step
        __stcs(o + (size_t)i * Hp, f[i]);
        __stcs(o + (size_t)(9 + i) * Hp, g[i]);
#else
        o[(size_t)i * Hp] = f[i];
        o[(size_t)(9 + i) * Hp] = g[i];
#endif
    }
}

template <typename T>
__device__ __forceinline__ void store_cell_hp(const LbmParams<T> &P, int Hp, int xl, int y, const T f[9], const T g[9])
{
    store_cell_at(P.dst, Hp, xl, y, f, g);
    // halo push: the two edge columns also land in the neighbours' ghost columns (peer stores over NVLink)
    if (P.peer_lo && xl < G) store_cell_at(P.peer_lo, Hp, P.peer_lo_Wl + xl, y, f, g);
    if (P.peer_hi && xl >= P.Wl - G) store_cell_at(P.peer_hi, Hp, xl - P.Wl, y, f, g);
}

template <typename T>
__device__ __forceinline__ void store_cell(const LbmParams<T> &P, int xl, int y, const T f[9], const T g[9])
{
    store_cell_hp(P, P.Hp, xl, y, f, g);
}

}  // namespace fdlbm
