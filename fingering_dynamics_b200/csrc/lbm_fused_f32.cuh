// lbm_fused_f32.cuh -- the fused step kernel for fp32 storage: two rows per thread, PACKED arithmetic.
//
// Same algorithm as k_fused (lbm_fused.cuh): a CTA owns a strip of rows and marches along x; the g and f columns
// stream through two shared-memory stage rings filled by bulk copies (cp.async.bulk + mbarrier, one column ahead),
// psi of the rows above / below arrives by warp shuffle, one barrier per column.  In fp32 the memory time per
// column is half of fp64's and the scalar two-row kernel was bound by INSTRUCTION ISSUE (profiles/README.md:
// k_fused_vec<float> issued 1354 warp instructions per 64-cell warp column, 52 % of them integer / address /
// control).  This kernel cuts the instruction count:
//
//   * a thread owns the aligned row pair (yb, yb+1), both rows live in the two halves of float2 registers and
//     moments + collision run on the packed fp32 pipe (FFMA2 / FADD2 / FMUL2: one issue slot for two cells);
//   * the collision is written in its even / odd form over the four direction pairs (i, opp(i)): the even part
//     of f_eq, g_eq and of the forcing term is shared by both directions of a pair;
//   * every global address is ONE per-thread pointer plus an immediate: running pointers advance by the column
//     stride, population and row offsets fold into the instruction (the row pitch Hp is a template parameter for
//     the common heights; Hp = 0 reads it from the parameters and pays one IMAD per access); the four stage
//     pointers of columns x-1 .. x+2 rotate instead of being recomputed;
//   * bounce-back is a pair of complementary predicated loads into the same register ("@p ld bounce; @!p ld
//     stream"), no address selects; warps without a bounce-back cell take a branch with 8-byte loads;
//   * the two edge lanes of a warp evaluate their outer neighbour rows in ONE merged pass;
//   * the columns next to the Zou-He faces / the domain edge run the scalar code in FACE CTAs of the same grid;
//   * strips that fill in bulk and have no one-row-pair warp run a LEAN instantiation of the column loop (16.6 KB of
//     code instead of 25 KB of the 32 KB instruction-cache level).
//
// Requires an even grid height (aligned row pairs); odd heights use k_fused_vec.
//
// Measured on B200 (8192x2048; profiles/README.md, profiles/r2/): 41.0 GLUPS = 0.914 of the HBM roofline (round 1:
// 38.5 = 0.859; the scalar two-row kernel 30.0 = 0.66).  The fp64 kernel reaches 0.99 with the same memory pattern per
// CTA and column; what keeps fp32 below it is the per-warp instruction chain, about as long as the column's transfer
// time.  Measured and NOT adopted (round 1 + 2): f loads issued one column ahead -1.6 %; f in its own registers with g
// reloaded from a five-stage ring -4 %; L2 prefetch of f / g 4 columns ahead -13 % / -9 %, bulk L2 prefetch 1 / 2 / 4
// columns ahead of the stage fill +1.6 / -6 / -26 %; eight g stages (2 CTAs/SM) -4 %; warp-interleaved rows (no
// shared-memory bank conflicts, but two 4-byte stores per population) -4 %; flag loads earlier in the column 0 %, no
// flag loads at all (timing mock) -1.7 %; two steps per pass (mock): upper bound 1.11x before overheads.
#pragma once
#include "lbm_fused_vec.cuh"

namespace fdlbm {
#ifndef FDLBM_F32_BULK
#define FDLBM_F32_BULK 1  // g stages filled by cp.async.bulk + mbarrier as in k_fused (measured +2.8 % over per-thread cp.async; 0 for A/B)
#endif
namespace f32p {

typedef float2 p2;
#define FDLBM_DI __device__ __forceinline__
FDLBM_DI p2 mk(float a, float b) { return make_float2(a, b); }
FDLBM_DI p2 bc(float a) { return make_float2(a, a); }
FDLBM_DI p2 add(p2 a, p2 b) { return __fadd2_rn(a, b); }
FDLBM_DI p2 mul(p2 a, p2 b) { return __fmul2_rn(a, b); }
FDLBM_DI p2 fma2(p2 a, p2 b, p2 c) { return __ffma2_rn(a, b, c); }
FDLBM_DI p2 sub(p2 a, p2 b) { return __ffma2_rn(b, bc(-1.0f), a); }

// v <- bit ? *bounce : *stream, as two complementary predicated loads into one register (no address select;
// ptxas puts both on one scoreboard and does not serialise them)
FDLBM_DI float lds_pick(const float *stream, const float *bounce, unsigned bit)
{
    float v;
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q ld.shared.f32 %0, [%2];\n @!q ld.shared.f32 %0, [%1];\n}"
                 : "=f"(v)
                 : "r"(smem_u32(stream)), "r"(smem_u32(bounce)), "r"(bit));
    return v;
}

// unpredicated loads as volatile asm: they keep their place AFTER the predicated pairs.  ptxas turns a predicated
// load at the end of a block into "branch around + plain load"; the plain load then no longer pairs with its
// complement and waits for it (a full memory latency per column, measured) -- so the tail is never a pair.
FDLBM_DI p2 lds_v2(const float *p)
{
    p2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p)));
    return v;
}
FDLBM_DI float lds_f(const float *p)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
    return v;
}

struct Cfg {
    static constexpr int NT = 128, ROWS = 256, HALO = 4, PT = ROWS + 2 * HALO, NS = 4, FAM = 9 * PT;
    static constexpr int RINGS = 2;  // g ring + f ring
    static constexpr size_t SMEM = (size_t)RINGS * NS * FAM * sizeof(float);
};

// moments of the cell pair (fingering_periodic.py:123-152, 201-208), packed; see moments() in lbm_device.cuh
struct Macro2 {
    p2 rho, ux, uy, p, mu, inv_mt, psi, gx, gy;
};

FDLBM_DI void moments2(const LbmParams<float> &P, const p2 f[9], p2 psi, p2 gx, p2 gy, p2 lap, bool solid0, bool solid1,
                       Macro2 &m)
{
    m.psi = psi, m.gx = gx, m.gy = gy;
    p2 rho = add(add(add(f[0], f[1]), add(f[2], f[3])), add(add(f[4], f[5]), add(f[6], f[7])));
    rho = add(rho, f[8]);
    // solid rows are streamed, never collided: the collision below runs with omega = 0 and no force for them,
    // on a harmless density (what streams into a solid may sum to 0)
    if (solid0) rho.x = 1.0f;
    if (solid1) rho.y = 1.0f;
    m.rho = rho;
    const p2 t = fma2(mul(psi, bc(-1.0f)), psi, bc(1.0f));                       // 1 - psi^2
    m.mu = fma2(mul(psi, bc(P.a)), t, mul(lap, bc(-P.kappa)));                    // a psi (1 - psi^2) - kappa lap
    const p2 jx = add(sub(f[1], f[3]), add(sub(f[5], f[6]), sub(f[8], f[7])));
    const p2 jy = add(sub(f[2], f[4]), add(sub(f[5], f[8]), sub(f[6], f[7])));
    // tau_mix: one reciprocal serves 1/rho and 1/tau_mix (see moments())
    const p2 D = fma2(psi, bc(P.M - 1.0f), bc(P.M + 1.0f));                       // (1 - psi) + M (1 + psi)
    const p2 rD = mul(rho, D);
    const p2 X = fma2(rD, bc(0.5f), bc(P.eta6m));
    const p2 den = mul(rho, X);
    const p2 r = mk(__frcp_rn(den.x), __frcp_rn(den.y));
    const p2 inv_rho = mul(r, X);
    m.inv_mt = mul(mul(rD, rho), r);
    const p2 hm = mul(m.mu, bc(0.5f));
    m.ux = mul(fma2(hm, gx, jx), inv_rho);
    m.uy = mul(fma2(hm, gy, jy), inv_rho);
    m.p = fma2(psi, m.mu, mul(rho, bc(1.0f / 3.0f)));
}

// BGK collision + Guo forcing of the cell pair (fingering_periodic.py:155-199, 258-264) in even / odd form:
// for a direction pair (i, o = opp(i)), e_o = -e_i, so with eu = e_i.u, eF = e_i.F
//   f_eq(i,o) = E +- O,  E = w (3p + rho (4.5 eu^2 - 1.5 u^2)),  O = 3 w rho eu       (same for g_eq with psi, gamma mu)
//   F(i,o)    = Fe +- Fo, Fe = w (9 eu eF - 3 u.F),  Fo = 3 w eF
//   f_i <- (1 - om) f_i + (om E + Fe) +- (om O + Fo)
FDLBM_DI void collide2(const LbmParams<float> &P, const Macro2 &m, bool solid0, bool solid1, p2 f[9], p2 g[9])
{
    const float w0 = 4.0f / 9.0f, w1 = 1.0f / 9.0f, w5 = 1.0f / 36.0f, c0 = 5.0f / 3.0f;
    p2 om = m.inv_mt;                       // 1 / tau_mix
    p2 omg = bc(P.inv_tau);
    p2 pref = fma2(om, bc(-0.5f), bc(1.0f));  // 1 - 1/(2 tau_mix)
    if (solid0) om.x = 0.0f, omg.x = 0.0f, pref.x = 0.0f;
    if (solid1) om.y = 0.0f, omg.y = 0.0f, pref.y = 0.0f;
    const p2 mp = mul(m.mu, pref);
    const p2 Fx = mul(mp, m.gx), Fy = mul(mp, m.gy);
    const p2 uF = fma2(m.ux, Fx, mul(m.uy, Fy));
    const p2 nusq15 = mul(fma2(m.ux, m.ux, mul(m.uy, m.uy)), bc(-1.5f));  // -1.5 u^2
    const p2 c1 = fma2(om, bc(-1.0f), bc(1.0f));                          // 1 - om
    const p2 c1g = fma2(omg, bc(-1.0f), bc(1.0f));
    const p2 p3 = mul(m.p, bc(3.0f));
    const p2 gm3 = mul(m.mu, bc(3.0f * P.gamma));
    const p2 og_psi = mul(omg, m.psi);                                    // om_g psi
    const p2 og_gm3 = mul(omg, gm3);                                      // om_g 3 gamma mu
    {  // rest direction
        const p2 feq = fma2(mul(m.rho, bc(w0)), nusq15, fma2(m.p, bc(-c0), m.rho));
        const p2 geq = fma2(mul(m.psi, bc(w0)), nusq15, fma2(m.mu, bc(-c0 * P.gamma), m.psi));
        const p2 Fi = mul(uF, bc(-3.0f * w0));
        f[0] = add(fma2(om, sub(feq, f[0]), f[0]), Fi);
        g[0] = fma2(omg, sub(geq, g[0]), g[0]);
    }
    const p2 n3uF = mul(uF, bc(-3.0f));
    // per weight class: w rho, 3 w rho, w 3p, -3 w u.F, w om_g psi, 3 w om_g psi, w om_g 3 gamma mu
    const p2 wr1 = mul(m.rho, bc(w1)), wr5 = mul(m.rho, bc(w5)), wr31 = mul(m.rho, bc(3.0f * w1)), wr35 = mul(m.rho, bc(3.0f * w5));
    const p2 wp1 = mul(p3, bc(w1)), wp5 = mul(p3, bc(w5)), wF1 = mul(n3uF, bc(w1)), wF5 = mul(n3uF, bc(w5));
    const p2 gp1 = mul(og_psi, bc(w1)), gp5 = mul(og_psi, bc(w5)), gp31 = mul(og_psi, bc(3.0f * w1)), gp35 = mul(og_psi, bc(3.0f * w5));
    const p2 gg1 = mul(og_gm3, bc(w1)), gg5 = mul(og_gm3, bc(w5));
#define FDLBM_PAIR(I, O, W, EU, EF)                                                                  \
    {                                                                                                \
        const bool ax = (W) > 0.05f;                                                                 \
        const p2 eu = (EU), eF = (EF);                                                               \
        const p2 A = fma2(mul(eu, bc(4.5f)), eu, nusq15);                                            \
        const p2 E = fma2(ax ? wr1 : wr5, A, ax ? wp1 : wp5);                                        \
        const p2 Od = mul(ax ? wr31 : wr35, eu);                                                     \
        const p2 Fe = fma2(mul(eu, bc(9.0f * (W))), eF, ax ? wF1 : wF5);                             \
        const p2 se = fma2(om, E, Fe), so = fma2(om, Od, mul(eF, bc(3.0f * (W))));                   \
        f[I] = fma2(c1, f[I], add(se, so));                                                          \
        f[O] = fma2(c1, f[O], sub(se, so));                                                          \
        const p2 Eg = fma2(ax ? gp1 : gp5, A, ax ? gg1 : gg5);                                       \
        const p2 Og = mul(ax ? gp31 : gp35, eu);                                                     \
        g[I] = fma2(c1g, g[I], add(Eg, Og));                                                         \
        g[O] = fma2(c1g, g[O], sub(Eg, Og));                                                         \
    }
    FDLBM_PAIR(1, 3, w1, m.ux, Fx)
    FDLBM_PAIR(2, 4, w1, m.uy, Fy)
    FDLBM_PAIR(5, 7, w5, add(m.ux, m.uy), add(Fx, Fy))
    FDLBM_PAIR(8, 6, w5, sub(m.ux, m.uy), sub(Fx, Fy))
#undef FDLBM_PAIR
}

// The launch grid: `n_fast` CTAs march over the plain columns [fx0, fx1) (strip = blockIdx % nyt, chunks of `chunk`
// columns), followed by the FACE CTAs: one per strip and side for the columns next to the Zou-He faces, [0, fx0)
// and [fx1, Wl), which run the scalar code of k_fused_vec (fused_vec_strip).  Face CTAs come last in dispatch
// order and are tiny (2-3 columns): they slip into the CTA slots the one-wave chunking leaves free.
// HPC > 0: compile-time row pitch (every population offset is an immediate); HPC == 0: P.Hp
// BULK: every stage of this strip is filled by bulk copies (the common case: no y wrap inside the apron, or all pieces
// 16-byte multiples) and no warp of it has a single active row pair; the per-thread cp.async fill paths and the in-loop
// flag load of such warps are then compiled OUT of the column loop -- they made its body
// 25 KB of the 32 KB instruction-cache level (ncu r2d: no_instruction 0.49 stalls per issue against 0.17 in the fp64 kernel)
template <int HPC, bool BULK>
__device__ __forceinline__ void fast_strip(const LbmParams<float> &P, const int nyt, const int chunk, const int fx0, const int fx1)
{
    typedef float T;
    constexpr int D = FUSED_D, NS = Cfg::NS, PT = Cfg::PT, HALO = Cfg::HALO, FAM = Cfg::FAM, ROWS = Cfg::ROWS, NT = Cfg::NT;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(D == 1 && NS == 4, "stage ring of four columns, one column ahead");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *gst = reinterpret_cast<T *>(smem_raw);  // [NS][9][PT]
    T *fst = gst + NS * FAM;                   // [NS][9][PT]
    const int t = threadIdx.x, lane = t & 31;
    const int H = P.H, Hp = HPC > 0 ? HPC : P.Hp;
    const ptrdiff_t S = (ptrdiff_t)NPOP * Hp;  // column stride (elements)

    const int yt = blockIdx.x % nyt;
    const int xs = fx0 + (blockIdx.x / nyt) * chunk;
    const int xe = min(fx1, xs + chunk);
    const int y0 = yt * ROWS;
    const int yb = y0 + 2 * t;                  // first row of this thread (even)
    const int ny = min(ROWS, H - y0);           // rows of this strip (even)
    const bool has = 2 * t < ny;                // this thread owns rows yb, yb+1
    const int nv = has ? 2 : 0;
    const int jb = HALO + 2 * t;                // stage row of the thread's first cell

    // outer neighbours of the row pair: row yb-1 and row yb+2.  Inside a warp they belong to lane-1 / lane+1
    // (shuffles); the first and the last active lane of a warp evaluate theirs themselves, in one merged pass
    const bool edge_lo = has && lane == 0;
    const bool edge_hi = has && (lane == 31 || 2 * (t + 1) >= ny);
    // a lane that is both (a warp with ONE active row pair: ny % 64 == 2) loads the flags of its second outer row inside
    // the loop; strips without such a warp run the BULK instantiation, which has no such load (cf. edge2 in lbm_fused.cuh)
    const bool edge = edge_lo || edge_hi, edge2 = !BULK && edge_lo && edge_hi;
    auto wrap_row = [&](int yy) {  // wrapped global row, or -1 for a ghost row of a y-wall variant
        if (yy < 0 || yy >= H) return P.y_wall ? -1 : (yy < 0 ? yy + H : yy - H);
        return yy;
    };
    const int ye_lo = wrap_row(yb - 1), ye_hi = wrap_row(yb + 2);
    const int ye = edge_lo ? ye_lo : ye_hi;      // the merged pass: this lane's outer row ...
    const bool e_ghost = ye < 0;                 // ghost row of a y wall: psi_wall, no flags

    auto slot = [](int c) { return c & (NS - 1); };

    // ---- g stage ring: per-thread 16-byte cp.async, one chunk of every population row per thread ------------
    constexpr int EPC = 4;  // elements per 16-byte chunk
    const bool fill_thread = t < (ny + 2 * HALO + EPC - 1) / EPC;  // chunks cover rows [y0-HALO, y0+ny+HALO)
    const int fy = y0 - HALO + EPC * t;          // first global row of this thread's chunk (before the wrap)
    // a chunk that wraps in y is still one aligned 16-byte run when H % 4 == 0 (rows -4..-1 are rows H-4..H-1): the
    // strips at the wrap then fill as fast as the others -- the kernel ends with its slowest CTA (+1.8 %)
    const int fyw = fy < 0 ? fy + H : (fy >= H ? fy - H : fy);
    const bool fill_fast = fill_thread && ((H % EPC) == 0 || (fy >= 0 && fy + EPC <= H));
#if FDLBM_F32_BULK
    // bulk fill (as in k_fused): one thread issues 9 bulk copies of the contiguous piece of a stage row, completion
    // counted in bytes on the stage's mbarrier; an apron that wraps in y is one 16-byte cp.async per population
    __shared__ __align__(8) unsigned long long bars[NS];
    const bool wrap_lo = y0 - HALO < 0, wrap_hi = y0 + ny + HALO > H;  // CTA-uniform
    const bool bulk = BULK;  // = bulk_strip(H, y0, ny), decided by the kernel
    if (t == 0) {
#pragma unroll
        for (int s_ = 0; s_ < NS; ++s_) mbar_init(&bars[s_], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto landed = [&](int c) {
        if (bulk) mbar_wait(&bars[slot(c)], (unsigned)(((c - (xs - 2)) / NS) & 1));
    };
#else
    constexpr bool bulk = false;
    auto landed = [](int) {};
#endif
    // g column v+2+D and f column v+1+D: f is consumed one column behind g, its ring holds
    // x-1..x+1 plus the column in flight; an f column always travels with a g column (same mbarrier / commit group)
    auto prefetch = [&](int v) {
        const int cg = v + 2 + D;
        if (cg >= xs - 2 && cg <= xe + 1) {
            T *stage = gst + slot(cg) * FAM + EPC * t;
            const T *col = P.src + lat_idx(Hp, cg, 9, 0);
            const bool with_f = cg - 1 >= xs - 1 && cg - 1 <= xe;
            T *fstage = fst + slot(cg - 1) * FAM + EPC * t;
            const T *fcol = P.src + lat_idx(Hp, cg - 1, 0, 0);
#if FDLBM_F32_BULK
            if (bulk) {
                T *st0 = gst + slot(cg) * FAM, *fs0 = fst + slot(cg - 1) * FAM;
                if (t == 0) {
                    const int r0 = wrap_lo ? y0 : y0 - HALO;               // first row of the contiguous piece
                    const int r1 = wrap_hi ? y0 + ny : y0 + ny + HALO;     // one past its last row
                    const unsigned bytes = (unsigned)((r1 - r0) * sizeof(T));
                    mbar_expect_tx(&bars[slot(cg)], (with_f ? 18u : 9u) * bytes);
#pragma unroll
                    for (int pop = 0; pop < 9; ++pop)
                        bulk_g2s(st0 + pop * PT + (r0 - (y0 - HALO)), col + (size_t)pop * Hp + r0, bytes, &bars[slot(cg)]);
                    if (with_f) {
#pragma unroll
                        for (int pop = 0; pop < 9; ++pop)
                            bulk_g2s(fs0 + pop * PT + (r0 - (y0 - HALO)), fcol + (size_t)pop * Hp + r0, bytes, &bars[slot(cg)]);
                    }
                }
                if (wrap_lo && t >= 32 && t < 41) cp_async16(st0 + (t - 32) * PT, col + (size_t)(t - 32) * Hp + (y0 - HALO + H));
                if (wrap_hi && t >= 64 && t < 73)
                    cp_async16(st0 + (t - 64) * PT + HALO + ny, col + (size_t)(t - 64) * Hp + (y0 + ny - H));
                if (with_f) {
                    if (wrap_lo && t >= 41 && t < 50) cp_async16(fs0 + (t - 41) * PT, fcol + (size_t)(t - 41) * Hp + (y0 - HALO + H));
                    if (wrap_hi && t >= 73 && t < 82)
                        cp_async16(fs0 + (t - 73) * PT + HALO + ny, fcol + (size_t)(t - 73) * Hp + (y0 + ny - H));
                }
            } else
#endif
            if (fill_fast) {
                const T *s = col + fyw, *sf = fcol + fyw;
#pragma unroll
                for (int pop = 0; pop < 9; ++pop) cp_async16(stage + pop * PT, s + (ptrdiff_t)pop * Hp);
                if (with_f) {
#pragma unroll
                    for (int pop = 0; pop < 9; ++pop) cp_async16(fstage + pop * PT, sf + (ptrdiff_t)pop * Hp);
                }
            } else if (fill_thread) {
#pragma unroll
                for (int e = 0; e < EPC; ++e) {
                    int yy = (fy + e) % H;
                    if (yy < 0) yy += H;
#pragma unroll
                    for (int pop = 0; pop < 9; ++pop) {
                        cp_async_small<4>(stage + pop * PT + e, col + (size_t)pop * Hp + yy);
                        if (with_f) cp_async_small<4>(fstage + pop * PT + e, fcol + (size_t)pop * Hp + yy);
                    }
                }
            }
        }
        cp_async_commit();
    };

    // ---- flags: reflect bytes of the row pair (one 16-bit load) + solid-mask word; one outer row for edge lanes --
    auto load_flags = [&](int c, int yy, int n) -> RawFlags {  // indexed form (warm-up, one-lane warps)
        RawFlags r{0u, 0u};
        if (c > xe + 1 || yy < 0 || n <= 0) return r;
        const uint8_t *p = P.reflect + cell_idx(Hp, c, yy);
        r.refl = n == 2 ? (unsigned)*reinterpret_cast<const unsigned short *>(p) : (unsigned)p[0];
        r.word = P.solid[(size_t)(c + G) * (Hp >> 5) + (yy >> 5)];
        return r;
    };
    auto decode = [](RawFlags r, int yy, int v) -> unsigned {  // reflect bits | solid << 8 of row yy + v
        return ((r.refl >> (8 * v)) & 0xffu) | (((r.word >> ((yy + v) & 31)) & 1u) << 8);
    };

    // g of the row pair of column c after streaming + bounce-back, straight from the stages
    // qm / q0 / qp: the thread's first row in the stages of columns c-1 / c / c+1.  The column loop keeps these
    // pointers and rotates them (one address computation per column instead of one per use: the slot arithmetic was
    // ~70 of the 757 warp instructions per column, ncu r2b)
    auto stage_row = [&](int c) -> const T * { return gst + slot(c) * FAM + jb; };
    constexpr int FOFF = NS * FAM;  // the f ring lies FOFF elements behind the g ring, same slots
    auto pull_pair = [&](const T *qm, const T *q0, const T *qp, unsigned b0bits, unsigned b1bits, bool anyb, p2 g[9]) {
        if (!anyb) {
            g[1] = *reinterpret_cast<const p2 *>(qm + 1 * PT);
            g[3] = *reinterpret_cast<const p2 *>(qp + 3 * PT);
            g[2] = mk(q0[2 * PT - 1], q0[2 * PT]);
            g[4] = mk(q0[4 * PT + 1], q0[4 * PT + 2]);
            g[5] = mk(qm[5 * PT - 1], qm[5 * PT]);
            g[6] = mk(qp[6 * PT - 1], qp[6 * PT]);
            g[7] = mk(qp[7 * PT + 1], qp[7 * PT + 2]);
            g[8] = mk(qm[8 * PT + 1], qm[8 * PT + 2]);
        } else {
            const T *o = q0, *o1 = q0 + 1;  // the cell itself: bounced-back directions read its opposite population
            g[1] = mk(lds_pick(qm + 1 * PT, o + 3 * PT, b0bits & 0x01u), lds_pick(qm + 1 * PT + 1, o1 + 3 * PT, b1bits & 0x01u));
            g[2] = mk(lds_pick(q0 + 2 * PT - 1, o + 4 * PT, b0bits & 0x02u), lds_pick(q0 + 2 * PT, o1 + 4 * PT, b1bits & 0x02u));
            g[3] = mk(lds_pick(qp + 3 * PT, o + 1 * PT, b0bits & 0x04u), lds_pick(qp + 3 * PT + 1, o1 + 1 * PT, b1bits & 0x04u));
            g[4] = mk(lds_pick(q0 + 4 * PT + 1, o + 2 * PT, b0bits & 0x08u), lds_pick(q0 + 4 * PT + 2, o1 + 2 * PT, b1bits & 0x08u));
            g[5] = mk(lds_pick(qm + 5 * PT - 1, o + 7 * PT, b0bits & 0x10u), lds_pick(qm + 5 * PT, o1 + 7 * PT, b1bits & 0x10u));
            g[6] = mk(lds_pick(qp + 6 * PT - 1, o + 8 * PT, b0bits & 0x20u), lds_pick(qp + 6 * PT, o1 + 8 * PT, b1bits & 0x20u));
            g[7] = mk(lds_pick(qp + 7 * PT + 1, o + 5 * PT, b0bits & 0x40u), lds_pick(qp + 7 * PT + 2, o1 + 5 * PT, b1bits & 0x40u));
            g[8] = mk(lds_pick(qm + 8 * PT + 1, o + 6 * PT, b0bits & 0x80u), lds_pick(qm + 8 * PT + 2, o1 + 6 * PT, b1bits & 0x80u));
        }
        g[0] = lds_v2(q0);  // last: a predicated load is never the tail of its block (see lds_v2)
    };
    // psi_new of column c on the row pair (q) and on its outer neighbours (q_lo, q_hi); every column a fast CTA
    // touches is in the domain and carries no Zou-He rule.  KEEP = false: the pulled g is dropped -- the collision
    // of column c reloads it from the stages one iteration later (pull_pair_g), which is cheaper than 18
    // registers held across the collision of column c-1.
    auto psi_column = [&](int c, const T *sm_, const T *s0_, const T *sp_, unsigned fl_e, const unsigned fl[2], p2 g[9], p2 &q,
                          T &q_lo, T &q_hi) {  // sm_ / s0_ / sp_ = stage_row(c-1 / c / c+1)
        q = mk(0.0f, 0.0f);
        const unsigned b0bits = fl[0] & 0xffu, b1bits = fl[1] & 0xffu;
        const bool anyb = __any_sync(FULL, has && ((b0bits | b1bits) != 0u));
        if (has) {
            pull_pair(sm_, s0_, sp_, b0bits, b1bits, anyb, g);
            p2 s = add(add(add(g[0], g[1]), add(g[2], g[3])), add(add(g[4], g[5]), add(g[6], g[7])));
            s = add(s, g[8]);
            q.x = (fl[0] & 0x100u) ? P.psi_wall : s.x;
            q.y = (fl[1] & 0x100u) ? P.psi_wall : s.y;
        }
        T e = 0.0f, e2 = 0.0f;
        if (edge) {
            auto one_row = [&](int dj, unsigned fe) -> T {  // dj = stage row - jb
                const T *qm = sm_ + dj, *q0 = s0_ + dj, *qp = sp_ + dj;
                const unsigned bb = fe & 0xffu;
                T h[9];
                if (!bb) {
                    h[1] = qm[1 * PT], h[2] = q0[2 * PT - 1], h[3] = qp[3 * PT], h[4] = q0[4 * PT + 1];
                    h[5] = qm[5 * PT - 1], h[6] = qp[6 * PT - 1], h[7] = qp[7 * PT + 1], h[8] = qm[8 * PT + 1];
                } else {
                    h[1] = lds_pick(qm + 1 * PT, q0 + 3 * PT, bb & 0x01u);
                    h[2] = lds_pick(q0 + 2 * PT - 1, q0 + 4 * PT, bb & 0x02u);
                    h[3] = lds_pick(qp + 3 * PT, q0 + 1 * PT, bb & 0x04u);
                    h[4] = lds_pick(q0 + 4 * PT + 1, q0 + 2 * PT, bb & 0x08u);
                    h[5] = lds_pick(qm + 5 * PT - 1, q0 + 7 * PT, bb & 0x10u);
                    h[6] = lds_pick(qp + 6 * PT - 1, q0 + 8 * PT, bb & 0x20u);
                    h[7] = lds_pick(qp + 7 * PT + 1, q0 + 5 * PT, bb & 0x40u);
                    h[8] = lds_pick(qm + 8 * PT + 1, q0 + 6 * PT, bb & 0x80u);
                }
                h[0] = lds_f(q0);
                const T s = (((h[0] + h[1]) + (h[2] + h[3])) + ((h[4] + h[5]) + (h[6] + h[7]))) + h[8];
                return (fe & 0x100u) ? P.psi_wall : s;
            };
            e = e_ghost ? P.psi_wall : one_row(edge_lo ? -1 : 2, fl_e);
            if (edge2) e2 = ye_hi < 0 ? P.psi_wall : one_row(2, decode(load_flags(c, ye_hi, 1), ye_hi, 0));
        }
        const T dn = __shfl_up_sync(FULL, q.y, 1), up = __shfl_down_sync(FULL, q.x, 1);
        q_lo = edge_lo ? e : dn;
        q_hi = edge_hi ? (edge_lo ? e2 : e) : up;
    };

    p2 g_cur[9];
    p2 pm, p0, pp;                                   // psi_new of the row pair on columns x-1, x, x+1
    T pm_lo, pm_hi, p0_lo, p0_hi, pp_lo, pp_hi;      // ... and of the rows below / above the pair
    unsigned fl_cur[2], fl_nxt[2];
    const RawFlags z{0u, 0u};
    RawFlags fq0 = z, fq1 = z, eq0 = z, eq1 = z;     // look-ahead queues (columns x+1, x+2): own pair / outer row
    const int yef = e_ghost ? 0 : ye;                // row 0 stands in for a ghost row (flags masked at decode)
    const unsigned e_mask = e_ghost ? 0u : 0x1ffu;

    // ---- warm-up: g columns xs-2 .. xs+1, psi of columns xs-1 and xs ------------------------------------
    for (int v = xs - 4 - D; v < xs - 1; ++v) prefetch(v);
    {
        const RawFlags rf_m1 = load_flags(xs - 1, yb, nv), re_m1 = edge ? load_flags(xs - 1, yef, 1) : z;
        const RawFlags rf_0 = load_flags(xs, yb, nv), re_0 = edge ? load_flags(xs, yef, 1) : z;
        fq0 = load_flags(xs + 1, yb, nv);
        fq1 = load_flags(xs + 2, yb, nv);
        if (edge) {
            eq0 = load_flags(xs + 1, yef, 1);
            eq1 = load_flags(xs + 2, yef, 1);
        }
        cp_async_wait<D>();
        landed(xs - 2);
        landed(xs - 1);
        landed(xs);
        __syncthreads();
        fl_nxt[0] = decode(rf_m1, yb, 0), fl_nxt[1] = decode(rf_m1, yb, 1);
        psi_column(xs - 1, stage_row(xs - 2), stage_row(xs - 1), stage_row(xs), decode(re_m1, yef, 0) & e_mask, fl_nxt, g_cur, pm,
                   pm_lo, pm_hi);
        __syncthreads();
        prefetch(xs - 1);
        cp_async_wait<D>();
        landed(xs + 1);
        __syncthreads();
        fl_cur[0] = decode(rf_0, yb, 0), fl_cur[1] = decode(rf_0, yb, 1);
        psi_column(xs, stage_row(xs - 1), stage_row(xs), stage_row(xs + 1), decode(re_0, yef, 0) & e_mask, fl_cur, g_cur, p0, p0_lo,
                   p0_hi);
    }
    const T *a_m = stage_row(xs - 1), *a_0 = stage_row(xs), *a_p = stage_row(xs + 1), *a_pp = stage_row(xs + 2);  // columns x-1 .. x+2

    // ---- running pointers: everything the iteration touches is pointer + immediate -----------------------
    T *pd = P.dst + lat_idx(Hp, xs, 0, 0) + yb;
    // flags of column x+3: own pair / this lane's outer row
    const uint8_t *fr_own = P.reflect + cell_idx(Hp, xs + 3, 0) + yb;
    const uint32_t *fs_own = P.solid + (size_t)(xs + 3 + G) * (Hp >> 5) + (yb >> 5);
    const uint8_t *fr_edge = P.reflect + cell_idx(Hp, xs + 3, 0) + yef;
    const uint32_t *fs_edge = P.solid + (size_t)(xs + 3 + G) * (Hp >> 5) + (yef >> 5);

    for (int x = xs; x < xe; ++x) {
        cp_async_wait<D - 1>();  // g column x+2 has landed
        landed(x + 2);
        __syncthreads();         // ... for every thread; and everybody is done with iteration x-1
        // decode the flags of column x+1 before any new global load is issued (see k_fused)
        unsigned fe_nxt = decode(eq0, yef, 0) & e_mask;
        fl_nxt[0] = decode(fq0, yb, 0), fl_nxt[1] = decode(fq0, yb, 1);
        asm volatile("" : "+r"(fl_nxt[0]), "+r"(fl_nxt[1]), "+r"(fe_nxt)::"memory");
        p2 f[9];
        {   // f of column x: stream + bounce-back straight into registers
            const unsigned b0bits = fl_cur[0] & 0xffu, b1bits = fl_cur[1] & 0xffu;
            const bool anyb = __any_sync(FULL, has && ((b0bits | b1bits) != 0u));
            if (has) pull_pair(a_m + FOFF, a_0 + FOFF, a_p + FOFF, b0bits, b1bits, anyb, f);
        }
        prefetch(x);  // after the f loads: f is what the iteration waits for first (+0.9 %)
        // flags of column x+3 (decoded two iterations from now); the flag arrays carry one spare column
        RawFlags fq2 = z, eq2 = z;
        if (has) {
            fq2.refl = *reinterpret_cast<const unsigned short *>(fr_own);
            fq2.word = *fs_own;
        }
        if (edge) {
            eq2.refl = *fr_edge;
            eq2.word = *fs_edge;
        }
        {
            p2 gd[9];  // dropped (see psi_column)
            psi_column(x + 1, a_0, a_p, a_pp, fe_nxt, fl_nxt, gd, pp, pp_lo, pp_hi);
        }
        if (has) {
            const bool s0 = fl_cur[0] & 0x100u, s1 = fl_cur[1] & 0x100u;
            if (!(s0 && s1)) {
                T gxa, gya, lapa, gxb, gyb, lapb;
                //        C     E     W     N      S      NE     NW     SW     SE
                stencil9(p0.x, pp.x, pm.x, p0.y, p0_lo, pp.y, pm.y, pm_lo, pp_lo, gxa, gya, lapa);
                stencil9(p0.y, pp.y, pm.y, p0_hi, p0.x, pp_hi, pm_hi, pm.x, pp.x, gxb, gyb, lapb);
                Macro2 m;
                moments2(P, f, p0, mk(gxa, gxb), mk(gya, gyb), mk(lapa, lapb), s0, s1, m);
                collide2(P, m, s0, s1, f, g_cur);
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                *reinterpret_cast<p2 *>(pd + (ptrdiff_t)i * Hp) = f[i];
                *reinterpret_cast<p2 *>(pd + (ptrdiff_t)(9 + i) * Hp) = g_cur[i];
            }
            // halo push: the two edge columns also land in the neighbours' ghost columns (peer stores over NVLink).  It
            // stays in this loop (the fp64 kernel leaves it to its face CTAs): the face CTAs here run the SCALAR code, whose
            // rounding order differs from the packed collision, and N slabs must reproduce one slab bit for bit
            if (P.peer_lo && x < G) {
                T *o = P.peer_lo + lat_idx(Hp, P.peer_lo_Wl + x, 0, yb);
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    *reinterpret_cast<p2 *>(o + (ptrdiff_t)i * Hp) = f[i];
                    *reinterpret_cast<p2 *>(o + (ptrdiff_t)(9 + i) * Hp) = g_cur[i];
                }
            }
            if (P.peer_hi && x >= P.Wl - G) {
                T *o = P.peer_hi + lat_idx(Hp, x - P.Wl, 0, yb);
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    *reinterpret_cast<p2 *>(o + (ptrdiff_t)i * Hp) = f[i];
                    *reinterpret_cast<p2 *>(o + (ptrdiff_t)(9 + i) * Hp) = g_cur[i];
                }
            }
        }
        {   // g of column x+1 for the next iteration's collision (the stages x .. x+2 are still in place)
            const unsigned b0bits = fl_nxt[0] & 0xffu, b1bits = fl_nxt[1] & 0xffu;
            const bool anyb = __any_sync(FULL, has && ((b0bits | b1bits) != 0u));
            if (has) pull_pair(a_0, a_p, a_pp, b0bits, b1bits, anyb, g_cur);
        }
        {   // four stages: column x+3 takes the slot of column x-1, the pointers only rotate
            const T *a_n = a_m;
            a_m = a_0, a_0 = a_p, a_p = a_pp, a_pp = a_n;
        }
        pm = p0, p0 = pp;
        pm_lo = p0_lo, pm_hi = p0_hi, p0_lo = pp_lo, p0_hi = pp_hi;
        fl_cur[0] = fl_nxt[0], fl_cur[1] = fl_nxt[1];
        fq0 = fq1, fq1 = fq2;
        eq0 = eq1, eq1 = eq2;
        pd += S;
        fr_own += Hp, fr_edge += Hp;
        fs_own += Hp >> 5, fs_edge += Hp >> 5;
    }
    cp_async_wait<0>();
}

template <int HPC>
__global__ void __launch_bounds__(Cfg::NT, 3)
    k_fused_f32p(const __grid_constant__ LbmParams<float> P, int nyt, int chunk, int fx0, int fx1, int n_fast)
{
    constexpr int ROWS = Cfg::ROWS, NT = Cfg::NT, HALO = Cfg::HALO;
    static_assert(VecCfg<float, NT, 2>::SMEM <= Cfg::SMEM && VecCfg<float, NT, 2>::ROWS == ROWS, "same strips as k_fused_vec");
    if ((int)blockIdx.x >= n_fast) {  // face CTA
        const int k = (int)blockIdx.x - n_fast, side = k / nyt;
        const bool left = fx0 > 0 && side == 0;
        fused_vec_strip<float, NT, 2>(P, k % nyt, left ? 0 : fx1, left ? fx0 : P.Wl);
        return;
    }
    const int y0 = (int)(blockIdx.x % nyt) * ROWS, ny = min(ROWS, P.H - y0);
    const bool wrap = y0 - HALO < 0 || y0 + ny + HALO > P.H;
    if (FDLBM_F32_BULK && (!wrap || ((P.H % 4) == 0 && (ny % 4) == 0)) && (ny % 64) != 2)
        fast_strip<HPC, true>(P, nyt, chunk, fx0, fx1);
    else
        fast_strip<HPC, false>(P, nyt, chunk, fx0, fx1);
}

// the plain column range of this slab: columns x with x and x+1 in the domain and away from the Zou-He faces
// (the faces, their neighbours -- psi_new is stored there for the next step's Zou-He -- and nothing else are left
// to the face CTAs)
inline void plain_range(const LbmParams<float> &P, int &fx0, int &fx1)
{
    fx0 = 0, fx1 = P.Wl;
    if (!P.x_periodic) {
        fx0 = 2 - P.gx0 > 0 ? 2 - P.gx0 : 0;
        fx1 = P.W - 3 - P.gx0 < P.Wl ? P.W - 3 - P.gx0 : P.Wl;
    }
}
inline bool applicable(const LbmParams<float> &P)
{
    int fx0, fx1;
    plain_range(P, fx0, fx1);
    return P.H % 2 == 0 && fx1 - fx0 >= 16;
}

template <int HPC>
int launch_hp(const LbmParams<float> &P, cudaStream_t stream)
{
    auto kern = k_fused_f32p<HPC>;
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
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::NT, Cfg::SMEM);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) occ = 1;
        n_cta = sms * occ;
    }
    int fx0, fx1;
    plain_range(P, fx0, fx1);
    const int nyt = (P.H + Cfg::ROWS - 1) / Cfg::ROWS;
    const int n_face = nyt * ((fx0 > 0) + (fx1 < P.Wl));
    // one wave of plain-column CTAs; keep a slot free for the face CTAs to slip into
    const int slots = n_face && n_cta > nyt ? n_cta - 1 : n_cta;
    const int chunk = fused_chunk(nyt, slots, fx1 - fx0);
    const int nchunks = (fx1 - fx0 + chunk - 1) / chunk;
    kern<<<nyt * nchunks + n_face, Cfg::NT, Cfg::SMEM, stream>>>(P, nyt, chunk, fx0, fx1, nyt * nchunks);
    return 0;
}

// even grid heights only; the common row pitches get their own instantiation
inline int launch(const LbmParams<float> &P, cudaStream_t stream)
{
    switch (P.Hp) {
    case 2048: return launch_hp<2048>(P, stream);
    case 4096: return launch_hp<4096>(P, stream);
    case 8192: return launch_hp<8192>(P, stream);
    default: return launch_hp<0>(P, stream);
    }
}

#undef FDLBM_DI
}  // namespace f32p
}  // namespace fdlbm
