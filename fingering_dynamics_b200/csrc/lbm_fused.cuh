// lbm_fused.cuh -- the one-pass step kernel: stream + bounce-back + Zou-He + moments + interaction
// force (psi stencils) + BGK collision, 18 populations read once and written once per lattice update.
//
// The collision of cell x needs grad/lap of the NEW psi at x, i.e. psi_new on the 8 neighbours, each a
// sum of g pulled from ITS neighbours (two-ring dependency, SURVEY.md finding 5).  Instead of staging
// psi_new through HBM (+30% traffic), a CTA owns a strip of TY consecutive y and MARCHES along x:
//
//   * the g AND f columns stream from HBM into two shared-memory stage rings (4 stages each, one column ahead of the
//     compute front) through the bulk-copy engine: one thread issues 18 cp.async.bulk of a whole stage row each,
//     completion is counted in bytes on the stage's mbarrier; the loads of the next column are in flight while the
//     current one is collided, and the strip's neighbours in y are reachable for the psi halo (apron rows);
//   * iteration x:  wait for column x+2's stage (mbarrier) and the one __syncthreads -> issue the fill of column x+3
//     -> pull f of column x and g of column x+1 from the stages (bounce-back = address select) -> psi_new(x+1, y);
//     rows y-1 / y+1 arrive by WARP SHUFFLE from the neighbouring lanes (the two end lanes of a warp evaluate their
//     outer neighbour themselves), the 3x3 psi neighbourhood lives in registers and rotates with x -> moments of
//     column x, stencils, collide with the g pulled one iteration earlier, store the 18 populations (coalesced);
//   * two bodies: PLAIN (every column the strip touches is inside the domain, no Zou-He rule, no halo push, no
//     one-row warp; row pitch a template parameter for H = 2048 / 4096 / 8192) for almost all columns, the general
//     one for the 2-3 columns next to a face / slab edge, which get their own small CTAs at the end of the grid.
//
// Scheduling: one resident wave.  Strips are the fast CTA index, so the CTAs working on the same column
// range advance in step and the apron rows a strip reads are L2 hits on lines its neighbour streams.
// The y wrap is resolved when a stage is filled; the x wrap / slab halo comes from the ghost columns.
// Measured on B200, 8192x2048 fp64: 22.2 GLUPS = 0.99 of the HBM roofline (round 1: 0.89); profiles/README.md.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "lbm_device.cuh"

namespace fdlbm {

#ifndef FDLBM_FUSED_TY
#define FDLBM_FUSED_TY 128
#endif
#ifndef FDLBM_FUSED_D
#define FDLBM_FUSED_D 1  // NS = 3 + D = 4 stages: a power of two, the slot of a column is c & 3
#endif
// resident CTAs per SM the register allocation is bounded for: fp64 needs ~166 registers to stay free of
// spills (3 CTAs = 12 warps, 76 KB of stages each); the generic fp32 instantiation fits 128 registers (4 CTAs)
#ifndef FDLBM_FUSED_MINB64
#define FDLBM_FUSED_MINB64 3
#endif
#ifndef FDLBM_FUSED_MINB32
#define FDLBM_FUSED_MINB32 4
#endif
#ifndef FDLBM_BULK_COPY
#define FDLBM_BULK_COPY 1  // g stages of interior strips through cp.async.bulk + mbarrier (0: per-thread cp.async only)
#endif
constexpr int FUSED_TY = FDLBM_FUSED_TY;  // rows per strip = threads per CTA
constexpr int FUSED_D = FDLBM_FUSED_D;    // cp.async prefetch distance in columns
#ifndef FDLBM_L2_AHEAD
#define FDLBM_L2_AHEAD 2
#endif
constexpr int FUSED_L2_AHEAD = FDLBM_L2_AHEAD;  // L2 prefetch distance of the f columns (0 = off)

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
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- bulk asynchronous copies (the TMA engine's 1-D path, cp.async.bulk) completing on an mbarrier ----------
#ifndef FDLBM_MIN_CHUNK
#define FDLBM_MIN_CHUNK 2  // shortest column chunk of a CTA: small grids are bound by the per-column chain x columns per CTA, so they get many short CTAs (400x400: 22.5 -> 18.4 us per step, 200x250: 29.8 -> 20.5)
#endif
#ifndef FDLBM_FUSED_PLAIN
#define FDLBM_FUSED_PLAIN 1  // plain columns run a body without the domain / Zou-He tests, the faces get their own CTAs; 0 for A/B
#endif
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <typename T, int TY>
struct FusedCfg {
    static constexpr int HALO = 16 / (int)sizeof(T);      // rows of apron per side: keeps 16-byte chunks aligned
    static constexpr int PT = TY + 2 * HALO;              // stage row pitch (elements)
    static constexpr int NS = 3 + FUSED_D;                // stages of the g ring (columns x..x+2+D)
    static constexpr int FAM = 9 * PT;                    // elements of one stage (one family of one column)
    static constexpr int RINGS = 2;                       // g ring + f ring
    static constexpr size_t SMEM = (size_t)(RINGS * NS * FAM) * sizeof(T);
};

// Fill one stage: 9 populations of one family of the column whose record starts at `col` (pop 0 of the
// family), rows [y0-HALO, y0+rows+HALO) with the y wrap applied.  16-byte cp.async where the chunk is
// contiguous in memory, element-wise at the wrap.
template <typename T, int TY, int PT, int HALO>
__device__ __forceinline__ void stage_fill(T *stage, const T *col, int Hp, int H, int y0, int rows)
{
    constexpr int EPC = 16 / (int)sizeof(T);  // elements per chunk
    const int nch = (rows + 2 * HALO + EPC - 1) / EPC;
    for (int ch = threadIdx.x; ch < nch; ch += TY) {
        const int r0 = ch * EPC;
        const int y = y0 - HALO + r0;
        if (y >= 0 && y + EPC <= H) {
#pragma unroll
            for (int pop = 0; pop < 9; ++pop) cp_async16(stage + pop * PT + r0, col + (size_t)pop * Hp + y);
        } else {
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
                int yy = (y + e) % H;
                if (yy < 0) yy += H;
#pragma unroll
                for (int pop = 0; pop < 9; ++pop)
                    cp_async_small<(int)sizeof(T)>(stage + pop * PT + r0 + e, col + (size_t)pop * Hp + yy);
            }
        }
    }
}

// pull-stream + bounce-back from three stages: v_i <- stage_i(column c - e_x)[row j - e_y]
template <typename T, int PT>
__device__ __forceinline__ void pull_staged(const T *sm, const T *s0, const T *sp, int j, unsigned bits, T v[9])
{
    const T *s1 = sm + 1 * PT + j, *s2 = s0 + 2 * PT + j - 1, *s3 = sp + 3 * PT + j, *s4 = s0 + 4 * PT + j + 1;
    const T *s5 = sm + 5 * PT + j - 1, *s6 = sp + 6 * PT + j - 1, *s7 = sp + 7 * PT + j + 1, *s8 = sm + 8 * PT + j + 1;
    if (bits) {  // bounced-back directions: select the address (one load per register)
        const T *o = s0 + j;
        if (bits & 0x01u) s1 = o + 3 * PT;
        if (bits & 0x02u) s2 = o + 4 * PT;
        if (bits & 0x04u) s3 = o + 1 * PT;
        if (bits & 0x08u) s4 = o + 2 * PT;
        if (bits & 0x10u) s5 = o + 7 * PT;
        if (bits & 0x20u) s6 = o + 8 * PT;
        if (bits & 0x40u) s7 = o + 5 * PT;
        if (bits & 0x80u) s8 = o + 6 * PT;
    }
    v[0] = s0[j];
    v[1] = *s1;
    v[2] = *s2;
    v[3] = *s3;
    v[4] = *s4;
    v[5] = *s5;
    v[6] = *s6;
    v[7] = *s7;
    v[8] = *s8;
}

struct RawFlags {
    unsigned refl, word;  // reflect byte and solid-mask word exactly as loaded
};

// HPC > 0: the row pitch Hp is the compile-time constant HPC (all population offsets become immediates of the
// loads / stores / cp.async); HPC == 0: Hp is read from the parameters (any grid height).
// PLAIN: every column the strip touches (xs-1 .. xe+1) is inside the domain and carries no Zou-He rule, so the
// per-column domain / face tests (and the loads of W, gx0, psi_left, psi_right they need: 24 LDC per warp and column
// in the one-body kernel) are compiled out; the face columns run the general body in their own small CTAs.
// all stages of the strip starting at row y0 can be filled by bulk copies: no y wrap inside the apron, or every piece a
// 16-byte multiple
template <typename T, int TY>
__device__ __forceinline__ bool fused_bulk_strip(int H, int y0)
{
    constexpr int HALO = FusedCfg<T, TY>::HALO;
    const int ny = min(TY, H - y0);
    const bool wrap = y0 - HALO < 0 || y0 + TY + HALO > H;
    return FDLBM_BULK_COPY && (!wrap || ((H * sizeof(T)) % 16 == 0 && (ny * sizeof(T)) % 16 == 0));
}

template <typename T, int TY, int HPC, bool PLAIN>
__device__ __forceinline__ void fused_strip(const LbmParams<T> &P, const int yt, const int xs, const int xe)
{
    using C = FusedCfg<T, TY>;
    constexpr int D = FUSED_D, NS = C::NS, PT = C::PT, HALO = C::HALO, FAM = C::FAM;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *gst = reinterpret_cast<T *>(smem_raw);        // [NS][9][PT]
    T *fst = gst + NS * FAM;                         // [NS][9][PT]
    __shared__ __align__(8) unsigned long long bars[NS];  // one mbarrier per g stage (bulk-copy path)
    const int t = threadIdx.x, lane = t & 31;
    const int H = P.H, Hp = HPC > 0 ? HPC : P.Hp;

    // strips are the fast CTA index: the CTAs of a column chunk start together and advance in step.  The chunks
    // are EQUAL and the hardware places the CTAs: evening out the CTA / SM finishing times (the launch ends with
    // its slowest SM, ~5 % after the median one) by measured chunk lengths or by placement was tried three ways in
    // round 2 and lost 2-5 % every time (profiles/README.md, profiles/r2/tail_experiments/): anything that pulls the
    // strips of a chunk out of lock step costs more in L2 / DRAM locality than the tail it removes.
    const int y0 = yt * TY;
    const int y = y0 + t;
    const int ny = min(TY, H - y0);
    const bool active = t < ny;
    const int j = t + HALO;  // stage row of this thread's cell

    // psi of the rows y-1 / y+1 comes from the neighbouring lanes by warp shuffle.  The first and last
    // active lane of a warp have no such neighbour: they evaluate psi of that row themselves (same
    // staged data, one extra pull for two lanes per warp), so warps never wait for each other.
    const bool edge_lo = active && lane == 0;
    const bool edge_hi = active && (lane == 31 || t == ny - 1);
    // A lane that is both (a warp of ONE active row: H % 32 == 1 only) needs a second neighbour row, whose flags it
    // loads in the loop.  The PLAIN body is only launched when no warp is like that and has no such load: ptxas put it
    // on the scoreboard of the look-ahead flag loads and then made the psi shuffles -- which reuse its register --
    // wait for that scoreboard on EVERY path: a full memory latency per column (ncu r2c: 14 % of all stall samples;
    // without it +4.1 %, profiles/r2/r2_ab_edge2.txt; profiles/tools/sass_scoreboards.py checks the SASS for the pattern).
    const bool edge = edge_lo || edge_hi, edge2 = !PLAIN && edge_lo && edge_hi;
    auto wrap_row = [&](int yy) {  // wrapped global row, or -1 for a ghost row of a y-wall variant
        if (yy < 0 || yy >= H) return P.y_wall ? -1 : (yy < 0 ? yy + H : yy - H);
        return yy;
    };
    const int ye1 = wrap_row(edge_lo ? y - 1 : y + 1), je1 = edge_lo ? j - 1 : j + 1;
    const int ye2 = wrap_row(y + 1), je2 = j + 1;  // second neighbour of a one-row warp (edge2)

    // the stage of column c, counted from the strip's first column
    auto slot = [xs](int c) { return (NS & (NS - 1)) == 0 ? ((c - xs) & (NS - 1)) : (((c - xs) % NS) + NS) % NS; };
    auto slot_add = [](int sc, int d) { return (NS & (NS - 1)) == 0 ? ((sc + d) & (NS - 1)) : (((sc + d) % NS) + NS) % NS; };
    auto in_domain = [&](int c) { return PLAIN || P.x_periodic || (P.gx0 + c >= 0 && P.gx0 + c < P.W); };

    // Interior strips (no y wrap inside the apron) fill their g stages with BULK asynchronous copies: one thread
    // issues 9 cp.async.bulk of a whole stage row each (1056 bytes), completion is counted in bytes on the stage's
    // mbarrier.  Strips that touch the wrap use per-thread 16-byte cp.async (stage_fill).
    // Strips whose apron wraps in y copy the contiguous piece in bulk and the wrapped apron with 16-byte cp.async when
    // the pieces keep the 16-byte granularity; the kernel ends with its slowest CTA, and strips on the per-thread path
    // were that CTA.
    const bool wrap_lo = y0 - HALO < 0, wrap_hi = y0 + TY + HALO > H;  // CTA-uniform
    // the PLAIN body is only launched for strips whose stages all fill in bulk (fused_bulk_strip): the per-thread fill path is
    // then compiled out of its column loop
    const bool bulk = PLAIN || fused_bulk_strip<T, TY>(H, y0);
    if (t == 0) {
#pragma unroll
        for (int s_ = 0; s_ < NS; ++s_) mbar_init(&bars[s_], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // one pipeline step: g column v+2+D (and f column v+1+D: f is consumed one column behind g, so
    // its ring holds x-1..x+1 plus the column in flight); only columns this run reads.  Whenever an f column is
    // fetched a g column is fetched with it, on the same mbarrier / commit group.
    auto prefetch_s = [&](int v, const int scg) {  // scg = slot(v + 2 + D)
        const int cg = v + 2 + D;
        if (cg >= xs - 2 && cg <= xe + 1) {
            T *stage = gst + scg * FAM;
            const T *col = P.src + lat_idx(Hp, cg, 9, 0);
            const bool with_f = cg - 1 >= xs - 1 && cg - 1 <= xe;
            T *fstage = fst + slot_add(scg, -1) * FAM;
            const T *fcol = P.src + lat_idx(Hp, cg - 1, 0, 0);
            if (bulk) {
                // one bulk copy per population for the contiguous piece of the stage row ...
                if (t == 0) {
                    const int r0 = wrap_lo ? y0 : y0 - HALO;               // its first row
                    const int r1 = wrap_hi ? y0 + ny : y0 + ny + HALO;     // one past its last row
                    const unsigned bytes = (unsigned)((r1 - r0) * sizeof(T));
                    mbar_expect_tx(&bars[scg], (with_f ? 18u : 9u) * bytes);
#pragma unroll
                    for (int pop = 0; pop < 9; ++pop)
                        bulk_g2s(stage + pop * PT + (r0 - (y0 - HALO)), col + (size_t)pop * Hp + r0, bytes, &bars[scg]);
                    if (with_f) {
#pragma unroll
                        for (int pop = 0; pop < 9; ++pop)
                            bulk_g2s(fstage + pop * PT + (r0 - (y0 - HALO)), fcol + (size_t)pop * Hp + r0, bytes, &bars[scg]);
                    }
                }
                // ... and one 16-byte cp.async per population for an apron that wraps in y (16-byte bulk copies
                // measured far slower: 17.0 instead of 19.5 GLUPS)
                if (wrap_lo && t >= 32 && t < 41)
                    cp_async16(stage + (t - 32) * PT, col + (size_t)(t - 32) * Hp + (y0 - HALO + H));
                if (wrap_hi && t >= 64 && t < 73)
                    cp_async16(stage + (t - 64) * PT + HALO + ny, col + (size_t)(t - 64) * Hp + (y0 + ny - H));
                if (with_f) {
                    if (wrap_lo && t >= 41 && t < 50)
                        cp_async16(fstage + (t - 41) * PT, fcol + (size_t)(t - 41) * Hp + (y0 - HALO + H));
                    if (wrap_hi && t >= 73 && t < 82)
                        cp_async16(fstage + (t - 73) * PT + HALO + ny, fcol + (size_t)(t - 73) * Hp + (y0 + ny - H));
                }
            } else {
                stage_fill<T, TY, PT, HALO>(stage, col, Hp, H, y0, ny);
                if (with_f) stage_fill<T, TY, PT, HALO>(fstage, fcol, Hp, H, y0, ny);
            }
        }
        cp_async_commit();
    };
    auto prefetch = [&](int v) { prefetch_s(v, slot(v + 2 + D)); };
    // wait until g column c has landed in its stage (bulk path: the fill of column c is the
    // ((c - (xs-2)) / NS)-th use of its mbarrier, whose parity is waited for)
    auto landed_s = [&](int c, const int sc) {
        if (bulk) mbar_wait(&bars[sc], (unsigned)(((c - (xs - 2)) / NS) & 1));
    };
    auto landed = [&](int c) { landed_s(c, slot(c)); };
    // Per-cell flags are plain global loads issued two columns ahead of their use and carried RAW in
    // registers: nothing depends on them until they are decoded two iterations later, so their latency
    // never sits on the critical path.
    auto load_flags = [&](int c, int yy) -> RawFlags {
        RawFlags r{0u, 0u};
        if (c > xe + 1 || yy < 0 || !in_domain(c)) return r;
        r.refl = P.reflect[cell_idx(Hp, c, yy)];
        r.word = P.solid[(size_t)(c + G) * (Hp >> 5) + (yy >> 5)];
        return r;
    };
    auto decode = [](RawFlags r, int yy) -> unsigned {  // reflect bits | solid << 8
        return r.refl | (((r.word >> (yy & 31)) & 1u) << 8);
    };
    // psi_new of the cell (column c, global row yy >= 0 already wrapped, stage row jj) from the g stages
    // bm / b0 / bp: the g stages of columns c-1 / c / c+1.  The column loop keeps the four stage pointers of columns
    // x-1 .. x+2 and rotates them instead of recomputing slot * FAM at every use
    auto stage_of = [&](int c) -> const T * { return gst + slot(c) * FAM; };
    constexpr int FOFF = NS * FAM;  // the f ring lies FOFF elements behind the g ring, same slots
    auto psi_staged = [&](int c, const T *bm, const T *b0, const T *bp, int yy, int jj, unsigned flags, T g[9]) -> T {
        if (yy < 0) return P.psi_wall;
        if (!PLAIN) {
            const int gx = P.gx0 + c;
            if (!P.x_periodic) {
                if (gx < 0) return P.psi_left;
                if (gx >= P.W) return P.psi_right;
            }
        }
        pull_staged<T, PT>(bm, b0, bp, jj, flags & 0xffu, g);
        if (!PLAIN && P.zou_he) {
            const int gx = P.gx0 + c;
            if (gx == 0 || gx == P.W - 1) zou_he_g(P, gx, yy, g);
        }
        if (flags & 0x100u) return P.psi_wall;
        return (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) + g[8];
    };
    // psi_new of column c on rows y-1, y, y+1 (q_m, q_0, q_p); the pulled g of the own cell is returned
    auto psi_column = [&](int c, const T *bm, const T *b0, const T *bp, unsigned fl_own, unsigned fl_edge, T g[9], T &q_m,
                          T &q_0, T &q_p) {
        q_0 = T(0);
        if (active) q_0 = psi_staged(c, bm, b0, bp, y, j, fl_own, g);
        T e1 = T(0), e2 = T(0);
        if (edge) {
            T gh[9];
            e1 = psi_staged(c, bm, b0, bp, ye1, je1, fl_edge, gh);
            if (edge2) e2 = psi_staged(c, bm, b0, bp, ye2, je2, decode(load_flags(c, ye2), ye2), gh);
        }
        const T dn = __shfl_up_sync(FULL, q_0, 1), up = __shfl_down_sync(FULL, q_0, 1);
        q_m = edge_lo ? e1 : dn;
        q_p = edge_hi ? (edge_lo ? e2 : e1) : up;
    };

    T g_a[9], g_b[9];  // g of columns x and x+1 (roles alternate in the unrolled loop)
    T pm_m, pm_0, pm_p, p0_m, p0_0, p0_p, pp_m, pp_0, pp_p;  // psi_new on columns x-1, x, x+1 x rows y-1, y, y+1
    unsigned fl_cur = 0, fl_nxt = 0;                          // flags of this thread's cell in columns x, x+1
    const RawFlags z{0u, 0u};
    RawFlags fq0 = z, fq1 = z, eq0 = z, eq1 = z;               // look-ahead queues: own / edge-neighbour cell, columns x+1, x+2

    // pipeline warm-up.  Steps v = xs-4-D .. xs-2 bring in g columns xs-2 .. xs+D (all NS slots);
    // psi(xs-1) needs g columns xs-2..xs, then g column xs-2 makes room for column xs+1+D.
    for (int v = xs - 4 - D; v < xs - 1; ++v) prefetch(v);
    const RawFlags rf_m1 = active ? load_flags(xs - 1, y) : z, re_m1 = edge ? load_flags(xs - 1, ye1) : z;
    const RawFlags rf_0 = active ? load_flags(xs, y) : z, re_0 = edge ? load_flags(xs, ye1) : z;
    if (active) {
        fq0 = load_flags(xs + 1, y);
        fq1 = load_flags(xs + 2, y);
    }
    if (edge) {
        eq0 = load_flags(xs + 1, ye1);
        eq1 = load_flags(xs + 2, ye1);
    }
    cp_async_wait<D>();  // g columns <= xs have landed (this thread's share)
    landed(xs - 2);
    landed(xs - 1);
    landed(xs);
    __syncthreads();
    psi_column(xs - 1, stage_of(xs - 2), stage_of(xs - 1), stage_of(xs), decode(rf_m1, y), decode(re_m1, ye1), g_b, pm_m, pm_0, pm_p);
    __syncthreads();
    prefetch(xs - 1);
    cp_async_wait<D>();  // g column xs+1 has landed
    landed(xs + 1);
    __syncthreads();
    fl_cur = decode(rf_0, y);
    psi_column(xs, stage_of(xs - 1), stage_of(xs), stage_of(xs + 1), fl_cur, decode(re_0, ye1), g_a, p0_m, p0_0, p0_p);
    const T *b_m = stage_of(xs - 1), *b_0 = stage_of(xs), *b_p = stage_of(xs + 1), *b_pp = stage_of(xs + 2);  // columns x-1 .. x+2

    // one column.  sx = slot(x); g_cur = g of column x (in), g_nxt = g of column x+1 (out).  Unrolling this loop was
    // measured on B200 (profiles/r2/r2_ab_unroll.txt, r2_ab_lean1.txt): by 4 with every stage slot a literal -18 % (four
    // copies of the 12 KB body overflow the 32 KB instruction cache level), by 2 (no register moves for g) -2 %; running
    // pointers for the flag loads / stores instead of index arithmetic: -0.3 % (r2_ab_lean2.txt).
    auto column = [&](const int x, const int sx, T (&g_cur)[9], T (&g_nxt)[9]) {
        cp_async_wait<D - 1>();  // g column x+2 has landed
        landed_s(x + 2, slot_add(sx, 2));
        __syncthreads();         // ... for every thread; and everybody is done with iteration x-1
        // Decode the flags of column x+1 BEFORE any new global load is issued: they were loaded two iterations
        // ago, but the hardware scoreboard slots are shared -- decoding them after this iteration's f loads
        // would wait for those loads too (measured: 18 % of all stall samples on the first psi shuffle).
        fl_nxt = decode(fq0, y);
        unsigned fe_nxt = decode(eq0, ye1);
        asm volatile("" : "+r"(fl_nxt), "+r"(fe_nxt)::"memory");
        prefetch_s(x, slot_add(sx, 2 + D));  // overwrites the stage of g column x-1: no longer read
        T f[9];
        {
            // f of column x from its stages into registers; consumed after the psi phase below
            if (active) pull_staged<T, PT>(b_m + FOFF, b_0 + FOFF, b_p + FOFF, j, fl_cur & 0xffu, f);
        }
        const RawFlags fq2 = active ? load_flags(x + 3, y) : z;  // decoded two iterations from now
        const RawFlags eq2 = edge ? load_flags(x + 3, ye1) : z;
        psi_column(x + 1, b_0, b_p, b_pp, fl_nxt, fe_nxt, g_nxt, pp_m, pp_0, pp_p);
        if (active) {
            if (!PLAIN && P.zou_he) {
                const int gx_ = P.gx0 + x;
                if (gx_ == 0 || gx_ == P.W - 1) zou_he_f(P, x, gx_, y, f, PullRow<T>());
            }
            if (!(fl_cur & 0x100u)) {
                T gx, gy, lap;
                //        C     E     W     N     S     NE    NW    SW    SE
                stencil9(p0_0, pp_0, pm_0, p0_p, p0_m, pp_p, pm_p, pm_m, pp_m, gx, gy, lap);
                Macro<T> m;
                moments(P, f, p0_0, gx, gy, lap, m);
                collide(P, m, f, g_cur);
            }
            if (PLAIN) store_cell_at(P.dst, Hp, x, y, f, g_cur);  // no halo push in the plain range (fused_plain_range)
            else store_cell_hp(P, Hp, x, y, f, g_cur);
            if (!PLAIN && P.zou_he) {  // the next step's Zou-He needs grad psi and mu at the face columns
                const int gx_ = P.gx0 + x;
                if (gx_ < 2 || gx_ >= P.W - 2) P.psi_new[cell_idx(Hp, x, y)] = p0_0;
            }
        }
        pm_m = p0_m, pm_0 = p0_0, pm_p = p0_p;
        p0_m = pp_m, p0_0 = pp_0, p0_p = pp_p;
        fl_cur = fl_nxt;
        fq0 = fq1, fq1 = fq2;
        eq0 = eq1, eq1 = eq2;
        {   // column x+3 takes the stage of column x-1 (four stages: the pointers only rotate)
            const T *b_n = NS == 4 ? b_m : stage_of(x + 3);
            b_m = b_0, b_0 = b_p, b_p = b_pp, b_pp = b_n;
        }
    };
    for (int x = xs; x < xe; ++x) {
        column(x, slot(x), g_a, g_b);
#pragma unroll
        for (int i = 0; i < 9; ++i) g_a[i] = g_b[i];
    }
    cp_async_wait<0>();
}

// The launch grid (same scheme as f32p::k_fused_f32p): `n_fast` CTAs march over the plain columns [fx0, fx1) (strip =
// blockIdx % nyt, chunks of `chunk` columns) with the PLAIN body, followed by the FACE CTAs -- one per strip and side for
// the columns next to the Zou-He faces / the domain edge, [0, fx0) and [fx1, Wl) -- with the general body.  Face CTAs
// come last in dispatch order and are tiny (2-3 columns): they slip into the CTA slots the one-wave chunking leaves free.
template <typename T, int TY, int HPC>
__global__ void __launch_bounds__(TY, sizeof(T) == 8 ? FDLBM_FUSED_MINB64 : FDLBM_FUSED_MINB32)
    k_fused(const __grid_constant__ LbmParams<T> P, int nyt, int chunk, int fx0, int fx1, int n_fast, int plain_ok)
{
    int yt, xs, xe;
    const bool face = (int)blockIdx.x >= n_fast;
    if (face) {
        const int k = (int)blockIdx.x - n_fast, side = k / nyt;
        const bool left = fx0 > 0 && side == 0;
        yt = k % nyt, xs = left ? 0 : fx1, xe = left ? fx0 : P.Wl;
    } else {
        yt = (int)blockIdx.x % nyt, xs = fx0 + ((int)blockIdx.x / nyt) * chunk, xe = min(fx1, xs + chunk);
    }
    if (FDLBM_FUSED_PLAIN && plain_ok && !face && fused_bulk_strip<T, TY>(P.H, yt * TY))
        fused_strip<T, TY, HPC, true>(P, yt, xs, xe);
    else
        fused_strip<T, TY, HPC, false>(P, yt, xs, xe);
}

// the plain column range of this slab: columns x with x-1 .. x+2 inside the domain and away from the Zou-He faces and
// their neighbours (where psi_new is stored for the next step's Zou-He); everything on an x-periodic grid
template <typename T>
inline void fused_plain_range(const LbmParams<T> &P, int &fx0, int &fx1)
{
    fx0 = 0, fx1 = P.Wl;
    if (FDLBM_FUSED_PLAIN && P.H % 32 != 1 && !P.x_periodic) {
        fx0 = 2 - P.gx0 > 0 ? 2 - P.gx0 : 0;
        fx1 = P.W - 3 - P.gx0 < P.Wl ? P.W - 3 - P.gx0 : P.Wl;
    }
    if (FDLBM_FUSED_PLAIN && P.H % 32 != 1) {
        // the halo push into a neighbour's ghost columns (peer stores) is face-CTA work too
        if (P.peer_lo && fx0 < G) fx0 = G;
        if (P.peer_hi && fx1 > P.Wl - G) fx1 = P.Wl - G;
        if (fx0 > P.Wl) fx0 = P.Wl;
        if (fx1 < fx0) fx1 = fx0;
    }
}

// Column chunk length for `nyt` strips on `n_cta` resident CTA slots.  With c chunks per strip the kernel takes
// ceil(nyt*c / n_cta) waves of Wl/c columns each; pick the c that minimises waves/c (one wave when the strips
// divide the slots well, a few shorter waves for tall grids); the 3-column pipeline warm-up of every chunk is in the
// cost.  Small grids (the reference's own 400x400 / 380x380 / 200x250) get chunks as short as 2 columns: filling the
// SMs matters more than the warm-up there (400x400 fp64: 176 -> 25 us per step with 8-column chunks in round 1, 18 us now).
inline int fused_chunk(int nyt, int n_cta, int Wl)
{
    int best_c = 1;
    double best = 1e30;
    const int c_max = Wl / FDLBM_MIN_CHUNK > 1 ? Wl / FDLBM_MIN_CHUNK : 1;
    for (int c = 1; c <= c_max && c <= 4096; ++c) {
        const int waves = (nyt * c + n_cta - 1) / n_cta;
        const double cost = (double)waves / c * (1.0 + 3.0 * c / Wl);
        if (cost < best * 0.97) {  // prefer fewer, longer chunks (one lock-step wave) unless clearly better
            best = cost;
            best_c = c;
        }
    }
    int chunk = (Wl + best_c - 1) / best_c;
    return chunk < FDLBM_MIN_CHUNK ? FDLBM_MIN_CHUNK : chunk;
}

// returns 0 or a cudaError_t
template <typename T, int HPC>
int launch_fused_hp(const LbmParams<T> &P, cudaStream_t stream)
{
    using C = FusedCfg<T, FUSED_TY>;
    auto kern = k_fused<T, FUSED_TY, HPC>;
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
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FUSED_TY, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) occ = 1;
        n_cta = sms * occ;
        if (getenv("FDLBM_DEBUG")) fprintf(stderr, "fdlbm: k_fused %d B smem, %d CTAs/SM x %d SMs\n", (int)C::SMEM, occ, sms);
    }
    // column chunks per strip, all strips of a chunk in step (strips are the fast CTA index)
    const int nyt = (P.H + FUSED_TY - 1) / FUSED_TY;
    int fx0, fx1;
    fused_plain_range(P, fx0, fx1);
    const int n_face = nyt * ((fx0 > 0) + (fx1 < P.Wl));
    // one wave of plain-column CTAs; keep a slot free for the face CTAs to slip into
    const int slots = n_face && n_cta > nyt ? n_cta - 1 : n_cta;
    int chunk = 8, nchunks = 0;
    if (fx1 > fx0) {
        chunk = fused_chunk(nyt, slots, fx1 - fx0);
        nchunks = (fx1 - fx0 + chunk - 1) / chunk;
    }
    kern<<<nyt * nchunks + n_face, FUSED_TY, C::SMEM, stream>>>(P, nyt, chunk, fx0, fx1, nyt * nchunks, P.H % 32 != 1);
    return 0;
}

// the common row pitches get a kernel with Hp folded into the instruction immediates
template <typename T>
int launch_fused(const LbmParams<T> &P, cudaStream_t stream)
{
#ifndef FDLBM_NO_HP_SPECIALISATION  // +2.1 .. 2.6 % with the PLAIN body (profiles/r2/r2_ab_lean1.txt; it was +-0 before the split)
    switch (P.Hp) {
    case 2048: return launch_fused_hp<T, 2048>(P, stream);
    case 4096: return launch_fused_hp<T, 4096>(P, stream);
    case 8192: return launch_fused_hp<T, 8192>(P, stream);
    default: break;
    }
#endif
    return launch_fused_hp<T, 0>(P, stream);
}

}  // namespace fdlbm
