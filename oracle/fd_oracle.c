/*
 * fd_oracle.c -- CPU restatement of the reference's two-phase D2Q9 time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or as the timed CPU arm).  The product path is the CUDA library
 * in fingering_dynamics_b200/csrc and fails loudly without it.
 *
 * What it restates (file:line into /root/reference/lattice_boltzmann):
 *   fingering_periodic.py:123-264  macroscopic moments, equilibria, forcing, BGK   (FP)
 *   fingering_periodic.py:268-324  Zou-He inlet/outlet, Gaussian profile           (FP)
 *   fingering_periodic.py:327-343  stream (periodic roll)                          (FP/FG/VA)
 *   fingering.py:219-286           stencils with top/bottom ghost rows             (FG)
 *   fingering.py:298-390           Zou-He inlet/outlet with corner nodes, 1.5      (FG)
 *   fingering.py:432-451,573       wall-row reflection on rows 1 and H-2           (FG)
 *   validation.py:98-190,228-320   full-grid variant with float e, cs^2            (VA)
 *   validation.py:357-376          wall rows 0 and H-1                             (VA)
 *   bounce_back.py:13-22,25-86,89-167  left_boundary / rectangle / circle tables   (BB)
 *
 * Pinning: parity is NOT pinned by reference tests (the reference has none).  It is pinned by
 * tests/golden/*.npz, produced by tests/golden/make_golden.py which executes the unmodified
 * reference in the build container.  This file is compiled with -ffp-contract=off and keeps the
 * reference's operation order, so it reproduces those fixtures BIT FOR BIT (tests/test_oracle.py).
 *
 * Layout: every array is the reference's own: C-order double, populations (9,H,W), fields (H,W),
 * x = axis 1 = flow direction.  "Masked" 1-D reference arrays are carried as full (H,W) grids whose
 * solid entries are ignored.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct {
    int H, W;
    double tau, gamma, a, kappa, Eta_n, M, psi_wall;
    double psi_left, psi_right; /* ghost columns, fingering_periodic.py:94-95 */
    int x_periodic;             /* 1: validation.py (no ghost columns)        */
    int y_wall;                 /* 1: ghost rows = psi_wall (FG, VA); 0: y periodic (FP) */
    int lap_order;              /* 0: -20C,N,E,W,S (FP:247-256, FG:267-285); 1: -20C,E,N,W,S (VA:291-309) */
    double outlet_f3_coef;      /* 2/3 (FP:317) or 1.5 (FG:378) */
} fdo_params;

static const int EX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int EY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

#define IDX(y, x) ((size_t)(y) * (size_t)W + (size_t)(x))
#define POP(a, i, y, x) ((a)[(size_t)(i) * (size_t)H * (size_t)W + IDX(y, x)])

/* psi seen by the stencils at (y,x) including ghost rows / columns.
 * FP:216-218 (left/right ghost columns, y wraps through np.roll(axis=0));
 * FG:221-224 (+ ghost rows = psi_wall across the padded width, so ghost corners are psi_wall);
 * VA:246    (ghost rows, x wraps). */
static inline double psi_at(const fdo_params *P, const double *psi, int y, int x)
{
    const int H = P->H, W = P->W;
    if (P->y_wall) {
        if (y < 0 || y >= H) return P->psi_wall;
    } else {
        if (y < 0) y += H;
        if (y >= H) y -= H;
    }
    if (P->x_periodic) {
        if (x < 0) x += W;
        if (x >= W) x -= W;
    } else {
        if (x < 0) return P->psi_left;
        if (x >= W) return P->psi_right;
    }
    return psi[IDX(y, x)];
}

/* Isotropic 9-point stencils; accumulation order is the reference's (see header). */
static inline void stencil_cell(const fdo_params *P, const double *psi, int y, int x,
                                double *gx, double *gy, double *lap)
{
    const double C = psi_at(P, psi, y, x);
    const double E = psi_at(P, psi, y, x + 1), Wv = psi_at(P, psi, y, x - 1);
    const double N = psi_at(P, psi, y + 1, x), S = psi_at(P, psi, y - 1, x);
    const double NE = psi_at(P, psi, y + 1, x + 1), NW = psi_at(P, psi, y + 1, x - 1);
    const double SW = psi_at(P, psi, y - 1, x - 1), SE = psi_at(P, psi, y - 1, x + 1);
    double acc;
    if (gx) {
        acc = 0.0;
        acc += 4 * E;
        acc += -4 * Wv;
        acc += NE;
        acc += -NW;
        acc += -SW;
        acc += SE;
        *gx = acc / 12;
    }
    if (gy) {
        acc = 0.0;
        acc += 4 * N;
        acc += -4 * S;
        acc += NE;
        acc += NW;
        acc += -SW;
        acc += -SE;
        *gy = acc / 12;
    }
    if (lap) {
        acc = 0.0;
        acc += -20 * C;
        if (P->lap_order == 0) {
            acc += 4 * N;
            acc += 4 * E;
        } else {
            acc += 4 * E;
            acc += 4 * N;
        }
        acc += 4 * Wv;
        acc += 4 * S;
        acc += NE;
        acc += NW;
        acc += SW;
        acc += SE;
        *lap = acc / 6;
    }
}

EXPORT void fdo_stencils(const fdo_params *P, const double *psi, double *gx, double *gy, double *lap)
{
    const int H = P->H, W = P->W;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            stencil_cell(P, psi, y, x, gx ? gx + IDX(y, x) : NULL, gy ? gy + IDX(y, x) : NULL,
                         lap ? lap + IDX(y, x) : NULL);
}

/* mu = a psi (1 - psi^2) - kappa lap   (FP:141-149).  VA's a psi (psi^2-1) with a>0 (VA:117-118) is
 * the same value bit for bit when called with -a (negation is exact). */
static inline double mu_of(const fdo_params *P, double psi, double lap)
{
    return P->a * psi * (1.0 - psi * psi) - P->kappa * lap;
}

/* tau_mix, FP:201-208 */
static inline double mix_tau_of(const fdo_params *P, double rho, double psi)
{
    const double v1 = P->Eta_n / rho;
    const double v2 = P->Eta_n * P->M / rho;
    const double mix_v = (2 * v1 * v2) / (v1 * (1.0 - psi) + v2 * (1.0 + psi));
    return 3 * mix_v + 0.5;
}

/* ------------------------------------------------------------------------------------------ */
/* collision on fluid cells: FP:455-460 with FP:155-199, 258-264                                */
/* ------------------------------------------------------------------------------------------ */
EXPORT void fdo_collide(const fdo_params *P, const uint8_t *mask, double *f, double *g,
                        const double *rho_, const double *ux_, const double *uy_, const double *p_,
                        const double *mu_, const double *mix_tau_, const double *psi_,
                        const double *gx_, const double *gy_)
{
    const int H = P->H, W = P->W;
    const double w[9] = {4.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36};
    const double c0 = 3.0 * (1.0 - w[0]);
    const double c0g = c0 * P->gamma;
    const double g3 = 3 * P->gamma;
    const double inv_tau = 1.0 / P->tau;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (!mask[IDX(y, x)]) continue;
            const size_t c = IDX(y, x);
            const double rho = rho_[c], ux = ux_[c], uy = uy_[c], p = p_[c], mu = mu_[c];
            const double mt = mix_tau_[c], psi = psi_[c], gx = gx_[c], gy = gy_[c];
            const double usq = ux * ux + uy * uy;
            const double A0 = (rho - c0 * p) / w[0], A18 = 3 * p;
            const double B0 = (psi - c0g * mu) / w[0], B18 = g3 * mu;
            const double pref = 1 - 1 / (2 * mt);
            const double inv_mt = 1.0 / mt;
            for (int i = 0; i < 9; ++i) {
                const double ex = EX[i], ey = EY[i];
                const double eu = ex * ux + ey * uy;
                const double poly = 3 * eu + 4.5 * (eu * eu) - 1.5 * usq;
                const double Fi = mu * w[i] * pref *
                                  (((ex - ux) * 3 + ex * eu * 9) * gx + ((ey - uy) * 3 + ey * eu * 9) * gy);
                const double feq = w[i] * ((i == 0 ? A0 : A18) + rho * poly);
                const double geq = w[i] * ((i == 0 ? B0 : B18) + psi * poly);
                double *fp = &POP(f, i, y, x), *gp = &POP(g, i, y, x);
                *fp = *fp - inv_mt * (*fp - feq) + Fi;
                *gp = *gp - inv_tau * (*gp - geq);
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* stream: f_i(y,x) <- f_i(y-ey, x-ex) with wrap on both axes, all cells (FP:327-343)           */
/* ------------------------------------------------------------------------------------------ */
EXPORT void fdo_stream(int H, int W, double *f, double *g)
{
    double *tmp = (double *)malloc(sizeof(double) * (size_t)H * W);
    for (int pass = 0; pass < 2; ++pass) {
        double *a = pass ? g : f;
        for (int i = 1; i < 9; ++i) {
            double *ai = a + (size_t)i * H * W;
            memcpy(tmp, ai, sizeof(double) * (size_t)H * W);
#pragma omp parallel for schedule(static)
            for (int y = 0; y < H; ++y) {
                int ys = y - EY[i];
                ys = ys < 0 ? ys + H : (ys >= H ? ys - H : ys);
                for (int x = 0; x < W; ++x) {
                    int xs = x - EX[i];
                    xs = xs < 0 ? xs + W : (xs >= W ? xs - W : xs);
                    ai[IDX(y, x)] = tmp[IDX(ys, xs)];
                }
            }
        }
    }
    free(tmp);
}

static const int OPP[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

static void par_copy(double *dst, const double *src, size_t n)
{
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)n; ++k) dst[k] = src[k];
}

static inline void reflect(int H, int W, const double *fb, const double *gb, double *f, double *g, int i,
                           int y, int x)
{
    POP(f, i, y, x) = POP(fb, OPP[i], y, x);
    POP(g, i, y, x) = POP(gb, OPP[i], y, x);
}

/* bounce_back.py:89-167.  masks = (12,H,W) uint8 in the order create_block.py:207-218 returns them:
 * side[top,bottom,right,left], concave[tr,tl,br,bl], convex[tr,tl,br,bl]; consumed as
 * n,s,e,w / nw,ne,sw,se (bounce_back.py:90-101). */
EXPORT void fdo_bb_circle(int H, int W, const uint8_t *masks, const double *fb, const double *gb,
                          double *f, double *g)
{
    /* directions reflected by each of the 12 classes */
    static const int tab[12][3] = {
        {2, 5, 6},  /* n_barrier  = side[0]  BB:112,137,146 */
        {4, 7, 8},  /* s_barrier  = side[1]  BB:126,155,164 */
        {3, 6, 7},  /* e_barrier  = side[2]  BB:119,148,157 */
        {1, 5, 8},  /* w_barrier  = side[3]  BB:105,139,166 */
        {1, 2, 5},  /* nw_cave    = concave[0] BB:107,114,135 */
        {2, 3, 6},  /* ne_cave    = concave[1] BB:116,121,144 */
        {1, 4, 8},  /* sw_cave    = concave[2] BB:109,128,162 */
        {3, 4, 7},  /* se_cave    = concave[3] BB:123,130,153 */
        {5, -1, -1}, /* nw_vex    = convex[0] BB:133 */
        {6, -1, -1}, /* ne_vex    = convex[1] BB:142 */
        {8, -1, -1}, /* sw_vex    = convex[2] BB:160 */
        {7, -1, -1}, /* se_vex    = convex[3] BB:151 */
    };
    for (int k = 0; k < 12; ++k) {
        const uint8_t *m = masks + (size_t)k * H * W;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                if (m[IDX(y, x)])
                    for (int t = 0; t < 3; ++t)
                        if (tab[k][t] > 0) reflect(H, W, fb, gb, f, g, tab[k][t], y, x);
    }
}

/* bounce_back.py:25-86.  corners = n x 8 ints:
 * top_left(x,y), bottom_left(x,y), top_right(x,y), bottom_right(x,y)  (create_block.py:39-48) */
EXPORT void fdo_bb_rect(int H, int W, const int *corners, int n, const double *fb, const double *gb,
                        double *f, double *g)
{
    uint8_t *cls = (uint8_t *)calloc((size_t)8 * H * W, 1); /* n,s,e,w,nw,ne,sw,se */
#define CLS(k, y, x) cls[(size_t)(k) * H * W + IDX(y, x)]
    for (int r = 0; r < n; ++r) {
        const int *c = corners + 8 * r;
        const int tlx = c[0], tly = c[1], blx = c[2], bly = c[3], trx = c[4], brx = c[6];
        for (int x = tlx; x <= trx; ++x) {
            CLS(0, tly + 1, x) = 1; /* BB:37 */
            CLS(1, bly - 1, x) = 1; /* BB:38 */
        }
        for (int y = bly; y <= tly; ++y) {
            CLS(3, y, tlx - 1) = 1; /* w_barrier BB:39 */
            CLS(2, y, trx + 1) = 1; /* e_barrier BB:40 */
        }
        CLS(4, tly + 1, tlx - 1) = 1; /* nw BB:41 */
        CLS(5, tly + 1, trx + 1) = 1; /* ne BB:42 */
        CLS(6, bly - 1, blx - 1) = 1; /* sw BB:43 */
        CLS(7, bly - 1, brx + 1) = 1; /* se BB:44 */
    }
    static const int tab[8][3] = {
        {2, 5, 6},   /* n: BB:51,62,69 */
        {4, 7, 8},   /* s: BB:57,76,83 */
        {1, 5, 8},   /* e: BB:48,64,85 */
        {3, 6, 7},   /* w: BB:54,71,78 */
        {6, -1, -1}, /* nw corner BB:67 */
        {5, -1, -1}, /* ne corner BB:60 */
        {7, -1, -1}, /* sw corner BB:74 */
        {8, -1, -1}, /* se corner BB:81 */
    };
    for (int k = 0; k < 8; ++k)
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                if (CLS(k, y, x))
                    for (int t = 0; t < 3; ++t)
                        if (tab[k][t] > 0) reflect(H, W, fb, gb, f, g, tab[k][t], y, x);
#undef CLS
    free(cls);
}

/* Wall rows: row_lo reflects {2,5,6}, row_hi reflects {4,7,8}.
 * FG: bottom_top_wall on the [:,1:-1] row slice => rows 1 and H-2 (fingering.py:432-451,573);
 * VA: halfway_bounceback rows 0 and H-1 (validation.py:357-376). */
EXPORT void fdo_wall_rows(int H, int W, int row_lo, int row_hi, const double *fb, const double *gb,
                          double *f, double *g)
{
    static const int lo[3] = {2, 5, 6}, hi[3] = {4, 7, 8};
    for (int x = 0; x < W; ++x)
        for (int t = 0; t < 3; ++t) {
            reflect(H, W, fb, gb, f, g, lo[t], row_lo, x);
            reflect(H, W, fb, gb, f, g, hi[t], row_hi, x);
        }
}

/* bounce_back.py:13-22 (not called by any driver) */
EXPORT void fdo_left_boundary(int H, int W, int hole, const double *fb, const double *gb, double *f, double *g)
{
    static const int d[3] = {1, 5, 8};
    const int a = (int)(H / 2.0 - hole), b = (int)(H / 2.0 + hole);
    for (int y = 0; y < H; ++y)
        if (y < a || y >= b)
            for (int t = 0; t < 3; ++t) reflect(H, W, fb, gb, f, g, d[t], y, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* Zou-He faces                                                                                 */
/* ------------------------------------------------------------------------------------------ */
/* FP:268-324 (corners=0, all rows) and FG:298-390 (corners=1: inlet rows 1..H-2 + corner nodes,
 * outlet corner copies).  inlet_ux / outlet_ux are per-row profiles (FP:270-271; FG: u0 everywhere).
 * psi, lap are the PREVIOUS psi and its stored Laplacian (getMu_plain uses self.nabla_psi2, FP:146-149). */
EXPORT void fdo_zou_he(const fdo_params *P, int corners, const double *inlet_ux, const double *outlet_ux,
                       double *f, double *g, const double *psi, const double *lap)
{
    const int H = P->H, W = P->W;
    const double w1 = 1.0 / 9, w5 = 1.0 / 36;
    double *rho_in = (double *)malloc(sizeof(double) * H);
    /* inlet, column 0 */
    for (int y = 0; y < H; ++y) {
        double psx, psy;
        stencil_cell(P, psi, y, 0, &psx, &psy, NULL);
        const double mu = mu_of(P, psi[IDX(y, 0)], lap[IDX(y, 0)]);
        const double ux = inlet_ux[y];
#define F(i) POP(f, i, y, 0)
#define G(i) POP(g, i, y, 0)
        rho_in[y] = 1 / (1 - ux) * (F(0) + F(2) + F(4) + 2 * (F(3) + F(6) + F(7)) - psx * mu / 2);
        const double psi_in = 1.0 - (G(0) + G(2) + G(3) + G(4) + G(6) + G(7));
        if (corners && (y == 0 || y == H - 1)) continue;
        const double den = w1 + w5 + w5;
        F(1) = F(3) + 2.0 / 3 * ux * rho_in[y] - psx * mu / 6;
        G(1) = w1 * psi_in / den;
        F(5) = F(7) - 0.5 * (F(2) - F(4)) + 1.0 / 6.0 * ux * rho_in[y] - psx * mu / 6 - psy * mu / 4;
        G(5) = w5 * psi_in / den;
        F(8) = F(6) + 0.5 * (F(2) - F(4)) + 1.0 / 6.0 * ux * rho_in[y] - psx * mu / 6 + psy * mu / 4;
        G(8) = w5 * psi_in / den;
#undef F
#undef G
    }
    if (corners) {
        /* left bottom corner node, FG:335-348 (sequential: later lines read earlier results) */
        {
            const int y = 0;
#define F(i) POP(f, i, y, 0)
#define G(i) POP(g, i, y, 0)
            F(1) = F(3); G(1) = G(3);
            F(2) = F(4); G(2) = G(4);
            F(5) = F(7); G(5) = G(7);
            F(6) = 0.5 * (rho_in[1] - (F(0) + F(1) + F(2) + F(3) + F(4) + F(5) + F(7)));
            G(6) = w5 * (1.0 - (G(0) + G(1) + G(2) + G(3) + G(4) + G(5) + G(7))) / (w5 + w5);
            F(8) = F(6); G(8) = G(6);
#undef F
#undef G
        }
        /* left top corner node, FG:351-364 */
        {
            const int y = H - 1;
#define F(i) POP(f, i, y, 0)
#define G(i) POP(g, i, y, 0)
            F(1) = F(3); G(1) = G(3);
            F(4) = F(2); G(4) = G(2);
            F(8) = F(6); G(8) = G(6);
            F(5) = 0.5 * (rho_in[H - 2] - (F(0) + F(1) + F(2) + F(3) + F(4) + F(6) + F(8)));
            G(5) = w5 * (1.0 - (G(0) + G(1) + G(2) + G(3) + G(4) + G(6) + G(8))) / (w5 + w5);
            F(7) = F(5); G(7) = G(5);
#undef F
#undef G
        }
    }
    free(rho_in);
    /* outlet, column W-1, all rows: FP:305-324 / FG:369-385 */
    for (int y = 0; y < H; ++y) {
        const int x = W - 1;
        double psx, psy;
        stencil_cell(P, psi, y, x, &psx, &psy, NULL);
        const double mu = mu_of(P, psi[IDX(y, x)], lap[IDX(y, x)]);
        const double ux = outlet_ux[y];
#define F(i) POP(f, i, y, x)
#define G(i) POP(g, i, y, x)
        const double rho_out = 1 / (1 + ux) * (F(0) + F(2) + F(4) + 2 * (F(1) + F(5) + F(8)) + psx * mu / 2);
        const double psi_out = -1.0 - (G(0) + G(1) + G(2) + G(4) + G(5) + G(8));
        const double den = w1 + w5 + w5;
        F(3) = F(1) - P->outlet_f3_coef * ux * rho_out + psx * mu / 6;
        G(3) = w1 * psi_out / den;
        F(6) = F(8) - 0.5 * (F(2) - F(4)) - 1.0 / 6.0 * ux * rho_out + psy * mu / 4 + psx * mu / 6;
        G(6) = w5 * psi_out / den;
        F(7) = F(5) + 0.5 * (F(2) - F(4)) - 1.0 / 6.0 * ux * rho_out - psy * mu / 4 + psx * mu / 6;
        G(7) = w5 * psi_out / den;
#undef F
#undef G
    }
    if (corners) { /* FG:387-390 */
        const int x = W - 1;
        POP(f, 2, 0, x) = POP(f, 4, 0, x);
        POP(g, 2, 0, x) = POP(g, 4, 0, x);
        POP(f, 4, H - 1, x) = POP(f, 2, H - 1, x);
        POP(g, 4, H - 1, x) = POP(g, 2, H - 1, x);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* moments, FP:470-479                                                                         */
/* ------------------------------------------------------------------------------------------ */
EXPORT void fdo_moments(const fdo_params *P, const uint8_t *mask, const double *f, const double *g,
                        double *psi, double *rho, double *ux, double *uy, double *p, double *mu,
                        double *mix_tau, double *gx, double *gy, double *lap)
{
    const int H = P->H, W = P->W;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t c = IDX(y, x);
            if (mask[c]) {
                double sr = POP(f, 0, y, x), sp = POP(g, 0, y, x);
                for (int i = 1; i < 9; ++i) {
                    sr += POP(f, i, y, x);
                    sp += POP(g, i, y, x);
                }
                rho[c] = sr; /* FP:151-152 */
                psi[c] = sp; /* FP:210-211 */
            } else {
                psi[c] = P->psi_wall; /* FP:212 */
            }
        }
    fdo_stencils(P, psi, gx, gy, lap);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t c = IDX(y, x);
            if (!mask[c]) continue;
            mu[c] = mu_of(P, psi[c], lap[c]);
            double tx = 0.0, ty = 0.0;
            for (int i = 0; i < 9; ++i) {
                tx += POP(f, i, y, x) * EX[i];
                ty += POP(f, i, y, x) * EY[i];
            }
            ux[c] = (tx + mu[c] * gx[c] / 2) / rho[c]; /* FP:126-131 */
            uy[c] = (ty + mu[c] * gy[c] / 2) / rho[c]; /* FP:134-139 */
            p[c] = 1.0 / 3 * rho[c] + psi[c] * mu[c]; /* FP:123-124 */
            mix_tau[c] = mix_tau_of(P, rho[c], psi[c]); /* FP:201-208 */
        }
}

/* ------------------------------------------------------------------------------------------ */
/* whole iterations                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double *f, *g, *psi, *rho, *ux, *uy, *p, *mu, *mix_tau, *gx, *gy, *lap; /* caller-owned */
} fdo_state;

/* geometry for the bounce-back of one iteration */
typedef struct {
    const uint8_t *mask;        /* fluid = 1 */
    const uint8_t *circ_masks;  /* (12,H,W) or NULL */
    const int *rect_corners;    /* n x 8 or NULL */
    int n_rects;
    int wall_lo, wall_hi;       /* wall rows, -1 = none */
    int zou_he;                 /* 0 none, 1 FP flavour, 2 FG flavour (corner nodes) */
    const double *inlet_ux, *outlet_ux;
} fdo_geom;

/* One reference iteration: FP:455-479 / FG:559-585. */
EXPORT void fdo_iterate(const fdo_params *P, const fdo_geom *G, fdo_state *S, int n_iter)
{
    const int H = P->H, W = P->W;
    const size_t np = (size_t)9 * H * W;
    double *fb = (double *)malloc(sizeof(double) * np), *gb = (double *)malloc(sizeof(double) * np);
    for (int it = 0; it < n_iter; ++it) {
        fdo_collide(P, G->mask, S->f, S->g, S->rho, S->ux, S->uy, S->p, S->mu, S->mix_tau, S->psi, S->gx, S->gy);
        par_copy(fb, S->f, np);
        par_copy(gb, S->g, np);
        fdo_stream(H, W, S->f, S->g);
        if (G->circ_masks) fdo_bb_circle(H, W, G->circ_masks, fb, gb, S->f, S->g);
        if (G->rect_corners) fdo_bb_rect(H, W, G->rect_corners, G->n_rects, fb, gb, S->f, S->g);
        if (G->wall_lo >= 0) fdo_wall_rows(H, W, G->wall_lo, G->wall_hi, fb, gb, S->f, S->g);
        if (G->zou_he) fdo_zou_he(P, G->zou_he == 2, G->inlet_ux, G->outlet_ux, S->f, S->g, S->psi, S->lap);
        fdo_moments(P, G->mask, S->f, S->g, S->psi, S->rho, S->ux, S->uy, S->p, S->mu, S->mix_tau, S->gx,
                    S->gy, S->lap);
    }
    free(fb);
    free(gb);
}

/* ------------------------------------------------------------------------------------------ */
/* validation.py, literal: float direction vectors e (VA:45-63), cs2 = cs**2, cs4 = cs**4       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double e[9][2];
    double w[9];
    double cs2, cs4;
    double a_va; /* +2 kappa / xi^2, VA:36 */
} fdo_va_consts;

EXPORT void fdo_va_iterate(const fdo_params *P, const fdo_va_consts *V, fdo_state *S, int n_iter)
{
    const int H = P->H, W = P->W;
    const size_t n = (size_t)H * W, np = 9 * n;
    double *fb = (double *)malloc(sizeof(double) * np), *gb = (double *)malloc(sizeof(double) * np);
    const double w0 = V->w[0];
    const double inv_tau = 1 / P->tau;
    const double c1w0g = (1.0 - w0) * P->gamma;
    for (int it = 0; it < n_iter; ++it) {
        /* VA:393-400: feq, geq from the stored macros; mix_tau from rho, psi; F with fresh gradients */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t c = IDX(y, x);
                const double rho = S->rho[c], psi = S->psi[c], ux = S->ux[c], uy = S->uy[c];
                const double p = S->p[c], mu = S->mu[c];
                const double mt = mix_tau_of(P, rho, psi); /* VA:228-236 */
                S->mix_tau[c] = mt;
                double gx, gy;
                stencil_cell(P, S->psi, y, x, &gx, &gy, NULL);
                const double usq = ux * ux + uy * uy;
                const double A0 = (rho - (1.0 - w0) * p / V->cs2) / w0, A18 = p / V->cs2; /* VA:130-138 */
                const double B0 = (psi - c1w0g * mu / V->cs2) / w0, B18 = P->gamma * mu / V->cs2; /* VA:140-146 */
                const double pref = 1 - 1 / (2 * mt);
                for (int i = 0; i < 9; ++i) {
                    const double ex = V->e[i][0], ey = V->e[i][1];
                    const double eu = ex * ux + ey * uy;
                    const double poly = 3 * eu + 4.5 * (eu * eu) - 1.5 * usq; /* c = 1: /c^2, /c^4 exact */
                    const double feq = V->w[i] * ((i == 0 ? A0 : A18) + rho * poly);
                    const double geq = V->w[i] * ((i == 0 ? B0 : B18) + psi * poly);
                    const double Fi = 1.0 * mu * V->w[i] * pref *
                                      (((ex - ux) / V->cs2 + ex * eu / V->cs4) * gx +
                                       ((ey - uy) / V->cs2 + ey * eu / V->cs4) * gy); /* VA:174-190 */
                    double *fp = &POP(S->f, i, y, x), *gp = &POP(S->g, i, y, x);
                    *fp = *fp - 1 / mt * (*fp - feq) + Fi; /* VA:313-315 */
                    *gp = *gp - inv_tau * (*gp - geq);     /* VA:318-320 */
                }
            }
        memcpy(fb, S->f, sizeof(double) * np);
        memcpy(gb, S->g, sizeof(double) * np);
        fdo_stream(H, W, S->f, S->g);
        fdo_wall_rows(H, W, 0, H - 1, fb, gb, S->f, S->g);
        /* VA:405-409 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                double sr = POP(S->f, 0, y, x), sp = POP(S->g, 0, y, x);
                for (int i = 1; i < 9; ++i) {
                    sr += POP(S->f, i, y, x);
                    sp += POP(S->g, i, y, x);
                }
                S->rho[IDX(y, x)] = sr;
                S->psi[IDX(y, x)] = sp;
            }
        fdo_stencils(P, S->psi, S->gx, S->gy, S->lap);
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t c = IDX(y, x);
                const double psi = S->psi[c];
                S->mu[c] = V->a_va * psi * (psi * psi - 1) - P->kappa * S->lap[c]; /* VA:122-123 */
                double tx = 0.0, ty = 0.0;
                for (int i = 0; i < 9; ++i) {
                    tx += POP(S->f, i, y, x) * V->e[i][0];
                    ty += POP(S->f, i, y, x) * V->e[i][1];
                }
                S->ux[c] = (tx + S->mu[c] * S->gx[c] * 1.0 / 2) / S->rho[c]; /* VA:113 */
                S->uy[c] = (ty + S->mu[c] * S->gy[c] * 1.0 / 2) / S->rho[c]; /* VA:114 */
                S->p[c] = V->cs2 * S->rho[c] + psi * S->mu[c];                /* VA:103-104 */
            }
    }
    free(fb);
    free(gb);
}

/* f = f_eq, g = g_eq on fluid cells from given macros (FP:119-121 with FP:171-192); solids untouched. */
EXPORT void fdo_equilibrium(const fdo_params *P, const uint8_t *mask, double *f, double *g,
                            const double *rho_, const double *ux_, const double *uy_, const double *p_,
                            const double *mu_, const double *psi_)
{
    const int H = P->H, W = P->W;
    const double w[9] = {4.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36};
    const double c0 = 3.0 * (1.0 - w[0]);
    const double c0g = c0 * P->gamma;
    const double g3 = 3 * P->gamma;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t c = IDX(y, x);
            if (!mask[c]) continue;
            const double rho = rho_[c], ux = ux_[c], uy = uy_[c], p = p_[c], mu = mu_[c], psi = psi_[c];
            const double usq = ux * ux + uy * uy;
            const double A0 = (rho - c0 * p) / w[0], A18 = 3 * p;
            const double B0 = (psi - c0g * mu) / w[0], B18 = g3 * mu;
            for (int i = 0; i < 9; ++i) {
                const double eu = EX[i] * ux + EY[i] * uy;
                const double poly = 3 * eu + 4.5 * (eu * eu) - 1.5 * usq;
                POP(f, i, y, x) = w[i] * ((i == 0 ? A0 : A18) + rho * poly);
                POP(g, i, y, x) = w[i] * ((i == 0 ? B0 : B18) + psi * poly);
            }
        }
}

EXPORT int fdo_abi_version(void) { return 1; }
