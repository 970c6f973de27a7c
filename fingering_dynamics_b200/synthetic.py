"""Synthetic porous domains for the benchmark configurations (BASELINE.json configs[3], configs[4];
SURVEY.md section 8(d) C4/C5): circular obstacles on a jittered lattice, fingering_periodic.py's step
variant (y-periodic, Gaussian Zou-He faces) and its constants and initial condition.

Everything here is host-side setup (geometry lists, initial arrays); the time step runs in the engine.
"""
import math

import numpy as np

from .lattice_boltzmann.create_block import Createblock
from . import geometry as geo


def fp_constants(H):
    """Derived constants of fingering_periodic.py:15-40 (independent of the grid except the profile)."""
    psi_wall = -0.5
    Pe = 15
    C_W = 1.5 * 10 ** (-7)
    Ca = 2.0 * 7.33 * 10 ** (-3)
    M = 20.0
    Eta = 0.001
    R_Nu = Eta / 1000
    tau = 1 / (3.0 - math.sqrt(3))
    rho0 = 1.0
    R_sigma = 0.045
    C_rho = 1.0 * 10 ** 3
    v0 = (tau - 0.5) / 3
    C_t = v0 / R_Nu * (C_W ** 2)
    Eta_n = Eta / (C_rho * (C_W ** 2) / C_t)
    sigma = R_sigma * (C_t ** 2) / (C_rho * (C_W ** 3))
    u0 = Ca * sigma / (rho0 * v0)
    xi = 2.0
    kappa = 0.75 * sigma * xi
    a = -2.0 * kappa / (xi ** 2)
    gamma = u0 * 20 / ((-a * Pe) * (tau - 0.5))
    t = np.array([i * 3 / (H / 2) for i in range(int(-H / 2), int(H / 2))])
    profile = u0 * np.exp(-(t ** 2) / 2)  # fingering_periodic.py:270-271
    return dict(tau=tau, gamma=gamma, a=a, kappa=kappa, Eta_n=Eta_n, M=M, psi_wall=psi_wall, u0=u0, rho0=rho0,
                inlet_ux=profile, outlet_ux=profile)


def porous_circles(H, W, seed=1234, pitch=40, r_lo=8, r_hi=12, jitter=4, free_cols=20):
    """Circle list [((cx, cy), r), ...] on a `pitch` lattice with centre jitter in [-jitter, jitter]^2 and
    radius in [r_lo, r_hi]; r+1 rings never touch (2*(r_hi+1) + 2*jitter < pitch); the first and last
    `free_cols` columns and one pitch next to the y edges stay free.  Deterministic in (H, W, seed):
    the lattice covers the GLOBAL grid, so every slab of a multi-GPU run sees the same obstacles."""
    assert 2 * (r_hi + 1) + 2 * jitter < pitch
    rng = np.random.default_rng(seed)
    nx = (W - 2 * free_cols) // pitch
    ny = (H - 2 * pitch) // pitch
    x_first = free_cols + (W - 2 * free_cols - nx * pitch) // 2 + pitch // 2
    y_first = pitch + (H - 2 * pitch - ny * pitch) // 2 + pitch // 2
    jx = rng.integers(-jitter, jitter + 1, size=(nx, ny))
    jy = rng.integers(-jitter, jitter + 1, size=(nx, ny))
    rr = rng.integers(r_lo, r_hi + 1, size=(nx, ny))
    out = []
    for k in range(nx):
        for m in range(ny):
            out.append(((int(x_first + k * pitch + jx[k, m]), int(y_first + m * pitch + jy[k, m])), int(rr[k, m])))
    return out


def porous_geometry(H, W, seed=1234, col0=0, ncols=None):
    """(solid, reflect) uint8 arrays of global columns [col0, col0+ncols) of the synthetic medium."""
    ncols = W if ncols is None else ncols
    # obstacles reach at most r_hi + 1 = 13 cells from their centre and are rasterised on a window of
    # r + 3 <= 15 cells: a 32-column apron keeps every obstacle that touches the slab un-clipped
    reach, pad = 16, 32
    lo, hi = max(0, col0 - pad), min(W, col0 + ncols + pad)
    local = [((cx - lo, cy), r) for (cx, cy), r in porous_circles(H, W, seed)
             if (lo == 0 or cx >= lo + reach) and (hi == W or cx < hi - reach)]
    bpa, side, cave, vex = Createblock(H, hi - lo).setCirleblock(local)
    solid = geo.solid_from_block_psi(bpa)
    refl = geo.reflect_bits_circle(side, cave, vex)
    s = slice(col0 - lo, col0 - lo + ncols)
    return np.ascontiguousarray(solid[:, s]), np.ascontiguousarray(refl[:, s])


def fp_initial_state(solid, c, col0=0, n_inject=5, alloc=np.zeros):
    """Initial arrays of fingering_periodic.py:90-121 for the columns held in `solid` (H, ncols):
    psi=-1, first 5 GLOBAL columns +1, solids psi_wall; rho=rho0, u=0, mu=0, p=rho/3, tau_mix from
    rho, psi; f=f_eq, g=g_eq on fluid cells, 0 on solids.  With u=0 and mu=0 the equilibria reduce to
    f_0 = rho - (5/3) p, f_i = 3 w_i p, g_0 = psi, g_i = 0 (fingering_periodic.py:155-192)."""
    H, ncols = solid.shape
    fluid = solid == 0
    psi = np.full((H, ncols), -1.0)
    k = max(0, min(ncols, n_inject - col0))
    psi[:, :k] = 1.0
    psi[~fluid] = c["psi_wall"]
    rho = np.full((H, ncols), float(c["rho0"]))
    zeros = np.zeros((H, ncols))
    p = 1 / 3 * rho
    v1 = c["Eta_n"] / rho
    v2 = c["Eta_n"] * c["M"] / rho
    mix_tau = 3 * (2 * v1 * v2 / (v1 * (1.0 - psi) + v2 * (1.0 + psi))) + 0.5
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    f = alloc((9, H, ncols))
    g = alloc((9, H, ncols))
    f[0] = np.where(fluid, w[0] * ((rho - 3.0 * (1.0 - w[0]) * p) / w[0]), 0.0)
    for i in range(1, 9):
        f[i] = np.where(fluid, w[i] * (3 * p), 0.0)
        g[i] = 0.0
    g[0] = np.where(fluid, w[0] * (psi / w[0]), 0.0)
    # nabla_psix / nabla_psiy: the first collision multiplies them by mu == 0 (fingering_periodic.py:111,
    # 194-199), so their value cannot influence the run; zeros keep the host set-up O(cells).
    return dict(f=f, g=g, psi=psi, rho=rho, ux=zeros, uy=zeros.copy(), p=p, mu=zeros.copy(), mix_tau=mix_tau,
                nabla_psix=zeros.copy(), nabla_psiy=zeros.copy())
