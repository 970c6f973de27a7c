"""Drop-in twin of the reference's lattice_boltzmann/fingering.py ("fingering with top and bottom wall"):
module constants, `Compute(mask)`, `stream`, `bottom_top_wall`, `main()`; the loop of main()
(fingering.py:558-585) runs on the GPU.

Variant: psi ghost rows = psi_wall, rectangular obstacles, wall reflection on rows 1 and H-2 (fingering.py:573),
uniform Zou-He inlet with corner nodes, outlet with the 1.5 coefficient (fingering.py:378).
"""
import math

import numpy as np

try:
    from ._compute import ComputeBase, E9, W9, run_loop, stream as _stream, wall_rows as _wall_rows
    from .create_block import Createblock
    from .bounce_back import Bounce_back
    from .. import geometry as _geo
except ImportError:
    from _compute import ComputeBase, E9, W9, run_loop, stream as _stream, wall_rows as _wall_rows
    from create_block import Createblock
    from bounce_back import Bounce_back
    from fingering_dynamics_b200 import geometry as _geo

# ---- constants (fingering.py:16-50) -----------------------------------------------------------------------
H = 380
W = 380
MAX_T = 1000
psi_wall = -1.0
Pe = 400
C_W = 1.5 * (10 ** (-7))
Ca = 3.0 * 7.33 * (10 ** (-3))
M = 20.0
Eta = 0.001
block_num = 10
R_Nu = Eta / 1000
tau = 1 / (3.0 - math.sqrt(3))
rho0 = 1.0
n_non = 1.0
R_sigma = 0.045
C_rho = 1000
v0 = (tau - 0.5) / 3
C_t = (v0 / R_Nu) * (C_W ** 2)
Eta_n = Eta * C_t / (C_rho * (C_W ** 2))
sigma = R_sigma * (C_t ** 2) / (C_rho * (C_W ** 3))
u0 = Ca * sigma / Eta_n
xi = 2.0
kappa = 0.75 * xi * sigma
a = - 2.0 * kappa / (xi ** 2)
gamma = u0 * H / ((-a * Pe) * (tau - 0.5))
Re = u0 * 20 / Eta_n


class Compute(ComputeBase):
    ZOU_HE, Y_WALL, X_PERIODIC, A_SIGN, F3 = "fg", True, False, 1.0, 1.5
    _full_grid = False

    def __init__(self, mask):
        """initial state of fingering.py:53-127: random rho from the GLOBAL NumPy RNG (same two draws as
        the reference, so np.random.seed(k) reproduces it), p with mu = 0, then mu, then uy (ux stays 0)."""
        self.mask = np.asarray(mask, dtype=bool)
        self.e = E9.copy()
        self.w = W9.copy()
        self.psi = np.full((H, W), -1.0)
        self.psi[:, :5] = 1.0
        self.block_mask = np.logical_not(self.mask)
        self.psi[self.block_mask] = psi_wall
        self.left_wall = np.full((H, 1), 1.0)
        self.right_wall = np.full((H, 1), -1.0)
        self.gamma = gamma
        self.top_bottom_wall = np.full((1, W + 2), psi_wall)
        sign = np.random.randint(0, 1, size=(H, W)) * 2 - 1.0
        n = int(self.mask.sum())
        self.rho = np.ones(n) + np.random.rand(H, W)[self.mask] * 0.001 * sign[self.mask]
        self.ux, self.uy, self.mu = np.zeros(n), np.zeros(n), np.zeros(n)
        self.f = np.zeros((9, H, W))
        self.g = np.zeros((9, H, W))
        self.nabla_psix, self.nabla_psiy, self.nabla_psi2 = self._stencils()
        self.p = self.getP()
        self.mu = self.getMu()
        self.uy = self.mu * self.nabla_psiy[self.mask] / 2 / self.rho  # getUy with f = 0 (fingering.py:123,140-145)
        self.mix_tau = self.getMix_tau()
        feq, geq, F = self._terms()
        self.feq = np.array([feq[i][self.mask] for i in range(9)])
        self.geq = np.array([geq[i][self.mask] for i in range(9)])
        self.F = np.zeros((9, n))
        for i in range(9):
            self.f[i][self.mask] = self.feq[i]
            self.g[i][self.mask] = self.geq[i]


def stream(f, g):
    _stream(f, g)


def bottom_top_wall(f_behind, g_behind, f, g):
    _wall_rows(f_behind, g_behind, f, g)


def default_rectangles():
    """the staggered 21x21 squares of main() (fingering.py:534-550)"""
    rects, count, flag = [], 1, True
    while (count + 1) * 20 <= 380:
        rows = range(4) if flag else range(5)
        for i in rows:
            if flag:
                rects.append(((count * 20, 60 * (i + 1) + i * 20), ((count + 1) * 20, 60 * (i + 1) + (i + 1) * 20)))
            else:
                rects.append(((count * 20, 60 * i + (i + 1) * 20), ((count + 1) * 20, 60 * i + (i + 2) * 20)))
        flag = not flag
        count += 2
    return rects


def reflect_bits(corner_list):
    """rectangle classes (bounce_back.py:25-86) OR the wall rows 1 and H-2 (fingering.py:573)"""
    return _geo.reflect_bits_rect(corner_list, H, W) | _geo.reflect_bits_wall_rows(H, W, 1, H - 2)


def update(i, x, y, cc):
    """animation callback of the reference (fingering.py:422-425): draw psi frame i of cc"""
    print(i)
    import matplotlib.pyplot as plt
    plt.cla()
    plt.pcolor(x, y, cc[i], label='MAX_T{}_Pe{}_M{}_Ca{}_wall{}'.format(MAX_T, Pe, M, Ca, psi_wall), cmap='RdBu')


def main(max_t=None, show=True):
    cr = Createblock(H, W)
    Bounce_back(H, W)
    block_psi_all, corner_list = cr.setblock(default_rectangles())
    mask = np.logical_not(block_psi_all == 1)
    cm = Compute(mask)
    n = MAX_T if max_t is None else max_t
    cc = run_loop(cm, reflect_bits(corner_list), n, frames_every=max(1, MAX_T // 100))  # fingering.py:565-566
    if show:
        try:
            import matplotlib.pyplot as plt
            plt.figure()
            plt.pcolor(list(range(W)), list(range(H)), cm.psi, cmap='RdBu')
            plt.colorbar()
            plt.show()
        except ImportError:
            pass
    cm.frames = cc
    return cm


if __name__ == '__main__':
    import time
    t1 = time.time()
    main()
    print((time.time() - t1) / 60)
