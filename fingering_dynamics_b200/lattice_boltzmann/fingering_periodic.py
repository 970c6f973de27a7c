"""Drop-in twin of the reference's lattice_boltzmann/fingering_periodic.py ("fingering without top and
bottom wall"): same module constants, `Compute(mask)`, `stream`, `bottom_top_wall`, `main()`; the time
loop of main() (fingering_periodic.py:454-479) runs on the GPU (fingering_dynamics_b200.Engine).

Variant: y periodic, circular obstacles, Gaussian-profile Zou-He velocity inlet AND outlet.
"""
import math

import numpy as np

try:
    from ._compute import ComputeBase, E9, W9, run_loop, stream as _stream, wall_rows as _wall_rows
    from .create_block import Createblock
    from .bounce_back import Bounce_back
    from .. import geometry as _geo
except ImportError:  # run from inside this directory, like the reference
    from _compute import ComputeBase, E9, W9, run_loop, stream as _stream, wall_rows as _wall_rows
    from create_block import Createblock
    from bounce_back import Bounce_back
    from fingering_dynamics_b200 import geometry as _geo

# ---- constants (fingering_periodic.py:15-44) -------------------------------------------------------------
H = 400
W = 400
MAX_T = 4000
psi_wall = -0.5
Pe = 15
C_W = 1.5 * 10 ** (-7)
Ca = 2.0 * 7.33 * 10 ** (-3)
M = 20.0
Eta = 0.001
block_num = 10
R_Nu = Eta / 1000
tau = 1 / (3.0 - math.sqrt(3))
rho0 = 1.0
n_non = 1.0
R_sigma = 0.045
C_rho = 1.0 * 10 ** 3
v0 = (tau - 0.5) / 3
C_t = v0 / R_Nu * (C_W ** 2)
Eta_n = Eta / (C_rho * (C_W ** 2) / C_t)
sigma = R_sigma * (C_t ** 2) / (C_rho * (C_W ** 3))
u0 = Ca * sigma / (rho0 * v0)
xi = 2.0
kappa = 0.75 * sigma * xi
a = - 2.0 * kappa / (xi ** 2)
gamma = u0 * 20 / ((-a * Pe) * (tau - 0.5))
Re = u0 * 20 / Eta_n


class Compute(ComputeBase):
    ZOU_HE, Y_WALL, X_PERIODIC, A_SIGN, F3 = "fp", False, False, 1.0, 2 / 3
    _full_grid = False

    def __init__(self, mask):
        """initial state of fingering_periodic.py:47-121 (mu stays 0 for the first collision)"""
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
        n = int(self.mask.sum())
        self.rho = np.ones(n) * rho0
        self.ux, self.uy, self.mu = np.zeros(n), np.zeros(n), np.zeros(n)
        self.f = np.zeros((9, H, W))
        self.g = np.zeros((9, H, W))
        self.nabla_psix, self.nabla_psiy, self.nabla_psi2 = self._stencils()
        self.p = self.getP()
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


def default_circles():
    """the obstacle list of main(): 9 columns x 10 rows of r=10 circles (fingering_periodic.py:406-420)"""
    circle_list, r, xx, count = [], 10, 10, 1
    z = r + xx - 25
    while count * (xx + r) - z <= 380:
        for i in range(block_num):
            circle_list.append(((count * (r + xx) - z, (2 * i + 1) * (r + xx)), r))
        count += 2
    return circle_list


def update(i, x, y, cc):
    """animation callback of the reference (fingering_periodic.py:346-349): draw psi frame i of cc"""
    print(i)
    import matplotlib.pyplot as plt
    plt.cla()
    plt.pcolor(x, y, cc[i], label='MAX_T{}_Pe{}_M{}_Ca{}_wall{}'.format(MAX_T, Pe, M, Ca, psi_wall), cmap='RdBu')


def main(max_t=None, show=True):
    cr = Createblock(H, W)
    Bounce_back(H, W)
    block_psi_all, side_list, concave_list, convex_list = cr.setCirleblock(default_circles())
    mask = np.logical_not(block_psi_all == 1)
    cm = Compute(mask)
    run_loop(cm, _geo.reflect_bits_circle(side_list, concave_list, convex_list), MAX_T if max_t is None else max_t)
    if show:
        try:
            import matplotlib.pyplot as plt
            plt.figure()
            plt.pcolor(list(range(W)), list(range(H)), cm.psi, cmap='RdBu')
            plt.colorbar()
            plt.gca().set_aspect('equal', adjustable='box')
            plt.show()
        except ImportError:
            pass
    return cm


if __name__ == '__main__':
    import time
    t1 = time.time()
    main()
    print((time.time() - t1) / 60)
