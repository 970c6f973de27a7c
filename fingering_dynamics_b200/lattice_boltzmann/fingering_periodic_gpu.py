"""Drop-in twin of the reference's lattice_boltzmann/fingering_periodic_gpu.py, the CuPy transliteration of
fingering_periodic.py with its own constant block (fingering_periodic_gpu.py:20-45: W = 420, MAX_T = 1000,
psi_wall = 1, Pe = 400, ...), a uniform face velocity u0 on both Zou-He faces (:295-379), an injected band of 10
columns (:92), random initial density 1 - 0.05 * rand (:102) and an empty obstacle list (:473).  The reference version
issues ~10^3 CuPy library kernels per step; here the same step runs in the fused sm_100a kernel.  Arrays are NumPy on
the host side of the boundary (the global NumPy RNG stands in for cp.random).
"""
import math

import numpy as np

try:
    from ._compute import ComputeBase, E9, W9, run_loop, stream as _stream
    from .create_block import Createblock
    from .bounce_back import Bounce_back
    from .. import geometry as _geo
except ImportError:  # run from inside this directory, like the reference
    from _compute import ComputeBase, E9, W9, run_loop, stream as _stream
    from create_block import Createblock
    from bounce_back import Bounce_back
    from fingering_dynamics_b200 import geometry as _geo

# ---- constants (fingering_periodic_gpu.py:20-45) ----------------------------------------------------------
H = 400
W = 420
MAX_T = 1000
psi_wall = 1.0
Pe = 400
C_W = 5.0 * (10 ** (-5)) / W
Ca = 7.33 * 10 ** (-3)
M = 20.0
R_Nu = 10 ** (-6)
tau = 1 / (3.0 - math.sqrt(3))
rho0 = 1.0
n_non = 1.0
R_sigma = 0.045
C_rho = 1.0 * 10 ** 3
v0 = (tau - 0.5) / 3
C_t = v0 / R_Nu * (C_W ** 2)
Eta_n = 0.001 / (C_rho * (C_W ** 2) / C_t)
sigma = R_sigma * (C_t ** 2) / (C_rho * (C_W ** 3))
u0 = Ca * sigma / (rho0 * v0)
xi = 2.0
kappa = 0.75 * sigma * xi
a = - 2.0 * kappa / (xi ** 2)
gamma = u0 * W / ((-a * Pe) * (tau - 0.5))


class Compute(ComputeBase):
    """fingering_periodic_gpu.py:48-384.  Same step as fingering_periodic.Compute (y periodic, Zou-He with the 2/3
    coefficient on all rows, no corner nodes) with a uniform face velocity."""
    ZOU_HE, Y_WALL, X_PERIODIC, A_SIGN, F3 = "fp", False, False, 1.0, 2 / 3
    _full_grid = False

    def _profiles(self):
        p = np.full(self._m.H, float(self._m.u0))  # ux = u0 on every row (fingering_periodic_gpu.py:296,364)
        return p, p

    def __init__(self, mask):
        """initial state of fingering_periodic_gpu.py:48-123: psi = +1 on the first 10 columns, random rho, u = 0;
        mu is computed into a local and dropped (:117), so the first collision sees mu = 0 like fingering_periodic.py"""
        self.mask = np.asarray(mask, dtype=bool)
        self.e = E9.copy()
        self.w = W9.copy()
        self.psi = np.full((H, W), -1.0)
        self.psi[:, :10] = 1.0
        self.block_mask = np.logical_not(self.mask)
        self.psi[self.block_mask] = psi_wall
        self.left_wall = np.full((H, 1), 1.0)
        self.right_wall = np.full((H, 1), -1.0)
        self.gamma = gamma
        self.top_bottom_wall = np.full((1, W + 2), psi_wall)
        n = int(self.mask.sum())
        self.rho = 1.0 - 0.05 * np.random.rand(H, W)[self.mask]
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


def update(i, x, y, cc):
    """animation callback of the reference (fingering_periodic_gpu.py, same as fingering_periodic.py:346-349): draw psi frame i of cc"""
    print(i)
    import matplotlib.pyplot as plt
    plt.cla()
    plt.pcolor(x, y, cc[i], label='MAX_T{}_Pe{}_M{}_Ca{}_wall{}'.format(MAX_T, Pe, M, Ca, psi_wall), cmap='RdBu')


def main(max_t=None, show=False):
    """fingering_periodic_gpu.py:446-560: no obstacles (:473), psi frames every MAX_T // 150 iterations (:479,527).
    Returns cm with the frames in cm.frames (as the fingering.py twin does)."""
    cr = Createblock(H, W)
    Bounce_back(H, W)
    block_psi_all, side_list, concave_list, convex_list = cr.setCirleblock([])
    mask = np.logical_not(block_psi_all == 1)
    cm = Compute(mask)
    n = MAX_T if max_t is None else max_t
    cc = run_loop(cm, _geo.reflect_bits_circle(side_list, concave_list, convex_list), n,
                  frames_every=max(1, MAX_T // 150))
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
    main()
