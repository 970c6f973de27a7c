"""Drop-in twin of the reference's lattice_boltzmann/fingering_periodic_gpu.py, the CuPy transliteration of
fingering_periodic.py (uniform inlet profile, W = 420, no obstacles, random initial density;
fingering_periodic_gpu.py:20-45, 102, 295-379, 473).  The reference version issues ~10^3 CuPy library kernels per
step; here the same step runs in the fused sm_100a kernel.  Arrays are NumPy on the host side of the boundary.
"""
import numpy as np

try:
    from . import fingering_periodic as _fp
    from ._compute import run_loop
    from .create_block import Createblock
    from .bounce_back import Bounce_back
    from .. import geometry as _geo
except ImportError:
    import fingering_periodic as _fp
    from _compute import run_loop
    from create_block import Createblock
    from bounce_back import Bounce_back
    from fingering_dynamics_b200 import geometry as _geo

H = 400
W = 420
MAX_T = 4000
psi_wall = _fp.psi_wall
M, tau, rho0, Eta_n, kappa, a, u0, gamma = _fp.M, _fp.tau, _fp.rho0, _fp.Eta_n, _fp.kappa, _fp.a, _fp.u0, _fp.gamma


class Compute(_fp.Compute):
    """fingering_periodic_gpu.py:48-123: as fingering_periodic.Compute with a uniform face velocity and
    rho = 1 - 0.001 * rand (the global NumPy RNG stands in for cp.random)."""
    ZOU_HE = "fg_uniform"

    def _profiles(self):
        p = np.full(self._m.H, float(self._m.u0))
        return p, p

    def _engine_kwargs(self):
        kw = super()._engine_kwargs()
        kw["zou_he"] = "fp"  # all rows, 2/3 coefficient, no corner nodes (fingering_periodic_gpu.py:295-379)
        return kw

    def _cfg(self):
        self.ZOU_HE = "fp"
        try:
            return super()._cfg()
        finally:
            self.ZOU_HE = "fg_uniform"


def stream(f, g):
    _fp.stream(f, g)


def main(max_t=None, show=False):
    _fp.H, _fp.W = H, W
    cr = Createblock(H, W)
    Bounce_back(H, W)
    block_psi_all, side_list, concave_list, convex_list = cr.setCirleblock([])  # fingering_periodic_gpu.py:473
    mask = np.logical_not(block_psi_all == 1)
    cm = Compute(mask)
    n = int(mask.sum())
    cm.rho = np.ones(n) - 0.001 * np.random.rand(H, W)[mask]
    run_loop(cm, _geo.reflect_bits_circle(side_list, concave_list, convex_list), MAX_T if max_t is None else max_t)
    return cm
