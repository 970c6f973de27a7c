"""Drop-in twin of the reference's lattice_boltzmann/validation.py (wettability check: a droplet on the
bottom wall): module constants, `Compute()`, `create_circle`, `stream`, `halfway_bounceback`, `main()`;
the loop of main() (validation.py:392-409) runs on the GPU.

Variant: full grid (no obstacles), x periodic, psi ghost rows = psi_wall, wall reflection on rows 0 and H-1,
no Zou-He.  The reference file is broken as shipped (its class body ends at validation.py:193); this twin
carries the methods the reference meant to have.  The reference's cos/sin direction vectors and cs**2 are
1-2 ulp off the exact values; the engine uses exact ones (difference ~1e-15).
"""
import numpy as np

try:
    from ._compute import ComputeBase, W9, run_loop, stream as _stream, wall_rows as _wall_rows
    from .. import geometry as _geo
except ImportError:
    from _compute import ComputeBase, W9, run_loop, stream as _stream, wall_rows as _wall_rows
    from fingering_dynamics_b200 import geometry as _geo

# ---- constants (validation.py:14-40) ------------------------------------------------------------------------
H = 200
W = 250
MAX_T = 1000
psi_wall = 0.0
Pe = 70
rho = 1.0
n_non = 1.4
M = 20.0
Theta = np.pi / 4
tau = 1 / (3 - np.sqrt(3))
C_W = 1.0 * (10.0 ** (-5)) / W
C_rho = 10.0 ** 3
v0 = (tau - 0.5) / 3
C_t = v0 / (10.0 ** (-6)) * (C_W ** 2)
DELTA_X = 1.0
DELTA_T = 1.0
sigma = 0.045 * (C_t ** 2) / (C_rho * (C_W ** 3))
u0 = C_t / C_W
c = DELTA_X / DELTA_T
cs = c / np.sqrt(3)
xi = 2.0 * DELTA_X
kappa = (3 / 4) * sigma * xi
a = 2 * kappa / (xi ** 2)
gamma = u0 * W / (a * Pe) / ((tau - 0.5) * DELTA_T)
Eta_n = 0.001 / (C_rho * (C_W ** 2) / C_t)


def create_circle(n, r):
    """validation.py:323-326"""
    y, x = np.ogrid[-int(W / 2): int(W / 2), -r: n - r]
    return x ** 2 + y ** 2 <= r ** 2


class Compute(ComputeBase):
    ZOU_HE, Y_WALL, X_PERIODIC, A_SIGN, F3 = "none", True, True, -1.0, 2 / 3
    _full_grid = True

    def __init__(self):
        """initial state of validation.py:43-96: droplet of radius 36 tangent to the bottom wall, u = 0"""
        self.mask = np.ones((H, W), dtype=bool)
        self.block_mask = np.zeros((H, W), dtype=bool)
        ang = [(i - 1) * np.pi / 2 for i in range(1, 5)]
        self.e = np.array([[0.0, 0.0]] + [[np.cos(t), np.sin(t)] for t in ang] +
                          [[np.cos((i - 5) * np.pi / 2 + np.pi / 4) * np.sqrt(2),
                            np.sin(np.pi * ((i - 5.0) / 2 + 1 / 4)) * np.sqrt(2)] for i in range(5, 9)]) * c
        self.e[np.abs(self.e) < 0.1] = 0
        self.w = W9.copy()
        self.psi = np.full((H, W), -1.0)
        self.psi[create_circle(W, 36).T[:H, :]] = 1.0
        self.gamma = gamma
        self.psi_wall_list = np.full((1, W), psi_wall).astype(float)
        self.rho = np.ones((H, W)) * rho
        self.ux = np.zeros((H, W))
        self.uy = np.zeros((H, W))
        self.f = np.zeros((9, H, W))
        self.g = np.zeros((9, H, W))
        self.nabla_psix, self.nabla_psiy, self.nabla_psi2 = self._stencils()
        self.mu = self.getMu()
        self.F = np.zeros((9, H, W))
        self.mix_tau = self.getMix_tau()
        self.p = self.getP()
        feq, geq, F = self._terms()
        self.feq, self.geq = feq.copy(), geq.copy()
        self.f, self.g = feq.copy(), geq.copy()

    def _fields(self, **extra):
        # validation.py keeps no gradient arrays up to date: getLarge_F evaluates getNabla_psix/psiy of the CURRENT psi
        # itself (validation.py:174-190), so the collision operators get fresh gradients, not the attributes
        gx, gy, _ = self._stencils()
        extra.setdefault("nabla_psix", gx)
        extra.setdefault("nabla_psiy", gy)
        return super()._fields(**extra)

    def power_law(self, temp2):
        return power_law(self, temp2)

    def updateP(self):
        self.p = self.getP()

    def updateRho(self):
        self.rho = self._moments()["rho"].copy()

    def updateMu(self):
        self.nabla_psi2 = self.getNabla_psi2()
        self.mu = self.getMu()

    def updateU(self):
        m = self._moments()
        self.ux, self.uy = m["ux"].copy(), m["uy"].copy()

    def updateF(self):
        self.f = self._collided("f")[0].copy()

    def updateG(self):
        self.g = self._collided("g")[1].copy()


def stream(f, g):
    _stream(f, g)


def halfway_bounceback(f_behind, g_behind, f, g):
    _wall_rows(f_behind, g_behind, f, g)


def update(i, x, y, cc):
    """animation callback of the reference (validation.py:379-384): draw psi frame i of cc"""
    print(i)
    import matplotlib.pyplot as plt
    plt.cla()
    plt.pcolor(x, y, cc[i], label="MAX_T:{}, Pe:{}, M:{}, wall{}".format(MAX_T, Pe, M, psi_wall))
    plt.legend()


# validation.py:193 ends `class Compute` early: the functions below are MODULE-level there (taking `self`), and callers
# bind them back onto the class.  Same names here, forwarding to the methods.
x_array = np.arange(0.1, 1.0, 0.01)  # validation.py:40


def power_law(self, temp2):
    """validation.py:219-225 (unused non-Newtonian relaxation time: nearest-neighbour inverse of
    x - temp2 x^(1-n) - dt/2 on x_array)"""
    from scipy import interpolate
    y_array = x_array - temp2 * x_array ** (1 - n_non) - 0.5 * DELTA_T
    return interpolate.interp1d(y_array.real, x_array, kind='nearest', fill_value='extrapolate')(0)


def getMix_tau(self):
    return Compute.getMix_tau(self)


def updatePsi(self):
    return Compute.updatePsi(self)


def getNabla_psix(self):
    return Compute.getNabla_psix(self)


def getNabla_psiy(self):
    return Compute.getNabla_psiy(self)


def getNabla_psi2(self):
    return Compute.getNabla_psi2(self)


def updateF(self):
    return Compute.updateF(self)


def updateG(self):
    return Compute.updateG(self)


def main(max_t=None, show=True):
    cm = Compute()
    run_loop(cm, _geo.reflect_bits_wall_rows(H, W, 0, H - 1), MAX_T if max_t is None else max_t)
    if show:
        try:
            import matplotlib.pyplot as plt
            plt.figure()
            plt.pcolor(list(range(W)), list(range(H)), cm.psi)
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
