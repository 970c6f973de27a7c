"""ctypes binding of the CPU oracle (oracle/fd_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of fd_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by the
package `fingering_dynamics_b200`.

Parity status: the reference has no tests or golden vectors of its own ("parity unpinned" by the
reference); this oracle is pinned bit-for-bit against tests/golden/*.npz, which were produced by
executing the unmodified reference (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fd_oracle.c")
_LIB = os.path.join(_HERE, "libfd_oracle.so")

_BASE = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden"]


def build(force=False):
    """Compile fd_oracle.c -> libfd_oracle.so (OpenMP when the toolchain has it)."""
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    tries = [
        ["gcc"] + _BASE + ["-fopenmp", "-B/usr/lib/gcc/x86_64-linux-gnu/13/"],
        ["gcc"] + _BASE + ["-fopenmp"],
        ["gcc"] + _BASE + ["-Wno-unknown-pragmas"],
    ]
    err = None
    for cmd in tries:
        r = subprocess.run(cmd + ["-o", _LIB, _SRC, "-lm"], capture_output=True, text=True)
        if r.returncode == 0:
            return _LIB
        err = r.stderr
    raise RuntimeError("cannot build the oracle: " + str(err))


class Params(ctypes.Structure):
    _fields_ = [("H", ctypes.c_int), ("W", ctypes.c_int),
                ("tau", ctypes.c_double), ("gamma", ctypes.c_double), ("a", ctypes.c_double),
                ("kappa", ctypes.c_double), ("Eta_n", ctypes.c_double), ("M", ctypes.c_double),
                ("psi_wall", ctypes.c_double), ("psi_left", ctypes.c_double), ("psi_right", ctypes.c_double),
                ("x_periodic", ctypes.c_int), ("y_wall", ctypes.c_int), ("lap_order", ctypes.c_int),
                ("outlet_f3_coef", ctypes.c_double)]


_dp = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_ip = ctypes.POINTER(ctypes.c_int)


class State(ctypes.Structure):
    _fields_ = [(n, _dp) for n in ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "gx", "gy", "lap")]


class Geom(ctypes.Structure):
    _fields_ = [("mask", _u8p), ("circ_masks", _u8p), ("rect_corners", _ip), ("n_rects", ctypes.c_int),
                ("wall_lo", ctypes.c_int), ("wall_hi", ctypes.c_int), ("zou_he", ctypes.c_int),
                ("inlet_ux", _dp), ("outlet_ux", _dp)]


class VaConsts(ctypes.Structure):
    _fields_ = [("e", (ctypes.c_double * 2) * 9), ("w", ctypes.c_double * 9),
                ("cs2", ctypes.c_double), ("cs4", ctypes.c_double), ("a_va", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fdo_abi_version.restype = ctypes.c_int
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _u8(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_u8p)


def set_threads(n):
    """OMP thread count for the next calls (no-op when built without OpenMP)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
    except OSError:
        pass


def make_params(H, W, *, tau, gamma, a, kappa, Eta_n, M, psi_wall, psi_left=1.0, psi_right=-1.0,
                x_periodic=0, y_wall=0, lap_order=0, outlet_f3_coef=2 / 3):
    return Params(H, W, tau, gamma, a, kappa, Eta_n, M, psi_wall, psi_left, psi_right,
                  x_periodic, y_wall, lap_order, outlet_f3_coef)


def stencils(P, psi):
    gx, gy, lap = (np.empty_like(psi) for _ in range(3))
    lib().fdo_stencils(ctypes.byref(P), _d(psi), _d(gx), _d(gy), _d(lap))
    return gx, gy, lap


def stream(f, g):
    _, H, W = f.shape
    lib().fdo_stream(H, W, _d(f), _d(g))


def bb_circle(masks12, fb, gb, f, g):
    _, H, W = f.shape
    m = np.ascontiguousarray(np.asarray(masks12, dtype=np.uint8))
    lib().fdo_bb_circle(H, W, _u8(m), _d(fb), _d(gb), _d(f), _d(g))


def bb_rect(corners, fb, gb, f, g):
    _, H, W = f.shape
    c = np.ascontiguousarray(np.asarray(corners, dtype=np.int32))
    lib().fdo_bb_rect(H, W, c.ctypes.data_as(_ip), c.shape[0], _d(fb), _d(gb), _d(f), _d(g))


def wall_rows(lo, hi, fb, gb, f, g):
    _, H, W = f.shape
    lib().fdo_wall_rows(H, W, lo, hi, _d(fb), _d(gb), _d(f), _d(g))


def left_boundary(hole, fb, gb, f, g):
    _, H, W = f.shape
    lib().fdo_left_boundary(H, W, hole, _d(fb), _d(gb), _d(f), _d(g))


class Run:
    """Holds the arrays of one simulation and advances them with fdo_iterate / fdo_va_iterate."""

    FIELDS = ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "gx", "gy", "lap")

    def __init__(self, P, arrays, *, mask=None, circ_masks=None, rect_corners=None, wall_rows=None,
                 zou_he=0, inlet_ux=None, outlet_ux=None, va=None):
        self.P = P
        self.a = {k: np.ascontiguousarray(np.array(arrays[k], dtype=np.float64)) for k in self.FIELDS}
        self.S = State(*[_d(self.a[k]) for k in self.FIELDS])
        H, W = P.H, P.W
        self.mask = np.ascontiguousarray((np.ones((H, W)) if mask is None else mask).astype(np.uint8))
        self.circ = None if circ_masks is None else np.ascontiguousarray(np.asarray(circ_masks, dtype=np.uint8))
        self.rect = None if rect_corners is None else np.ascontiguousarray(np.asarray(rect_corners, dtype=np.int32))
        self.inlet = None if inlet_ux is None else np.ascontiguousarray(inlet_ux, dtype=np.float64)
        self.outlet = None if outlet_ux is None else np.ascontiguousarray(outlet_ux, dtype=np.float64)
        lo, hi = (-1, -1) if wall_rows is None else wall_rows
        self.G = Geom(_u8(self.mask),
                      None if self.circ is None else _u8(self.circ),
                      None if self.rect is None else self.rect.ctypes.data_as(_ip),
                      0 if self.rect is None else self.rect.shape[0],
                      lo, hi, zou_he,
                      None if self.inlet is None else _d(self.inlet),
                      None if self.outlet is None else _d(self.outlet))
        self.va = va

    def iterate(self, n):
        if self.va is not None:
            lib().fdo_va_iterate(ctypes.byref(self.P), ctypes.byref(self.va), ctypes.byref(self.S), int(n))
        else:
            lib().fdo_iterate(ctypes.byref(self.P), ctypes.byref(self.G), ctypes.byref(self.S), int(n))
        return self.a


def va_consts(e, w, cs, a_va):
    v = VaConsts()
    for i in range(9):
        v.e[i][0], v.e[i][1] = float(e[i][0]), float(e[i][1])
        v.w[i] = float(w[i])
    cs = np.float64(cs)
    v.cs2, v.cs4, v.a_va = float(cs ** 2), float(cs ** 4), float(a_va)
    return v


# ------------------------------------------------------------------------------------------------
# initial states (restating Compute.__init__ of each driver) -- used by tests and the CPU bench arm
# ------------------------------------------------------------------------------------------------
def equilibrium(P, mask, rho, ux, uy, p, mu, psi):
    H, W = P.H, P.W
    f = np.zeros((9, H, W))
    g = np.zeros((9, H, W))
    m = np.ascontiguousarray(mask.astype(np.uint8))
    lib().fdo_equilibrium(ctypes.byref(P), _u8(m), _d(f), _d(g), _d(rho), _d(ux), _d(uy), _d(p), _d(mu), _d(psi))
    return f, g


def _mix_tau(P, rho, psi):
    v1 = P.Eta_n / rho
    v2 = P.Eta_n * P.M / rho
    return 3 * (2 * v1 * v2 / (v1 * (1.0 - psi) + v2 * (1.0 + psi))) + 0.5


def fp_initial_state(P, mask, n_inject=5):
    """fingering_periodic.py:90-121: psi=-1, first 5 columns +1, solids psi_wall; rho=1, u=0, mu=0 (!),
    p = rho/3 + psi*0, tau_mix from rho, psi; f=f_eq, g=g_eq on fluid."""
    H, W = P.H, P.W
    mask = np.asarray(mask, dtype=bool)
    psi = np.full((H, W), -1.0)
    psi[:, :n_inject] = 1.0
    psi[~mask] = P.psi_wall
    rho = np.ones((H, W))
    z = np.zeros((H, W))
    gx, gy, lap = stencils(P, psi)
    mu = z.copy()
    p = 1 / 3 * rho + psi * mu
    mt = _mix_tau(P, rho, psi)
    f, g = equilibrium(P, mask, rho, z, z, p, mu, psi)
    return dict(f=f, g=g, psi=psi, rho=rho, ux=z.copy(), uy=z.copy(), p=p, mu=mu, mix_tau=mt, gx=gx, gy=gy, lap=lap)


def fg_initial_state(P, mask, rho, n_inject=5):
    """fingering.py:95-127: as FP but rho given (random in the reference), p computed with mu=0, THEN mu,
    THEN uy from mu*nabla_psiy/2/rho (ux stays 0), tau_mix, f=f_eq, g=g_eq."""
    H, W = P.H, P.W
    mask = np.asarray(mask, dtype=bool)
    psi = np.full((H, W), -1.0)
    psi[:, :n_inject] = 1.0
    psi[~mask] = P.psi_wall
    z = np.zeros((H, W))
    gx, gy, lap = stencils(P, psi)
    p = 1 / 3 * rho + psi * z
    mu = P.a * psi * (1.0 - psi * psi) - P.kappa * lap
    uy = (z + mu * gy / 2) / rho
    mt = _mix_tau(P, rho, psi)
    f, g = equilibrium(P, mask, rho, z, uy, p, mu, psi)
    return dict(f=f, g=g, psi=psi, rho=rho, ux=z.copy(), uy=uy, p=p, mu=mu, mix_tau=mt, gx=gx, gy=gy, lap=lap)
