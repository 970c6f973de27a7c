"""Engine: the Python face of the C ABI (include/fdlbm.h).  One Engine replaces the body of the
reference's `for i in range(MAX_T)` loop (fingering_periodic.py:454-479, fingering.py:558-585,
validation.py:392-409): the state lives on the GPU and `step(n)` advances it n reference iterations.

NumPy arrays in, NumPy arrays out, reference layout ((9,H,W) populations, (H,W) fields, float64).
"""
import ctypes

import numpy as np

from . import _native as nat

_ZH = {"none": 0, "fp": 1, "fg": 2}
_DT = {"f64": 0, "f32": 1}
_KERNEL = {"auto": 0, "twopass": 1, "fused": 2}

STATE_IN = ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "nabla_psix", "nabla_psiy")


class Engine:
    def __init__(self, H, W, *, tau, gamma, a, kappa, Eta_n, M, psi_wall, psi_left=1.0, psi_right=-1.0,
                 psi_y_wall=False, x_periodic=False, zou_he="none", inlet_ux=None, outlet_ux=None,
                 outlet_f3_coef=2.0 / 3.0, dtype="f64", kernel="auto", device=0, slab=None, external_halo=False):
        self._h = None
        self.H, self.W = int(H), int(W)
        self.x0, self.x1 = (0, self.W) if slab is None else (int(slab[0]), int(slab[1]))
        cfg = nat.Config()
        cfg.H, cfg.W = self.H, self.W
        cfg.dtype = _DT[dtype]
        cfg.psi_y_wall = int(bool(psi_y_wall))
        cfg.x_periodic = int(bool(x_periodic))
        cfg.zou_he = _ZH[zou_he]
        cfg.kernel = _KERNEL[kernel]
        cfg.device = int(device)
        cfg.x0, cfg.x1 = self.x0, self.x1
        cfg.external_halo = int(bool(external_halo))
        cfg.tau, cfg.gamma, cfg.a, cfg.kappa = float(tau), float(gamma), float(a), float(kappa)
        cfg.Eta_n, cfg.M, cfg.psi_wall = float(Eta_n), float(M), float(psi_wall)
        cfg.psi_left, cfg.psi_right = float(psi_left), float(psi_right)
        cfg.outlet_f3_coef = float(outlet_f3_coef)
        self._profiles = []
        if cfg.zou_he:
            if inlet_ux is None or outlet_ux is None:
                raise ValueError("Zou-He faces need inlet_ux and outlet_ux (H values each)")
            for name, arr in (("inlet_ux", inlet_ux), ("outlet_ux", outlet_ux)):
                a64 = nat.as_f64(np.broadcast_to(np.asarray(arr, dtype=np.float64), (self.H,)), (self.H,))
                self._profiles.append(a64)
                setattr(cfg, name, a64.ctypes.data)
        self.cfg = cfg
        self.dtype = dtype
        h = ctypes.c_void_p()
        nat.check(nat.lib().fdlbm_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h

    # -- life cycle ---------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            nat.lib().fdlbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs -------------------------------------------------------------------------------
    def _window(self, arr2d_shape, col0):
        H, ncols = arr2d_shape
        if H != self.H:
            raise ValueError("array has %d rows, engine has H=%d" % (H, self.H))
        return int(col0), int(ncols)

    def set_geometry(self, solid, reflect, col0=0):
        """solid: (H,ncols) nonzero = block cell; reflect: (H,ncols) uint8 bits (see geometry.py)."""
        s = np.asarray(solid)
        if not (s.dtype == np.uint8 and s.flags.c_contiguous):   # uint8 planes go up as they are (page-locked ones too):
            s = np.ascontiguousarray(s != 0, dtype=np.uint8)      # the device packs "nonzero" into the bitfield
        r = np.ascontiguousarray(reflect, dtype=np.uint8)
        if s.shape != r.shape:
            raise ValueError("solid and reflect shapes differ")
        c0, nc = self._window(s.shape, col0)
        nat.check(nat.lib().fdlbm_set_geometry(self._h, c0, nc, nat.ptr(s), nat.ptr(r)))

    def set_state(self, col0=0, **arrays):
        """f, g (9,H,ncols) and psi, rho, ux, uy, p, mu, mix_tau, nabla_psix, nabla_psiy (H,ncols);
        nabla_psi2 optional.  These are the arrays the reference's Compute object holds before an
        iteration; the first collision uses them as given."""
        missing = [k for k in STATE_IN if k not in arrays]
        if missing:
            raise ValueError("set_state is missing " + ", ".join(missing))
        keep = {}
        F = nat.Fields()
        nc = None
        for k in nat.FIELD_NAMES:
            if arrays.get(k) is None:
                continue
            a = nat.as_f64(arrays[k])
            shape2 = a.shape[1:] if k in ("f", "g") else a.shape
            if k in ("f", "g") and a.shape[0] != 9:
                raise ValueError("%s must have 9 populations" % k)
            c0, n = self._window(shape2, col0)
            if nc is not None and n != nc:
                raise ValueError("inconsistent widths")
            nc = n
            keep[k] = a
            setattr(F, k, a.ctypes.data)
        nat.check(nat.lib().fdlbm_set_state(self._h, int(col0), nc, ctypes.byref(F)))

    def init_state(self, variant="fp", n_inject=5, psi_inject=1.0, psi_rest=-1.0, rho0=1.0, rho=None, col0=0):
        """Compute.__init__ on the device (fingering_periodic.py:90-121 / fingering.py:95-127): no host planes
        except an optional (H,ncols) `rho` (fingering.py draws it at random).  Leaves the engine where
        set_state would; fp64 engines hold the reference's bits."""
        spec = nat.Init()
        spec.variant = {"fp": 1, "fg": 2}[variant]
        spec.n_inject = int(n_inject)
        spec.psi_inject, spec.psi_rest, spec.rho0 = float(psi_inject), float(psi_rest), float(rho0)
        keep = None
        if rho is not None:
            keep = nat.as_f64(rho)
            _, spec.ncols = self._window(keep.shape, col0)
            spec.col0 = int(col0)
            spec.rho = keep.ctypes.data
        nat.check(nat.lib().fdlbm_init_state(self._h, ctypes.byref(spec)))

    # -- run ----------------------------------------------------------------------------------
    def step(self, n=1):
        nat.check(nat.lib().fdlbm_step(self._h, int(n)))

    def sync(self):
        nat.check(nat.lib().fdlbm_sync(self._h))

    @property
    def iterations(self):
        return int(nat.lib().fdlbm_iterations(self._h))

    @property
    def launch_count(self):
        return int(nat.lib().fdlbm_launch_count(self._h))

    @property
    def stream(self):
        """raw cudaStream_t of the engine (int)"""
        return int(nat.lib().fdlbm_stream(self._h) or 0)

    # -- restart / watchdog ---------------------------------------------------------------------------
    def checkpoint(self):
        """opaque bytes of the raw device state; a run continued from restore() is bit-identical"""
        n = int(nat.lib().fdlbm_checkpoint_bytes(self._h))
        buf = np.empty(n, dtype=np.uint8)
        nat.check(nat.lib().fdlbm_checkpoint_save(self._h, nat.ptr(buf), n))
        return buf

    def restore(self, blob):
        """load a checkpoint() blob taken from an engine of the same grid / slab / dtype (geometry must be set)"""
        buf = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
        nat.check(nat.lib().fdlbm_checkpoint_load(self._h, nat.ptr(buf), buf.size))

    def save(self, path):
        """checkpoint() to exactly `path` (np.save on a file name would append '.npy')"""
        with open(path, "wb") as fh:
            np.save(fh, self.checkpoint(), allow_pickle=False)

    def load(self, path):
        with open(path, "rb") as fh:
            self.restore(np.load(fh, allow_pickle=False))

    def count_nonfinite(self):
        n = ctypes.c_int64(0)
        nat.check(nat.lib().fdlbm_count_nonfinite(self._h, ctypes.byref(n)))
        return int(n.value)

    def check_finite(self):
        """the reference runs under np.seterr(all='raise') (fingering_periodic.py:497): a blow-up is a
        FloatingPointError there; same here, on demand"""
        n = self.count_nonfinite()
        if n:
            raise FloatingPointError("%d non-finite populations after %d iterations" % (n, self.iterations))

    def peer_export(self):
        """bytes describing this engine's lattices and flag words (fdlbm_peer_info) for a neighbouring engine"""
        info = nat.PeerInfo()
        nat.check(nat.lib().fdlbm_peer_export(self._h, ctypes.byref(info)))
        return bytes(info)

    def peer_attach(self, side, info_bytes):
        """side 0: the neighbour owns the columns just below x0; side 1: just above x1.  Afterwards the step
        kernel pushes its edge columns into that neighbour's ghost columns and step(n) may take many steps."""
        info = nat.PeerInfo.from_buffer_copy(info_bytes)
        nat.check(nat.lib().fdlbm_peer_attach(self._h, int(side), ctypes.byref(info)))

    def halo_regions(self):
        h = nat.Halo()
        nat.check(nat.lib().fdlbm_halo_regions(self._h, ctypes.byref(h)))
        return h

    # -- outputs ------------------------------------------------------------------------------
    def get_state(self, names=("psi", "rho", "ux", "uy"), col0=None, ncols=None, out=None):
        """What the reference holds after the iterations done so far.  Returns {name: ndarray}."""
        col0 = self.x0 if col0 is None else int(col0)
        ncols = (self.x1 - self.x0) if ncols is None else int(ncols)
        res = {} if out is None else out
        F = nat.Fields()
        for k in names:
            if k not in nat.FIELD_NAMES:
                raise KeyError(k)
            shape = (9, self.H, ncols) if k in ("f", "g") else (self.H, ncols)
            if k not in res:
                res[k] = np.zeros(shape, dtype=np.float64)
            a = res[k]
            if a.shape != shape or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError("bad output array for " + k)
            setattr(F, k, a.ctypes.data)
        nat.check(nat.lib().fdlbm_get_state(self._h, col0, ncols, ctypes.byref(F)))
        return res


def pinned_empty(shape, dtype=np.float64):
    """NumPy array over page-locked host memory (faster H2D/D2H for the e2e path)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = nat.lib().fdlbm_pinned_alloc(max(n, 1))
    if not p:
        raise MemoryError(nat.lib().fdlbm_last_error().decode())
    buf = (ctypes.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        nat.lib().fdlbm_pinned_free(p)
