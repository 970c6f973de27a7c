"""Build and load libfdlbm.so (the C ABI of include/fdlbm.h) with ctypes.

There is no CPU path: if the library cannot be loaded, or no CUDA device is visible when an engine
is created, the call raises.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libfdlbm.so")
SOURCES = ["fdlbm.cu", "lbm_device.cuh", "lbm_kernels.cuh", "lbm_fused.cuh", "lbm_fused_vec.cuh", "lbm_fused_f32.cuh", "lbm_ops.cuh", "lbm_init.cuh"]
HEADER = os.path.join(os.path.dirname(_HERE), "include", "fdlbm.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC"]


STAMP_PATH = LIB_PATH + ".srchash"


def _source_hash():
    """sha256 over the sources, the header and the flags: what the library was built from (mtimes do not survive
    a checkout or the copy to the GPU box)"""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in [os.path.join(CSRC, s) for s in SOURCES] + [HEADER]:
        h.update(b"\0" + os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(STAMP_PATH) as fh:
            return fh.read().strip() != _source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """nvcc-compile the CUDA library for sm_100a, in tree (cross-compiles without a GPU).  Safe under torchrun:
    one process compiles (file lock) into a temporary file that is renamed into place, so no rank ever dlopens
    a half-written library."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():   # another process built it while we waited
                return LIB_PATH
            nvcc = os.environ.get("NVCC", "nvcc")
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "fdlbm.cu")]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.unlink(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stderr)
            if verbose:
                print(r.stderr)
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH + ".tmp", "w") as fh:
                fh.write(_source_hash())
            os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


class Config(ctypes.Structure):
    """fdlbm_config of include/fdlbm.h"""
    _fields_ = [("H", ctypes.c_int32), ("W", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("psi_y_wall", ctypes.c_int32), ("x_periodic", ctypes.c_int32), ("zou_he", ctypes.c_int32),
                ("kernel", ctypes.c_int32), ("device", ctypes.c_int32), ("x0", ctypes.c_int32),
                ("x1", ctypes.c_int32), ("external_halo", ctypes.c_int32),
                ("tau", ctypes.c_double), ("gamma", ctypes.c_double), ("a", ctypes.c_double),
                ("kappa", ctypes.c_double), ("Eta_n", ctypes.c_double), ("M", ctypes.c_double),
                ("psi_wall", ctypes.c_double), ("psi_left", ctypes.c_double), ("psi_right", ctypes.c_double),
                ("outlet_f3_coef", ctypes.c_double),
                ("inlet_ux", ctypes.c_void_p), ("outlet_ux", ctypes.c_void_p)]


FIELD_NAMES = ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "nabla_psix", "nabla_psiy", "nabla_psi2")


class Fields(ctypes.Structure):
    """fdlbm_fields of include/fdlbm.h"""
    _fields_ = [(n, ctypes.c_void_p) for n in FIELD_NAMES]


class Init(ctypes.Structure):
    """fdlbm_init of include/fdlbm.h"""
    _fields_ = [("variant", ctypes.c_int32), ("n_inject", ctypes.c_int32), ("psi_inject", ctypes.c_double),
                ("psi_rest", ctypes.c_double), ("rho0", ctypes.c_double), ("rho", ctypes.c_void_p),
                ("col0", ctypes.c_int32), ("ncols", ctypes.c_int32)]


class Halo(ctypes.Structure):
    _fields_ = [("send_lo", ctypes.c_void_p), ("recv_lo", ctypes.c_void_p), ("send_hi", ctypes.c_void_p),
                ("recv_hi", ctypes.c_void_p), ("bytes", ctypes.c_size_t)]


class AlgebraOut(ctypes.Structure):
    """fdlbm_algebra_out of include/fdlbm.h"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("p", "mu", "mix_tau", "a0", "a1_8", "b0", "b1_8")]


class PeerInfo(ctypes.Structure):
    """fdlbm_peer_info of include/fdlbm.h (plain bytes: can be pickled and sent to another rank)"""
    _fields_ = [("pid", ctypes.c_int64), ("device", ctypes.c_int32), ("Wl", ctypes.c_int32), ("Hp", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("H", ctypes.c_int32), ("lat", ctypes.c_void_p * 2), ("flags", ctypes.c_void_p),
                ("ipc_lat", (ctypes.c_ubyte * 64) * 2), ("ipc_flags", ctypes.c_ubyte * 64)]


# every symbol include/fdlbm.h declares (tests check the export list against the header)
_SIGS = {
    "fdlbm_abi_version": (ctypes.c_int, []),
    "fdlbm_last_error": (ctypes.c_char_p, []),
    "fdlbm_device_count": (ctypes.c_int, []),
    "fdlbm_create": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(ctypes.c_void_p)]),
    "fdlbm_destroy": (None, [ctypes.c_void_p]),
    "fdlbm_set_geometry": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "fdlbm_set_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Fields)]),
    "fdlbm_init_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Init)]),
    "fdlbm_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "fdlbm_get_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Fields)]),
    "fdlbm_iterations": (ctypes.c_int64, [ctypes.c_void_p]),
    "fdlbm_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "fdlbm_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "fdlbm_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "fdlbm_halo_regions": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Halo)]),
    "fdlbm_checkpoint_bytes": (ctypes.c_size_t, [ctypes.c_void_p]),
    "fdlbm_checkpoint_save": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "fdlbm_checkpoint_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "fdlbm_count_nonfinite": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "fdlbm_peer_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(PeerInfo)]),
    "fdlbm_peer_attach": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(PeerInfo)]),
    "fdlbm_pinned_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "fdlbm_pinned_free": (None, [ctypes.c_void_p]),
    "fdlbm_op_release": (None, []),
    "fdlbm_op_stream": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "fdlbm_op_bounce_back": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "fdlbm_op_stencils": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "fdlbm_op_collide": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.c_void_p, ctypes.POINTER(Fields)]),
    "fdlbm_op_collision_terms": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.c_void_p, ctypes.POINTER(Fields),
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "fdlbm_op_algebra": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(Fields), ctypes.POINTER(AlgebraOut)]),
    "fdlbm_op_zou_he": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(Fields)]),
    "fdlbm_op_moments": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.c_void_p, ctypes.POINTER(Fields)]),
}

_lib = None


def lib():
    """The loaded CUDA library.  Raises if it is missing and cannot be built: there is no fallback."""
    global _lib
    if _lib is None:
        # FDLBM_LIB: load a specific build of the same sources (kernel tuning experiments)
        path = os.environ.get("FDLBM_LIB") or build()
        L = ctypes.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.fdlbm_abi_version() != 1:
            raise RuntimeError("libfdlbm.so ABI version mismatch")
        _lib = L
    return _lib


class FdlbmError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise FdlbmError("fdlbm error %d: %s" % (rc, lib().fdlbm_last_error().decode()))


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)
