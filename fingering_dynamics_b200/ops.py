"""NumPy-in / NumPy-out wrappers of the stateless C-ABI operators (fdlbm_op_*): the reference's
module-level functions one call at a time, computed on the GPU.  In-place semantics like the reference:
the caller's arrays (also non-contiguous views such as f[:, 1:-1]) are updated."""
import numpy as np

from . import _native as nat


def _writeback(dst, src):
    if src is not dst:
        dst[...] = src


def stream(f, g):
    """stream(f, g), fingering_periodic.py:327-343: every population shifted by e_i, periodic, all cells."""
    _, H, W = f.shape
    fo, go = nat.as_f64(f, (9, H, W)), nat.as_f64(g, (9, H, W))
    nat.check(nat.lib().fdlbm_op_stream(H, W, nat.ptr(fo), nat.ptr(go)))
    _writeback(f, fo)
    _writeback(g, go)


def bounce_back(reflect, f_behind, g_behind, f, g):
    """f_i(c) <- f_behind_opp(i)(c) wherever bit (i-1) of reflect[c] is set (and the same for g)."""
    _, H, W = f.shape
    fb, gb = nat.as_f64(f_behind, (9, H, W)), nat.as_f64(g_behind, (9, H, W))
    fo, go = nat.as_f64(f, (9, H, W)), nat.as_f64(g, (9, H, W))
    r = np.ascontiguousarray(reflect, dtype=np.uint8)
    if r.shape != (H, W):
        raise ValueError("reflect bits must be (H, W)")
    nat.check(nat.lib().fdlbm_op_bounce_back(H, W, nat.ptr(r), nat.ptr(fb), nat.ptr(gb), nat.ptr(fo), nat.ptr(go)))
    _writeback(f, fo)
    _writeback(g, go)
