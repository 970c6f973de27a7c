"""Error behaviour of the C ABI (include/fdlbm.h): every misuse returns a negative FDLBM_E_* code with a message in
fdlbm_last_error() and never throws or crashes; the Python binding turns codes into FdlbmError / ValueError.  The
reference's own error behaviour is Python exceptions (IndexError from Createblock, FloatingPointError under
np.seterr(all='raise'), fingering_periodic.py:497); those twins are covered in test_create_block / test_gpu_parity.

Argument validation happens before the device is touched, so the first half runs on CPU; the state-machine half
needs a GPU."""
import ctypes

import numpy as np
import pytest

from tests import helpers as hp

E_ARG, E_CUDA, E_STATE = -1, -2, -3


def _cfg(**kw):
    from fingering_dynamics_b200 import _native as nat
    c = nat.Config()
    c.H, c.W, c.x0, c.x1 = 32, 48, 0, 48
    c.x_periodic = 1
    c.tau, c.gamma, c.a, c.kappa, c.Eta_n, c.M = 0.8, 1.0, -0.04, 0.09, 0.1, 20.0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _create(cfg):
    from fingering_dynamics_b200 import _native as nat
    h = ctypes.c_void_p()
    rc = nat.lib().fdlbm_create(ctypes.byref(cfg), ctypes.byref(h))
    msg = nat.lib().fdlbm_last_error().decode()
    if rc == 0:
        nat.lib().fdlbm_destroy(h)
    return rc, msg


@pytest.mark.parametrize("kw,needle", [
    (dict(H=3), "grid too small"),
    (dict(W=2, x1=2), "grid too small"),
    (dict(dtype=7), "bad dtype"),
    (dict(x0=-1), "bad slab"),
    (dict(x1=49), "bad slab"),
    (dict(x0=47, external_halo=1), "bad slab"),
    (dict(zou_he=1), "mutually exclusive"),
    (dict(x_periodic=0, zou_he=0), "Zou-He"),
    (dict(x_periodic=0, zou_he=1), "inlet_ux"),
    (dict(x0=8), "external_halo"),
    (dict(tau=0.0), "tau"),
    (dict(tau=float("nan")), "tau"),
])
def test_create_rejects_bad_configurations_before_touching_the_device(kw, needle):
    rc, msg = _create(_cfg(**kw))
    assert rc == E_ARG, (rc, msg)
    assert needle in msg, msg


def test_null_arguments_are_refused_everywhere():
    from fingering_dynamics_b200 import _native as nat
    L = nat.lib()
    assert L.fdlbm_create(None, None) == E_ARG
    assert L.fdlbm_set_geometry(None, 0, 1, None, None) == E_ARG
    assert L.fdlbm_set_state(None, 0, 1, None) == E_ARG
    assert L.fdlbm_get_state(None, 0, 1, None) == E_ARG
    assert L.fdlbm_step(None, 1) == E_ARG
    assert L.fdlbm_sync(None) == E_ARG
    assert L.fdlbm_halo_regions(None, None) == E_ARG
    assert L.fdlbm_checkpoint_save(None, None, 0) == E_ARG
    assert L.fdlbm_checkpoint_load(None, None, 0) == E_ARG
    assert L.fdlbm_count_nonfinite(None, None) == E_ARG
    assert L.fdlbm_peer_export(None, None) == E_ARG
    assert L.fdlbm_peer_attach(None, 0, None) == E_ARG
    assert L.fdlbm_iterations(None) == -1 and L.fdlbm_launch_count(None) == -1
    assert L.fdlbm_checkpoint_bytes(None) == 0 and L.fdlbm_stream(None) is None
    L.fdlbm_destroy(None)  # a no-op
    assert b"null" in L.fdlbm_last_error() or b"bad argument" in L.fdlbm_last_error()


def test_python_binding_validates_shapes_without_a_device():
    from fingering_dynamics_b200 import _native as nat
    with pytest.raises(ValueError):
        nat.as_f64(np.zeros((3, 4)), (4, 3))
    from fingering_dynamics_b200 import Engine
    with pytest.raises(ValueError):   # Zou-He faces without profiles: refused before the library is called
        Engine(32, 48, tau=0.8, gamma=1.0, a=-0.04, kappa=0.09, Eta_n=0.1, M=20.0, psi_wall=0.0, zou_he="fp")
    with pytest.raises(KeyError):
        Engine(32, 48, tau=0.8, gamma=1.0, a=-0.04, kappa=0.09, Eta_n=0.1, M=20.0, psi_wall=0.0, x_periodic=True,
               dtype="f16")


@pytest.mark.gpu
def test_state_machine_errors_leave_the_engine_usable(golden):
    from fingering_dynamics_b200 import _native as nat
    L = nat.lib()
    d = golden("fp_small")
    H, W = int(d["H"]), int(d["W"])
    rc, msg = _create(_cfg(device=4096))
    assert rc == E_ARG and "bad device" in msg
    e = hp.fp_engine(d)
    F = nat.Fields()
    # nothing loaded yet
    assert L.fdlbm_step(e._h, 1) == E_STATE and b"set_state" in L.fdlbm_last_error()
    out = np.zeros((H, W))
    F.psi = out.ctypes.data
    assert L.fdlbm_get_state(e._h, 0, W, ctypes.byref(F)) == E_STATE
    n = ctypes.c_int64(0)
    assert L.fdlbm_count_nonfinite(e._h, ctypes.byref(n)) == E_STATE
    buf = np.zeros(int(L.fdlbm_checkpoint_bytes(e._h)), dtype=np.uint8)
    assert L.fdlbm_checkpoint_save(e._h, nat.ptr(buf), buf.size) == E_STATE
    # bad arguments
    st = hp.state_for_engine(d, "s0")
    with pytest.raises(ValueError):
        e.set_state(**{k: v for k, v in st.items() if k != "mu"})
    with pytest.raises(ValueError):
        e.set_state(**dict(st, psi=st["psi"][:-1]))
    assert L.fdlbm_set_state(e._h, 0, W, ctypes.byref(nat.Fields())) == E_ARG
    s8 = np.zeros((H, W), dtype=np.uint8)
    assert L.fdlbm_set_geometry(e._h, 0, W + 1, nat.ptr(s8), nat.ptr(s8)) == E_ARG
    assert L.fdlbm_set_geometry(e._h, -1, 4, nat.ptr(s8), nat.ptr(s8)) == E_ARG
    assert L.fdlbm_step(e._h, -1) in (E_ARG, E_STATE)
    with pytest.raises(KeyError):
        e.get_state(("vorticity",))
    # ... and the engine still runs the reference trajectory afterwards
    e.set_state(**st)
    assert L.fdlbm_step(e._h, -3) == E_ARG
    assert L.fdlbm_get_state(e._h, 0, W + 5, ctypes.byref(F)) == E_ARG
    assert L.fdlbm_checkpoint_save(e._h, nat.ptr(buf), 16) == E_ARG
    assert L.fdlbm_checkpoint_load(e._h, nat.ptr(buf), 16) == E_ARG
    junk = np.zeros(buf.size, dtype=np.uint8)
    assert L.fdlbm_checkpoint_load(e._h, nat.ptr(junk), junk.size) == E_ARG and b"not a checkpoint" in L.fdlbm_last_error()
    info = nat.PeerInfo()
    assert L.fdlbm_peer_export(e._h, ctypes.byref(info)) == 0
    assert L.fdlbm_peer_attach(e._h, 0, ctypes.byref(info)) == E_ARG and b"external_halo" in L.fdlbm_last_error()
    assert L.fdlbm_peer_attach(e._h, 2, ctypes.byref(info)) == E_ARG
    e.step(10)
    got = e.get_state(("psi", "rho"))
    assert hp.rel_err(got["psi"], d["s10_psi"]) <= 1e-10 and hp.rel_err(got["rho"], d["s10_rho"]) <= 1e-10
    assert e.iterations == 10
    e.close()
    e.close()  # idempotent


@pytest.mark.gpu
def test_slab_engines_refuse_multi_step_calls_without_attached_neighbours(golden):
    from fingering_dynamics_b200 import _native as nat
    d = golden("va_small")
    H, W = int(d["H"]), int(d["W"])
    e = hp.va_engine(d, slab=(0, W // 2), external_halo=True)
    st = {k: (v[..., :W // 2] if v.shape[-1] == W else v) for k, v in hp.state_for_engine(d, "s0").items()}
    e.set_state(**{k: np.ascontiguousarray(v) for k, v in st.items()})
    with pytest.raises(nat.FdlbmError) as ei:
        e.step(2)
    assert "one step per call" in str(ei.value)
    for other in (hp.fp_engine(golden("fp_small"), dtype="f32"),   # different row pitch and dtype
                  hp.fg_engine(golden("fg_small"))):               # H = 36 against 48: same padded row pitch (64)
        info = nat.PeerInfo.from_buffer_copy(other.peer_export())
        assert nat.lib().fdlbm_peer_attach(e._h, 1, ctypes.byref(info)) == E_ARG
        assert b"different H" in nat.lib().fdlbm_last_error()
        other.close()
    e.close()
