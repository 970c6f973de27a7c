"""Compute.__init__ on the device (fdlbm_init_state): the analytic initial state of
fingering_periodic.py:90-121 / fingering.py:95-127 built by a kernel instead of 29 uploaded host planes.

fp64: bit-identical to the oracle's restatement of Compute.__init__ AND to the fixtures written by the unmodified
reference (tests/golden/make_golden.py); a run started from it is the run started from fdlbm_set_state.
"""
import numpy as np
import pytest

from tests import helpers as hp

pytestmark = pytest.mark.gpu

MAC = ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "nabla_psix", "nabla_psiy", "nabla_psi2")


def _porous(H, W, dtype="f64", **kw):
    from fingering_dynamics_b200 import Engine, synthetic as syn
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
               psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype=dtype, **kw)
    e.set_geometry(solid, refl)
    return e, c, solid, refl


def test_fp_device_init_is_the_oracles_initial_state_bit_for_bit():
    from oracle import oracle as orc
    H, W = 256, 192
    e, c, solid, _ = _porous(H, W)
    e.init_state("fp", rho0=c["rho0"])
    got = e.get_state(MAC)
    e.close()
    P = orc.make_params(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                        psi_wall=c["psi_wall"])
    ref = orc.fp_initial_state(P, solid == 0)
    fluid = solid == 0
    names = {"nabla_psix": "gx", "nabla_psiy": "gy", "nabla_psi2": "lap"}
    for k in MAC:
        r = ref[names.get(k, k)]
        if k not in ("f", "g", "psi", "nabla_psix", "nabla_psiy", "nabla_psi2"):
            r = np.where(fluid, r, 0.0)   # masked 1-D arrays of the reference: no entry at solids
        assert np.array_equal(got[k], r), k


@pytest.mark.parametrize("name", ["fp_small", "fg_small"])
def test_device_init_reproduces_the_reference_fixture_s0(golden, name):
    """s0 of the fixtures is Compute.__init__ of the UNMODIFIED reference (FG: with its random rho handed in)"""
    d = golden(name)
    e = hp.ENGINES[name](d)
    mask = d["mask"]
    if name == "fp_small":
        e.init_state("fp")
    else:
        e.init_state("fg", rho=d["s0_rho"])
    got = e.get_state(MAC)
    e.close()
    for k in MAC:
        r = d["s0_" + k]
        if k not in ("f", "g", "psi", "nabla_psix", "nabla_psiy", "nabla_psi2"):
            r = np.where(mask, r, 0.0)
        assert np.array_equal(got[k], r), (name, k, float(np.max(np.abs(got[k] - r))))


@pytest.mark.parametrize("name", ["fp_small", "fg_small"])
def test_run_from_device_init_matches_the_reference_fixtures(golden, name):
    d = golden(name)
    e = hp.ENGINES[name](d)
    if name == "fp_small":
        e.init_state("fp")
    else:
        e.init_state("fg", rho=d["s0_rho"])
    e.step(10)
    got = e.get_state(("psi", "rho", "ux", "uy"))
    e.close()
    for k in got:
        r = d["s10_" + k] if k == "psi" else np.where(d["mask"], d["s10_" + k], 0.0)
        assert hp.rel_err(got[k], r) <= 1e-10, (name, k)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_device_init_equals_host_set_state_run(dtype):
    """the same engine started from fdlbm_init_state and from fdlbm_set_state(synthetic.fp_initial_state): same bits
    after 12 steps (fp32 too: rho0 = 1 and psi = +-1 / psi_wall are exact in fp32, and both paths round the same
    double results)"""
    from fingering_dynamics_b200 import synthetic as syn
    H, W = 256, 192
    e, c, solid, _ = _porous(H, W, dtype)
    e.set_state(**syn.fp_initial_state(solid, c))
    e.step(12)
    want = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
    e.init_state("fp", rho0=c["rho0"])
    assert e.iterations == 0
    e.step(12)
    got = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
    e.close()
    for k in got:
        assert np.array_equal(got[k], want[k]), (dtype, k)


def test_device_init_on_peer_slabs_equals_the_single_slab_run():
    from fingering_dynamics_b200 import Engine, synthetic as syn
    from fingering_dynamics_b200.slab import slab_bounds
    H, W, nslab = 256, 192, 3
    ref, c, solid, refl = _porous(H, W)
    ref.init_state("fp")
    ref.step(9)
    want = ref.get_state(("f", "g", "psi", "rho", "ux", "uy"))
    ref.close()
    kw = dict(tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"],
              zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    engs = [Engine(H, W, slab=slab_bounds(W, nslab, r), external_halo=True, **kw) for r in range(nslab)]
    for r, e in enumerate(engs):
        x0, x1 = slab_bounds(W, nslab, r)
        lo, hi = max(0, x0 - 2), min(W, x1 + 2)
        e.set_geometry(solid[:, lo:hi], refl[:, lo:hi], col0=lo)   # the minimal window: slab + ghost columns
    infos = [e.peer_export() for e in engs]
    for r, e in enumerate(engs):
        if r > 0:
            e.peer_attach(0, infos[r - 1])
        if r < nslab - 1:
            e.peer_attach(1, infos[r + 1])
    for e in engs:
        e.init_state("fp")
    for e in engs:
        e.sync()
    for n in (1, 3, 5):
        for e in engs:
            e.step(n)
    for r, e in enumerate(engs):
        x0, x1 = slab_bounds(W, nslab, r)
        got = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
        for k in got:
            assert np.array_equal(got[k], want[k][..., x0:x1]), (r, k)
        e.close()


def test_init_and_window_argument_errors():
    from fingering_dynamics_b200 import Engine, synthetic as syn
    from fingering_dynamics_b200._native import FdlbmError
    H, W = 128, 96
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    kw = dict(tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"],
              zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    e = Engine(H, W, **kw)
    with pytest.raises(FdlbmError, match="set_geometry"):
        e.init_state("fp")
    with pytest.raises(FdlbmError, match="does not cover"):      # a partial geometry window would silently drop the rest
        e.set_geometry(solid[:, :50], refl[:, :50], col0=0)
    e.set_geometry(solid, refl)
    with pytest.raises(FdlbmError, match="rho0"):
        e.init_state("fp", rho0=0.0)
    with pytest.raises(FdlbmError, match="rho window"):
        e.init_state("fg", rho=np.ones((H, 40)), col0=10)
    st = syn.fp_initial_state(solid, c)
    part = {k: np.ascontiguousarray(v[..., :50]) for k, v in st.items()}
    with pytest.raises(FdlbmError, match="does not cover the slab"):
        e.set_state(**part)
    e.init_state("fp")      # still usable after the errors
    e.step(2)
    assert e.iterations == 2 and e.count_nonfinite() == 0
    e.close()
    s = Engine(H, W, slab=(32, 64), external_halo=True, **kw)
    with pytest.raises(FdlbmError, match="ghost columns"):
        s.set_geometry(solid[:, 32:64], refl[:, 32:64], col0=32)   # owned columns only: the ghosts are missing
    s.set_geometry(solid[:, 30:66], refl[:, 30:66], col0=30)
    s.close()
