"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle would take minutes):
  * 8192x2048 synthetic porous medium (configs[3]): the fused kernel equals the two-pass kernel BIT FOR BIT
    (two independent schedules of the same arithmetic; the two-pass one is oracle-checked at small sizes);
  * y-translation equivariance of the y-periodic variant: rolling geometry, state and face profiles by k rows
    rolls the result by k rows, bit for bit;
  * closed box (validation.py variant, x periodic, walls): total mass and total order parameter are conserved.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(H, W, c, solid, refl, st, **kw):
    from fingering_dynamics_b200 import Engine
    e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
               psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], **kw)
    e.set_geometry(solid, refl)
    e.set_state(**st)
    return e


def test_fused_equals_twopass_at_8192x2048():
    from fingering_dynamics_b200 import synthetic as syn
    H, W = 2048, 8192
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    st = syn.fp_initial_state(solid, c)
    res = {}
    for kernel in ("fused", "twopass"):
        e = _engine(H, W, c, solid, refl, st, kernel=kernel)
        e.step(6)
        res[kernel] = e.get_state(("psi", "rho", "ux", "uy", "g"))
        e.close()
    for k in res["fused"]:
        assert np.array_equal(res["fused"][k], res["twopass"][k]), k
    assert np.isfinite(res["fused"]["psi"]).all() and res["fused"]["rho"][solid == 0].min() > 0.5


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_y_translation_equivariance(dtype):
    from fingering_dynamics_b200 import synthetic as syn
    H, W, k = 512, 1024, 77
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    st = syn.fp_initial_state(solid, c)
    e = _engine(H, W, c, solid, refl, st, dtype=dtype)
    e.step(20)
    base = e.get_state(("psi", "rho", "ux", "uy"))
    e.close()
    c2 = dict(c, inlet_ux=np.roll(c["inlet_ux"], k), outlet_ux=np.roll(c["outlet_ux"], k))
    st2 = {n: np.ascontiguousarray(np.roll(v, k, axis=-2)) for n, v in st.items()}
    e = _engine(H, W, c2, np.roll(solid, k, axis=0), np.roll(refl, k, axis=0), st2, dtype=dtype)
    e.step(20)
    moved = e.get_state(("psi", "rho", "ux", "uy"))
    e.close()
    for n in base:
        assert np.array_equal(np.roll(base[n], k, axis=0), moved[n]), n


def test_closed_box_conserves_mass_and_order_parameter():
    """validation.py variant at 2048x2048: nothing enters or leaves, so sum(rho) and sum(psi) stay put"""
    from fingering_dynamics_b200 import Engine, geometry as geo
    from fingering_dynamics_b200.lattice_boltzmann import validation as VA
    H = W = 2048
    rng = np.random.default_rng(3)
    psi = np.full((H, W), -1.0)
    yy, xx = np.ogrid[:H, :W]
    for _ in range(40):  # droplets
        cy, cx, r = rng.integers(100, H - 100), rng.integers(0, W), rng.integers(20, 60)
        psi[(yy - cy) ** 2 + (np.minimum(abs(xx - cx), W - abs(xx - cx))) ** 2 <= r * r] = 1.0
    rho = np.ones((H, W))
    z = np.zeros((H, W))
    e = Engine(H, W, tau=VA.tau, gamma=VA.gamma, a=-VA.a, kappa=VA.kappa, Eta_n=VA.Eta_n, M=VA.M, psi_wall=0.2,
               psi_y_wall=True, x_periodic=True, zou_he="none")
    e.set_geometry(np.zeros((H, W), np.uint8), geo.reflect_bits_wall_rows(H, W, 0, H - 1))
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    f = np.ascontiguousarray(w[:, None, None] * rho[None])   # rest state, u = 0, p = rho/3
    g = np.zeros((9, H, W))
    g[0] = psi
    e.set_state(f=f, g=g, psi=psi, rho=rho, ux=z, uy=z, p=rho / 3, mu=z, mix_tau=np.full((H, W), 0.8), nabla_psix=z,
                nabla_psiy=z)
    e.step(50)
    out = e.get_state(("psi", "rho"))
    e.close()
    assert abs(out["rho"].sum() - rho.sum()) <= 1e-11 * rho.sum()
    assert abs(out["psi"].sum() - psi.sum()) <= 1e-9 * abs(psi).sum()
    assert np.isfinite(out["psi"]).all()
