"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle would take minutes):
  * 8192x2048 synthetic porous medium (configs[3]) and 32768x8192 (configs[4], one GPU): the fused kernel equals the
    two-pass kernel BIT FOR BIT (two independent schedules of the same arithmetic; the two-pass one is oracle-checked
    at small sizes);
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


def test_fused_equals_twopass_at_32768x8192():
    """BASELINE configs[4], the north-star grid, on ONE GPU (77 GB of lattices per engine, one engine at a time): the
    fused step (row pitch 8192 compiled in, 64 strips, several waves of column chunks) equals the two-pass kernel bit
    for bit after 2 steps from the device-side initial state.  The geometry is the generator's 8192 x 4096 window
    tiled 8 times along x (generating 268 M cells on the host takes half a minute; the seams are just more obstacles)."""
    from fingering_dynamics_b200 import Engine, synthetic as syn
    H, W, Wt = 8192, 32768, 4096
    c = syn.fp_constants(H)
    s1, r1 = syn.porous_geometry(H, Wt)
    solid, refl = np.ascontiguousarray(np.tile(s1, (1, W // Wt))), np.ascontiguousarray(np.tile(r1, (1, W // Wt)))
    res = {}
    for kernel in ("fused", "twopass"):
        e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                   psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], kernel=kernel)
        e.set_geometry(solid, refl)
        e.init_state(variant="fp", rho0=c["rho0"])
        e.step(2)
        res[kernel] = e.get_state(("psi", "rho"))
        e.close()
    for k in ("psi", "rho"):
        assert np.array_equal(res["fused"][k], res["twopass"][k]), k
    fluid = solid == 0
    assert np.isfinite(res["fused"]["psi"]).all() and res["fused"]["rho"][fluid].min() > 0.5
    # the run has started: the injected phase has moved into the medium and rho is no longer flat
    assert res["fused"]["rho"][fluid].std() > 0


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


@pytest.mark.parametrize("variant", ["fp", "fg", "va"])
@pytest.mark.parametrize("H,W", [(75, 131), (129, 64), (33, 40), (257, 96)])
def test_fused_equals_twopass_on_random_geometry_and_odd_sizes(variant, H, W):
    """odd H / W (strip tails, partial warps, one-row threads of the two-rows-per-thread fp32 kernel), random
    solid flags and random reflect bits: the fused kernels must reproduce the two-pass kernel --
    bit for bit in fp64, to fp32 rounding in fp32."""
    _fused_vs_twopass(variant, H, W)


@pytest.mark.parametrize("variant", ["fp", "fg", "va"])
@pytest.mark.parametrize("H,W", [(4, 33), (64, 40), (66, 40), (128, 33), (130, 96), (194, 20), (256, 64), (258, 50), (320, 48),
                                 (512, 37), (2048, 24), (4096, 20), (8192, 20)])
def test_fused_equals_twopass_on_random_geometry_and_even_sizes(variant, H, W):
    """even H: the fp32 engine runs the packed two-row kernel (lbm_fused_f32.cuh) -- one-lane warps at the strip tail
    (H = 66, 130, 194, 258), idle warps (64, 128, 320), strips with a wrapping apron, a second strip (258, 320, 512),
    the compile-time row pitches (2048, 4096, 8192)"""
    _fused_vs_twopass(variant, H, W)


def _fused_vs_twopass(variant, H, W):
    from fingering_dynamics_b200 import Engine
    rng = np.random.default_rng(H * 1000 + W)
    solid = (rng.random((H, W)) < 0.12).astype(np.uint8)
    refl = np.where(rng.random((H, W)) < 0.15, rng.integers(0, 256, size=(H, W)), 0).astype(np.uint8)
    kw = dict(tau=0.79, gamma=1.2, a=-0.04, kappa=0.09, Eta_n=0.1, M=20.0, psi_wall=-0.3)
    prof = 0.01 * (1.0 + 0.5 * np.sin(np.arange(H) / 7.0))
    if variant == "fp":
        kw.update(zou_he="fp", inlet_ux=prof, outlet_ux=prof * 0.9)
    elif variant == "fg":
        kw.update(zou_he="fg", inlet_ux=prof, outlet_ux=prof, psi_y_wall=True, outlet_f3_coef=1.5)
    else:
        kw.update(zou_he="none", x_periodic=True, psi_y_wall=True)
    fluid = solid == 0
    # a smooth order parameter plus a little noise: white noise in psi means O(1) chemical-potential forces, single
    # cells then blow up within 3 steps and an fp32 comparison would only measure those cells
    yy, xx = np.mgrid[:H, :W]
    smooth = np.sin(2 * np.pi * xx / 23.0 + 0.3) * np.cos(2 * np.pi * yy / 17.0) + 0.3 * np.sin(2 * np.pi * (xx + yy) / 9.0)
    psi = np.where(fluid, 0.7 * smooth + 0.02 * np.tanh(rng.standard_normal((H, W))), kw["psi_wall"])
    rho = 1.0 + 0.01 * rng.standard_normal((H, W))
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)[:, None, None]
    f = np.ascontiguousarray(w * rho * (1 + 0.01 * rng.standard_normal((9, H, W))) * fluid)
    g = np.ascontiguousarray(w * psi * (1 + 0.01 * rng.standard_normal((9, H, W))) * fluid)
    z = np.zeros((H, W))
    st = dict(f=f, g=g, psi=psi, rho=rho, ux=z, uy=z, p=rho / 3, mu=0.001 * rng.standard_normal((H, W)),
              mix_tau=np.full((H, W), 0.8), nabla_psix=0.01 * rng.standard_normal((H, W)),
              nabla_psiy=0.01 * rng.standard_normal((H, W)))
    for dtype, exact in (("f64", True), ("f32", False)):
        res = {}
        for kernel in ("twopass", "fused"):
            e = Engine(H, W, dtype=dtype, kernel=kernel, **kw)
            e.set_geometry(solid, refl)
            e.set_state(**st)
            e.step(3)  # (random reflect bits are unphysical: the run blows up after ~5 steps)
            res[kernel] = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
            e.close()
        for k in res["fused"]:
            a, b = res["fused"][k], res["twopass"][k]
            if exact:
                assert np.array_equal(a, b, equal_nan=True) and np.isfinite(a).all(), (dtype, k)
            else:
                # (the packed fp32 kernel evaluates the collision in its even / odd form: same algebra, other rounding)
                assert np.max(np.abs(a - b)) <= 5e-5 * max(1.0, np.max(np.abs(b))), (dtype, k, np.max(np.abs(a - b)), np.max(np.abs(b)))


def test_grid_wider_than_65535_columns_matches_the_oracle():
    """W = 70000 columns (> the 65535 limit of gridDim.y): engine (fused and two-pass) vs the oracle, 3 iterations"""
    from oracle import oracle as orc
    from fingering_dynamics_b200 import Engine, synthetic as syn
    H, W = 40, 70000
    c = syn.fp_constants(H)
    solid = np.zeros((H, W), np.uint8)
    solid[10:14, 500:520] = 1
    refl = np.zeros((H, W), np.uint8)
    P = orc.make_params(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                        psi_wall=c["psi_wall"])
    s0 = orc.fp_initial_state(P, solid == 0)
    run = orc.Run(P, s0, mask=solid == 0, circ_masks=np.zeros((12, H, W), np.uint8), zou_he=1, inlet_ux=c["inlet_ux"],
                  outlet_ux=c["outlet_ux"])
    ref = run.iterate(3)
    for kernel in ("fused", "twopass"):
        e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                   psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], kernel=kernel)
        e.set_geometry(solid, refl)
        e.set_state(f=s0["f"], g=s0["g"], psi=s0["psi"], rho=s0["rho"], ux=s0["ux"], uy=s0["uy"], p=s0["p"], mu=s0["mu"],
                    mix_tau=s0["mix_tau"], nabla_psix=s0["gx"], nabla_psiy=s0["gy"])
        e.step(3)
        got = e.get_state(("psi", "rho", "ux", "uy"))
        e.close()
        fluid = solid == 0
        for k in got:
            r = ref[k] if k == "psi" else np.where(fluid, ref[k], 0.0)
            err = np.max(np.abs(got[k] - r)) / max(np.max(np.abs(r)), 1e-300)
            assert err <= 1e-10, (kernel, k, err)


def test_benchmark_generator_matches_the_oracle_at_1024x512():
    """the benchmark workload's own generator (radii 8-12 with centre jitter: class masks the
    r = 10 circles of config 1 never produce) at 1024 x 512, fused fp64 engine against the ORACLE after 1, 10 and 100
    iterations; started once from host arrays (fdlbm_set_state) and once from the device-side Compute.__init__."""
    from oracle import oracle as orc
    from fingering_dynamics_b200 import Engine, synthetic as syn, geometry as geo
    from fingering_dynamics_b200.lattice_boltzmann.create_block import Createblock
    H, W = 512, 1024
    c = syn.fp_constants(H)
    bpa, side, cave, vex = Createblock(H, W).setCirleblock(syn.porous_circles(H, W))
    solid, refl = syn.porous_geometry(H, W)
    assert np.array_equal(solid, geo.solid_from_block_psi(bpa)) and np.array_equal(refl, geo.reflect_bits_circle(side, cave, vex))
    mask = bpa != 1
    P = orc.make_params(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                        psi_wall=c["psi_wall"])
    s0 = orc.fp_initial_state(P, mask)
    run = orc.Run(P, s0, mask=mask, circ_masks=np.stack(list(side) + list(cave) + list(vex)).astype(np.uint8), zou_he=1,
                  inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    engs = {}
    for how in ("host", "device"):
        e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                   psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], kernel="fused")
        e.set_geometry(solid, refl)
        if how == "host":
            e.set_state(**syn.fp_initial_state(solid, c))
        else:
            e.init_state("fp", rho0=c["rho0"])
        engs[how] = e
    done = 0
    for n in (1, 10, 100):
        ref = run.iterate(n - done)
        for how, e in engs.items():
            e.step(n - done)
            got = e.get_state(("psi", "rho", "ux", "uy"))
            for k in got:
                r = ref[k] if k == "psi" else np.where(mask, ref[k], 0.0)
                err = np.max(np.abs(got[k] - r)) / np.max(np.abs(r))
                assert err <= 1e-10, (how, n, k, err)
        done = n
    for e in engs.values():
        e.close()


def test_fp32_packed_kernel_at_8192x2048():
    """configs[3] in fp32: the packed two-row kernel against the fp32 two-pass kernel (same arithmetic, other rounding
    order: 5e-5 of the field maximum) and against the fp64 engine within the stated fp32 tolerance (psi, rho 2e-5,
    u 2e-3 of max |u|), 6 steps from the device-side initial state; sign(psi) agrees wherever |psi| > 1e-3."""
    from fingering_dynamics_b200 import Engine, synthetic as syn
    H, W = 2048, 8192
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    res = {}
    for tag, dtype, kernel in (("f32", "f32", "fused"), ("f32_2p", "f32", "twopass"), ("f64", "f64", "fused")):
        e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                   psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype=dtype,
                   kernel=kernel)
        e.set_geometry(solid, refl)
        e.init_state("fp", rho0=c["rho0"])
        e.step(6)
        res[tag] = e.get_state(("psi", "rho", "ux", "uy"))
        e.close()
    for k in ("psi", "rho", "ux", "uy"):
        a, b, d = res["f32"][k], res["f32_2p"][k], res["f64"][k]
        scale = np.max(np.abs(d)) if k in ("psi", "rho") else max(np.max(np.abs(res["f64"]["ux"])), np.max(np.abs(res["f64"]["uy"])))
        assert np.max(np.abs(a - b)) <= 5e-5 * max(1.0, np.max(np.abs(b))), ("packed vs two-pass", k, np.max(np.abs(a - b)))
        tol = 2e-5 if k in ("psi", "rho") else 2e-3
        assert np.max(np.abs(a - d)) <= tol * scale, ("fp32 vs fp64", k, np.max(np.abs(a - d)), scale)
    sure = np.abs(res["f64"]["psi"]) > 1e-3
    assert np.array_equal(np.sign(res["f32"]["psi"][sure]), np.sign(res["f64"]["psi"][sure]))
