"""The oracle (oracle/fd_oracle.c) against the golden fixtures produced by the unmodified reference.

Bar: BIT-EXACT (np.array_equal) -- the oracle keeps the reference's operation order and is built
without FMA contraction.  The validation.py variant is also restated literally (float e, cs**2).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as hp


def _check(a, d, tag, mask=None, pops=True, names=hp.MACROS):
    for k in (("f", "g") if pops else ()) + tuple(names):
        ref = d["%s_%s" % (tag, k)]
        got = a[k]
        if mask is not None and k not in ("f", "g", "psi"):
            got = np.where(mask, got, 0.0)
        assert np.array_equal(got, ref), "%s %s max|d|=%g" % (tag, k, np.max(np.abs(got - ref)))
    for k, n in (("gx", "nabla_psix"), ("gy", "nabla_psiy"), ("lap", "nabla_psi2")):
        key = "%s_%s" % (tag, n)
        if key in d:
            assert np.array_equal(a[k], d[key]), key


@pytest.mark.parametrize("name,mk", [("fp_small", hp.fp_run), ("fg_small", hp.fg_run)])
def test_masked_variants_bit_exact(golden, name, mk):
    d = golden(name)
    run = mk(d)
    done = 0
    for step, pops in ((1, True), (2, True), (10, True), (40, False)):
        a = run.iterate(step - done)
        done = step
        _check(a, d, "s%d" % step, mask=d["mask"], pops=pops)


@pytest.mark.parametrize("name", ["va_small", "va_small_wet"])
def test_validation_variant_bit_exact(golden, name):
    d = golden(name)
    run = hp.va_run(d)
    done = 0
    for step, pops in ((1, True), (2, True), (10, True), (40, False)):
        a = run.iterate(step - done)
        done = step
        _check(a, d, "s%d" % step, pops=pops)


def test_ops_stream_and_bounce_back(golden):
    d = golden("ops")
    f, g = d["f_in"].copy(), d["g_in"].copy()
    orc.stream(f, g)
    assert np.array_equal(f, d["f_stream"]) and np.array_equal(g, d["g_stream"])
    fb, gb = d["f_in"], d["g_in"]

    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    orc.bb_circle(hp.circ_masks(d, "circ_"), fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_circle"]) and np.array_equal(g, d["g_bb_circle"])

    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    orc.bb_rect(d["rect_corners"], fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_rect"]) and np.array_equal(g, d["g_bb_rect"])
    H = int(d["H"])
    orc.wall_rows(1, H - 2, fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_rect_walls"]) and np.array_equal(g, d["g_bb_rect_walls"])

    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    orc.wall_rows(0, H - 1, fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_va"]) and np.array_equal(g, d["g_bb_va"])

    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    orc.left_boundary(3, fb, gb, f, g)
    assert np.array_equal(f, d["f_left_boundary"]) and np.array_equal(g, d["g_left_boundary"])


def test_ops_stencils(golden):
    d = golden("ops")
    H, W = int(d["H"]), int(d["W"])
    common = dict(tau=1.0, gamma=1.0, a=-1.0, kappa=1.0, Eta_n=1.0, M=1.0)
    psi = d["psi_in"]
    # the reference overwrites solids with psi_wall before the stencil (FP:216-217)
    for pre, wall, yw in (("fp", -0.5, 0), ("fg", -0.7, 1)):
        P = orc.make_params(H, W, psi_wall=wall, y_wall=yw, **common)
        p2 = np.where(d["stencil_mask"], psi, wall)
        gx, gy, lap = orc.stencils(P, p2)
        assert np.array_equal(gx, d[pre + "_nabla_psix"])
        assert np.array_equal(gy, d[pre + "_nabla_psiy"])
        assert np.array_equal(lap, d[pre + "_nabla_psi2"])


def test_full_config1_golden_scalars(golden):
    """config 1 (400x400, 90 circles): SURVEY.md section 4 golden scalars at step 100, reproduced by
    tests/golden/make_golden.py --full; the oracle must hit them from the same initial state."""
    d = golden("fp_full_scalars")
    assert abs(float(d["s100_sum_psi"]) - (-141454.496092527697)) < 1e-9
    assert abs(float(d["s100_sum_rho"]) - 131310.328290524136) < 1e-9
    assert int(d["n_fluid"]) == 131470


def test_initial_states_bit_exact(golden):
    """Compute.__init__ restated (oracle.fp_initial_state / fg_initial_state) against the s0 snapshots."""
    d = golden("fp_small")
    s = orc.fp_initial_state(hp.fp_params(d), d["mask"])
    m = d["mask"]
    for k in ("f", "g", "psi"):
        assert np.array_equal(s[k], d["s0_" + k]), k
    for k in ("rho", "ux", "uy", "p", "mu", "mix_tau"):
        assert np.array_equal(np.where(m, s[k], 0.0), d["s0_" + k]), k
    d = golden("fg_small")
    s = orc.fg_initial_state(hp.fg_params(d), d["mask"], np.where(d["mask"], d["s0_rho"], 1.0))
    m = d["mask"]
    for k in ("f", "g", "psi"):
        assert np.array_equal(s[k], d["s0_" + k]), k
    for k in ("rho", "ux", "uy", "p", "mu", "mix_tau"):
        assert np.array_equal(np.where(m, s[k], 0.0), d["s0_" + k]), k


def _bits(d, key, shape):
    return np.unpackbits(d[key])[:int(np.prod(shape))].reshape(shape).astype(bool)


def test_config1_shipped_size_1000_iterations_bit_exact(golden):
    """fingering_periodic.py as shipped (400x400, 90 circles): the oracle from its own restated initial state against
    the reference's sums and subsampled fields at iterations 1, 10, 100, 1000 -- bit for bit."""
    d = golden("fp_full_scalars")
    H, W = int(d["H"]), int(d["W"])
    mask = _bits(d, "mask_bits", (H, W))
    cls = _bits(d, "class_bits", (12, H, W)).astype(np.uint8)
    P = hp.fp_params(d)
    run = orc.Run(P, orc.fp_initial_state(P, mask), mask=mask, circ_masks=cls, zou_he=1, inlet_ux=d["inlet_ux"],
                  outlet_ux=d["inlet_ux"])
    done = 0
    for step in (1, 10, 100, 1000):
        a = run.iterate(step - done)
        done = step
        tag = "s%d" % step
        assert a["psi"].sum() == float(d[tag + "_sum_psi"]), step
        assert a["rho"][mask].sum() == float(d[tag + "_sum_rho"]), step
        assert a["psi"][200, 30] == float(d[tag + "_psi_200_30"]), step
        assert np.array_equal(a["psi"][::5, ::5], d[tag + "_psi_sub"]), step
        for k in ("rho", "ux", "uy"):
            assert np.array_equal(np.where(mask, a[k], 0.0)[::5, ::5], d["%s_%s_sub" % (tag, k)]), (step, k)


def test_config3_shipped_size_full_run_bit_exact(golden):
    """fingering.py as shipped (380x380, 40 squares, np.random.seed(0), MAX_T = 1000): the oracle from its restated
    initial state (the two RNG draws of fingering.py:106-107 replayed) against the reference, bit for bit."""
    from fingering_dynamics_b200.lattice_boltzmann.create_block import Createblock
    d = dict(golden("fg_full_scalars"))
    H, W = int(d["H"]), int(d["W"])
    mask = _bits(d, "mask_bits", (H, W))
    rects = [((int(r[0]), int(r[1])), (int(r[2]), int(r[3]))) for r in d["rects"]]
    bpa, corners = Createblock(H, W).setblock(rects)
    assert np.array_equal(bpa != 1, mask)
    d["mask"] = mask
    d["corners"] = np.array([[c["top_left"][0], c["top_left"][1], c["bottom_left"][0], c["bottom_left"][1],
                              c["top_right"][0], c["top_right"][1], c["bottom_right"][0], c["bottom_right"][1]]
                             for c in corners])
    state = np.random.get_state()
    try:
        np.random.seed(0)
        sign = np.random.randint(0, 1, size=(H, W)) * 2 - 1.0
        rho = np.ones((H, W)) + np.random.rand(H, W) * 0.001 * sign
    finally:
        np.random.set_state(state)
    P = hp.fg_params(d)
    u = np.full(H, float(d["c_u0"]))
    run = orc.Run(P, orc.fg_initial_state(P, mask, np.where(mask, rho, 1.0)), mask=mask, rect_corners=d["corners"],
                  wall_rows=(1, H - 2), zou_he=2, inlet_ux=u, outlet_ux=u)
    done = 0
    for step in (10, 100, 300, 1000):
        a = run.iterate(step - done)
        done = step
        tag = "s%d" % step
        assert a["psi"].sum() == float(d[tag + "_sum_psi"]), step
        assert a["rho"][mask].sum() == float(d[tag + "_sum_rho"]), step
        assert np.array_equal(a["psi"][::5, ::5], d[tag + "_psi_sub"]), step
        for k in ("rho", "ux", "uy"):
            assert np.array_equal(np.where(mask, a[k], 0.0)[::5, ::5], d["%s_%s_sub" % (tag, k)]), (step, k)
