"""Build-container only: the oracle against the UNMODIFIED reference executed live on random cases.

tests/test_oracle.py pins the oracle to the committed fixtures; here the same harness that wrote those fixtures
(tests/golden/make_golden.py, which imports /root/reference/lattice_boltzmann in place) runs the reference on
random grid sizes, obstacle lists, wettabilities and step counts, and the oracle must reproduce every array
BIT FOR BIT.  Skipped wherever /root/reference does not exist (the GPU box).
"""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as hp

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/lattice_boltzmann"),
                                reason="the reference tree only exists in the build container")


@pytest.fixture(scope="module")
def mg():
    with contextlib.redirect_stdout(io.StringIO()):
        from tests.golden import make_golden
    return make_golden


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _compare(a, want, mask, tag):
    for k in ("f", "g") + hp.MACROS:
        got = a[k]
        if mask is not None and k not in ("f", "g", "psi"):
            got = np.where(mask, got, 0.0)
        assert np.array_equal(got, want["s_" + k]), (tag, k, float(np.max(np.abs(got - want["s_" + k]))))


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_fingering_periodic_random_cases(mg, seed):
    """fingering_periodic.py:47-343,454-479 + bounce_back.py:89-167 + create_block.py:51-220 (even H: the
    reference's inlet profile has H-1 entries for odd H, fingering_periodic.py:270, and raises)."""
    FP = mg.FP
    rng = np.random.default_rng(100 + seed)
    H, W = 2 * int(rng.integers(12, 30)), int(rng.integers(30, 70))
    circles = []
    for _ in range(int(rng.integers(0, 5))):
        r = int(rng.integers(2, 7))
        circles.append(((int(rng.integers(8 + r, W - r - 5)), int(rng.integers(r + 3, H - r - 3))), r))
    pw = float(rng.choice([-1.0, -0.5, 0.0, 0.3]))
    steps = int(rng.integers(1, 9))
    old = (FP.H, FP.W, FP.psi_wall)
    try:
        with _quiet():
            cm, mask, bb, bpa, sl, cl, vl = mg.fp_setup(H, W, circles, psi_wall=pw)
        d = {}
        mg.snap_masked(cm, mask, "s0", d)
        t = np.array([i * 3 / (H / 2) for i in range(int(-H / 2), int(H / 2))])
        prof = FP.u0 * np.exp(-(t ** 2) / 2)
        P = orc.make_params(H, W, tau=FP.tau, gamma=FP.gamma, a=FP.a, kappa=FP.kappa, Eta_n=FP.Eta_n, M=FP.M,
                            psi_wall=pw, y_wall=0, outlet_f3_coef=2 / 3)
        run = orc.Run(P, hp.fixture_state(d, "s0"), mask=mask,
                      circ_masks=np.stack(list(sl) + list(cl) + list(vl)).astype(np.uint8), zou_he=1, inlet_ux=prof,
                      outlet_ux=prof)
        with _quiet():
            for _ in range(steps):
                mg.fp_iteration(cm, mask, bb, sl, cl, vl)
        want = {}
        mg.snap_masked(cm, mask, "s", want)
        _compare(run.iterate(steps), want, mask, (H, W, circles, pw, steps))
    finally:
        FP.H, FP.W, FP.psi_wall = old


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_fingering_random_cases(mg, seed):
    """fingering.py:53-451,558-585 + bounce_back.py:25-86 + create_block.py:395-407, seeded global RNG
    (fingering.py:106-107)."""
    FG, CB, BB = mg.FG, mg.CB, mg.BB
    rng = np.random.default_rng(200 + seed)
    H, W = int(rng.integers(28, 56)), int(rng.integers(36, 64))
    rects = []
    for _ in range(int(rng.integers(0, 4))):
        x0, y0 = int(rng.integers(6, W - 14)), int(rng.integers(4, H - 12))
        rects.append(((x0, y0), (x0 + int(rng.integers(2, 8)), y0 + int(rng.integers(2, 8)))))
    steps = int(rng.integers(1, 9))
    old = (FG.H, FG.W)
    try:
        FG.H, FG.W = H, W
        with _quiet():
            block_psi_all, corner_list = CB.Createblock(H, W).setblock(rects)
            bb = BB.Bounce_back(H, W)
            mask = np.logical_not(np.where(block_psi_all == 1, True, False))
            np.random.seed(seed)
            cm = FG.Compute(mask)
        d = {"H": H, "W": W, "mask": mask,
             "corners": np.array([[c["top_left"][0], c["top_left"][1], c["bottom_left"][0], c["bottom_left"][1],
                                   c["top_right"][0], c["top_right"][1], c["bottom_right"][0], c["bottom_right"][1]]
                                  for c in corner_list]).reshape(-1, 8)}
        d.update(mg.consts(FG, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "u0", "psi_wall"]))
        mg.snap_masked(cm, mask, "s0", d)
        run = hp.fg_run(d)
        with _quiet():
            for _ in range(steps):
                mg.fg_iteration(cm, mask, bb, corner_list)
        want = {}
        mg.snap_masked(cm, mask, "s", want)
        _compare(run.iterate(steps), want, mask, (H, W, rects, steps))
    finally:
        FG.H, FG.W = old


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_validation_random_cases(mg, seed):
    """validation.py:43-376,392-409 (literal float e-vectors and cs**2), random size and wettability."""
    VA = mg.VA
    rng = np.random.default_rng(300 + seed)
    H, W = int(rng.integers(80, 110)), int(rng.integers(90, 130))   # the droplet of validation.py:80-92 must fit
    pw = float(rng.choice([-0.6, -0.3, 0.0, 0.3]))
    steps = int(rng.integers(1, 7))
    old = (VA.H, VA.W, VA.psi_wall)
    try:
        VA.H, VA.W, VA.psi_wall = H, W, pw
        with _quiet():
            cm = VA.Compute()
        d = {"H": H, "W": W}
        d.update(mg.consts(VA, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "psi_wall", "cs", "c"]))
        d["e"], d["w"] = cm.e.copy(), cm.w.copy()
        mg.snap_va(cm, "s0", d)
        run = hp.va_run(d)
        with _quiet():
            for _ in range(steps):
                mg.va_iteration(cm)
        want = {}
        mg.snap_va(cm, "s", want)
        _compare(run.iterate(steps), want, None, (H, W, pw, steps))
    finally:
        VA.H, VA.W, VA.psi_wall = old


def test_validation_shipped_size_full_run_bit_exact(mg, golden):
    """config 2 as shipped (validation.py: 200x250, psi_wall = 0, MAX_T = 1000): the oracle, started from the reference's
    own Compute() state, against the reference's sums and subsampled fields at iterations 10, 100, 500, 1000."""
    VA = mg.VA
    d = dict(golden("va_full_scalars"))
    assert (VA.H, VA.W, VA.psi_wall) == (int(d["H"]), int(d["W"]), float(d["c_psi_wall"]))
    with _quiet():
        cm = VA.Compute()
    d["e"], d["w"] = cm.e.copy(), cm.w.copy()
    mg.snap_va(cm, "s0", d)
    run = hp.va_run(d)
    done = 0
    for step in (10, 100, 500, 1000):
        a = run.iterate(step - done)
        done = step
        tag = "s%d" % step
        assert a["psi"].sum() == float(d[tag + "_sum_psi"]), step
        assert a["rho"].sum() == float(d[tag + "_sum_rho"]), step
        for k in ("psi", "rho", "ux", "uy"):
            assert np.array_equal(a[k][::5, ::5], d["%s_%s_sub" % (tag, k)]), (step, k)
