#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING the unmodified reference.

Run in the build container only (needs /root/reference, cv2, scipy):

    python tests/golden/make_golden.py            # all small fixtures (seconds)
    python tests/golden/make_golden.py --full     # + config-1 400x400 scalars (minutes)

Nothing here is imported by the product or by the tests; the tests only read the
.npz files this script writes.  The reference modules are imported in place
(read-only) with three shims and no edits:

  * matplotlib is not installed -> stub modules in sys.modules
    (fingering_periodic.py:2,6  fingering.py:3,7  validation.py:3,9)
  * validation.py:193 ends `class Compute` early; the orphaned functions
    validation.py:219-320 are bound back onto the class with setattr
  * module globals H, W, psi_wall ... are patched before `Compute(...)`; the
    reference reads them at call time (e.g. fingering_periodic.py:90,127,212,270)

The loop bodies below call the reference's own methods in exactly the order of
fingering_periodic.py:454-479, fingering.py:558-585 and validation.py:392-409.
"""
import argparse
import contextlib
import copy
import io
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference/lattice_boltzmann"
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.simplefilter("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        import create_block
        import bounce_back
        import fingering_periodic
        import fingering
        import validation
    for fn in ("power_law", "getMix_tau", "updatePsi", "getNabla_psix", "getNabla_psiy",
               "getNabla_psi2", "updateF", "updateG"):
        setattr(validation.Compute, fn, getattr(validation, fn))
    return create_block, bounce_back, fingering_periodic, fingering, validation


CB, BB, FP, FG, VA = _import_reference()


def full(mask, v):
    """masked 1-D reference array -> full grid (zeros on solids)."""
    out = np.zeros(mask.shape)
    out[mask] = v
    return out


def consts(mod, names):
    return {"c_" + n: np.float64(getattr(mod, n)) for n in names}


def snap_masked(cm, mask, tag, d, with_pops=True):
    if with_pops:
        d[tag + "_f"] = cm.f.copy()
        d[tag + "_g"] = cm.g.copy()
    d[tag + "_psi"] = cm.psi.copy()
    for n in ("rho", "ux", "uy", "p", "mu", "mix_tau"):
        d[tag + "_" + n] = full(mask, getattr(cm, n))
    d[tag + "_nabla_psix"] = cm.nabla_psix.copy()
    d[tag + "_nabla_psiy"] = cm.nabla_psiy.copy()
    d[tag + "_nabla_psi2"] = cm.nabla_psi2.copy()


# ----------------------------------------------------------------------------------------------
# config 1: fingering_periodic.py
# ----------------------------------------------------------------------------------------------
def fp_iteration(cm, mask, bb, side_list, concave_list, convex_list):
    """fingering_periodic.py:455-479 verbatim order."""
    for j in range(9):
        cm.F[j] = cm.getLarge_F(j)
        cm.feq[j] = cm.getfeq(j)
        cm.geq[j] = cm.getgeq(j)
        cm.f[j][mask] = cm.getF(j)
        cm.g[j][mask] = cm.getG(j)
    f_behind = copy.deepcopy(cm.f)
    g_behind = copy.deepcopy(cm.g)
    FP.stream(cm.f, cm.g)
    bb.halfway_bounceback_circle(side_list, concave_list, convex_list, f_behind, g_behind, cm.f, cm.g)
    cm.zou_he_boundary_inlet()
    cm.zou_he_boundary_outlet()
    cm.rho = cm.getRho()
    cm.udpatePsi()
    cm.nabla_psix = cm.getNabla_psix()
    cm.nabla_psiy = cm.getNabla_psiy()
    cm.nabla_psi2 = cm.getNabla_psi2()
    cm.mu = cm.getMu()
    cm.ux = cm.getUx()
    cm.uy = cm.getUy()
    cm.p = cm.getP()
    cm.mix_tau = cm.getMix_tau()


def fp_setup(H, W, circle_list, psi_wall=None):
    FP.H, FP.W = H, W
    if psi_wall is not None:
        FP.psi_wall = psi_wall
    cr = CB.Createblock(H, W)
    bb = BB.Bounce_back(H, W)
    block_psi_all, side_list, concave_list, convex_list = cr.setCirleblock(circle_list)
    mask = np.logical_not(np.where(block_psi_all == 1, True, False))
    cm = FP.Compute(mask)
    return cm, mask, bb, block_psi_all, side_list, concave_list, convex_list


def geometry_dict(block_psi_all, side_list, concave_list, convex_list):
    d = {"block_psi_all": block_psi_all.astype(np.int64)}
    for k, m in enumerate(side_list):
        d["side_%d" % k] = m
    for k, m in enumerate(concave_list):
        d["concave_%d" % k] = m
    for k, m in enumerate(convex_list):
        d["convex_%d" % k] = m
    return d


def make_fp_small():
    H, W = 32, 48
    circles = [((12, 8), 4), ((12, 24), 5), ((26, 16), 5), ((36, 6), 3), ((37, 25), 4)]
    old = (FP.H, FP.W, FP.psi_wall)
    cm, mask, bb, bpa, sl, cl, vl = fp_setup(H, W, circles, psi_wall=-0.5)
    d = {"H": H, "W": W, "circles": np.array([[c[0][0], c[0][1], c[1]] for c in circles]),
         "mask": mask}
    d.update(geometry_dict(bpa, sl, cl, vl))
    d.update(consts(FP, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "u0", "psi_wall"]))
    temp = np.array([i * 3 / (H / 2) for i in range(int(-H / 2), int(H / 2))])
    d["inlet_ux"] = FP.u0 * np.exp(-(temp ** 2) / 2)
    snap_masked(cm, mask, "s0", d)
    checkpoints = {1: True, 2: True, 10: True, 40: False}
    for it in range(1, 41):
        fp_iteration(cm, mask, bb, sl, cl, vl)
        if it in checkpoints:
            snap_masked(cm, mask, "s%d" % it, d, with_pops=checkpoints[it])
    FP.H, FP.W, FP.psi_wall = old
    np.savez_compressed(os.path.join(OUT, "fp_small.npz"), **d)
    print("fp_small", sum(v.nbytes for v in d.values() if hasattr(v, "nbytes")) // 1024, "KiB raw")


def make_fp_full():
    """config 1 as shipped (fingering_periodic.py:15-40, 406-420): scalars + subsampled fields."""
    H, W = 400, 400
    circles = []
    r, xx, count = 10, 10, 1
    z = r + xx - 25
    while True:
        if count * (xx + r) - z > 380:
            break
        for i in range(FP.block_num):
            circles.append(((count * (r + xx) - z, (2 * i + 1) * (r + xx)), r))
        count += 2
    cm, mask, bb, bpa, sl, cl, vl = fp_setup(H, W, circles)
    d = {"H": H, "W": W, "circles": np.array([[c[0][0], c[0][1], c[1]] for c in circles]),
         "n_fluid": int(mask.sum())}
    d.update(consts(FP, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "u0", "psi_wall"]))
    # geometry, bit-packed: fluid mask + the 12 class masks in create_block.py:207-218 order
    d["mask_bits"] = np.packbits(mask)
    d["class_bits"] = np.packbits(np.stack(list(sl) + list(cl) + list(vl)))
    temp = np.array([i * 3 / (H / 2) for i in range(int(-H / 2), int(H / 2))])
    d["inlet_ux"] = FP.u0 * np.exp(-(temp ** 2) / 2)
    want = (1, 10, 100, 1000)
    for it in range(1, max(want) + 1):
        fp_iteration(cm, mask, bb, sl, cl, vl)
        if it in want:
            tag = "s%d" % it
            d[tag + "_sum_psi"] = np.float64(cm.psi.sum())
            d[tag + "_sum_rho"] = np.float64(cm.rho.sum())
            d[tag + "_sum_ux"] = np.float64(cm.ux.sum())
            d[tag + "_psi_200_30"] = np.float64(cm.psi[200, 30])
            d[tag + "_psi_sub"] = cm.psi[::5, ::5].copy()
            d[tag + "_rho_sub"] = full(mask, cm.rho)[::5, ::5].copy()
            d[tag + "_ux_sub"] = full(mask, cm.ux)[::5, ::5].copy()
            d[tag + "_uy_sub"] = full(mask, cm.uy)[::5, ::5].copy()
            print("fp_full step", it, d[tag + "_sum_psi"], d[tag + "_sum_rho"], flush=True)
    np.savez_compressed(os.path.join(OUT, "fp_full_scalars.npz"), **d)


def fp_default_circles():
    circles = []
    r, xx, count = 10, 10, 1
    z = r + xx - 25
    while True:
        if count * (xx + r) - z > 380:
            break
        for i in range(FP.block_num):
            circles.append(((count * (r + xx) - z, (2 * i + 1) * (r + xx)), r))
        count += 2
    return circles


def time_fp_default(steps=12, warmup=2):
    """Wall time of the UNMODIFIED reference loop body on config 1 as shipped (400x400, 90 circles): bench.py's
    `numpy_reference_mlups`.  NumPy ufuncs run on one thread: 1 core."""
    import time
    H, W = 400, 400
    cm, mask, bb, bpa, sl, cl, vl = fp_setup(H, W, fp_default_circles())
    for _ in range(warmup):
        fp_iteration(cm, mask, bb, sl, cl, vl)
    t0 = time.perf_counter()
    for _ in range(steps):
        fp_iteration(cm, mask, bb, sl, cl, vl)
    dt = time.perf_counter() - t0
    return {"value": H * W * steps / dt / 1e6, "unit": "MLUPS", "grid": "400x400 (config 1 as shipped)", "steps": steps,
            "ms_per_step": dt / steps * 1e3, "cores": "1 of %d" % (os.cpu_count() or 1),
            "what": "fingering_periodic.py:455-479 executed from /root/reference, unmodified (print calls dropped)"}


# ----------------------------------------------------------------------------------------------
# config 3: fingering.py
# ----------------------------------------------------------------------------------------------
def fg_iteration(cm, mask, bb, corner_list):
    """fingering.py:559-585 verbatim order (without the psi frame append)."""
    for j in range(9):
        cm.F[j] = cm.getLarge_F(j)
        cm.feq[j] = cm.getfeq(j)
        cm.geq[j] = cm.getgeq(j)
        cm.f[j][mask] = cm.getF(j)
        cm.g[j][mask] = cm.getG(j)
    f_behind = copy.deepcopy(cm.f)
    g_behind = copy.deepcopy(cm.g)
    FG.stream(cm.f, cm.g)
    bb.halfway_bounceback_rec(corner_list, f_behind, g_behind, cm.f, cm.g)
    FG.bottom_top_wall(f_behind[:, 1:-1], g_behind[:, 1:-1], cm.f[:, 1:-1], cm.g[:, 1:-1])
    cm.zou_he_boundary_inlet()
    cm.zou_he_boundary_outlet()
    cm.rho = cm.getRho()
    cm.udpatePsi()
    cm.nabla_psix = cm.getNabla_psix()
    cm.nabla_psiy = cm.getNabla_psiy()
    cm.nabla_psi2 = cm.getNabla_psi2()
    cm.mu = cm.getMu()
    cm.ux = cm.getUx()
    cm.uy = cm.getUy()
    cm.p = cm.getP()
    cm.mix_tau = cm.getMix_tau()


def make_fg_small():
    H, W = 36, 44
    rects = [((8, 6), (13, 11)), ((8, 22), (14, 27)), ((22, 13), (28, 19)), ((33, 4), (37, 9)),
             ((32, 24), (38, 30))]
    old = (FG.H, FG.W)
    FG.H, FG.W = H, W
    cr = CB.Createblock(H, W)
    bb = BB.Bounce_back(H, W)
    block_psi_all, corner_list = cr.setblock(rects)
    mask = np.logical_not(np.where(block_psi_all == 1, True, False))
    np.random.seed(7)
    cm = FG.Compute(mask)
    d = {"H": H, "W": W, "rects": np.array([[r[0][0], r[0][1], r[1][0], r[1][1]] for r in rects]),
         "mask": mask, "block_psi_all": block_psi_all.astype(np.int64),
         "corners": np.array([[c["top_left"][0], c["top_left"][1], c["bottom_left"][0], c["bottom_left"][1],
                               c["top_right"][0], c["top_right"][1], c["bottom_right"][0], c["bottom_right"][1]]
                              for c in corner_list])}
    d.update(consts(FG, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "u0", "psi_wall"]))
    snap_masked(cm, mask, "s0", d)
    checkpoints = {1: True, 2: True, 10: True, 40: False}
    for it in range(1, 41):
        fg_iteration(cm, mask, bb, corner_list)
        if it in checkpoints:
            snap_masked(cm, mask, "s%d" % it, d, with_pops=checkpoints[it])
    FG.H, FG.W = old
    np.savez_compressed(os.path.join(OUT, "fg_small.npz"), **d)
    print("fg_small", sum(v.nbytes for v in d.values() if hasattr(v, "nbytes")) // 1024, "KiB raw")


# ----------------------------------------------------------------------------------------------
# config 2: validation.py
# ----------------------------------------------------------------------------------------------
def va_iteration(cm):
    """validation.py:393-409 verbatim order."""
    for j in range(9):
        cm.feq[j] = cm.getfeq(j)
        cm.geq[j] = cm.getgeq(j)
    cm.mix_tau = cm.getMix_tau()
    for j in range(9):
        cm.F[j] = cm.getLarge_F(j)
    cm.updateF()
    cm.updateG()
    f_behind = copy.deepcopy(cm.f)
    g_behind = copy.deepcopy(cm.g)
    VA.stream(cm.f, cm.g)
    VA.halfway_bounceback(f_behind, g_behind, cm.f, cm.g)
    cm.updateRho()
    cm.updatePsi()
    cm.updateMu()
    cm.updateU()
    cm.updateP()


def snap_va(cm, tag, d, with_pops=True):
    if with_pops:
        d[tag + "_f"] = cm.f.copy()
        d[tag + "_g"] = cm.g.copy()
    for n in ("psi", "rho", "ux", "uy", "p", "mu", "mix_tau"):
        d[tag + "_" + n] = np.array(getattr(cm, n), dtype=np.float64).copy()


def make_va_small(psi_wall, name):
    H, W = 48, 88
    old = (VA.H, VA.W, VA.psi_wall)
    VA.H, VA.W, VA.psi_wall = H, W, psi_wall
    with contextlib.redirect_stdout(io.StringIO()):
        cm = VA.Compute()
    d = {"H": H, "W": W}
    d.update(consts(VA, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "psi_wall", "cs", "c"]))
    d["e"] = cm.e.copy()
    d["w"] = cm.w.copy()
    snap_va(cm, "s0", d)
    d["s0_nabla_psix"] = cm.getNabla_psix()
    d["s0_nabla_psiy"] = cm.getNabla_psiy()
    d["s0_nabla_psi2"] = cm.getNabla_psi2()
    checkpoints = {1: True, 2: True, 10: True, 40: False}
    for it in range(1, 41):
        va_iteration(cm)
        if it in checkpoints:
            snap_va(cm, "s%d" % it, d, with_pops=checkpoints[it])
    VA.H, VA.W, VA.psi_wall = old
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, sum(v.nbytes for v in d.values() if hasattr(v, "nbytes")) // 1024, "KiB raw")


# ----------------------------------------------------------------------------------------------
# unit-op fixtures: stream / bounce-back / stencils on random populations
# ----------------------------------------------------------------------------------------------
def make_ops():
    rng = np.random.default_rng(20261017)
    H, W = 24, 30
    d = {"H": H, "W": W}
    f = rng.random((9, H, W))
    g = rng.random((9, H, W))
    d["f_in"], d["g_in"] = f.copy(), g.copy()
    fs, gs = f.copy(), g.copy()
    FP.stream(fs, gs)
    d["f_stream"], d["g_stream"] = fs.copy(), gs.copy()

    # class-table bounce back on circles (bounce_back.py:89-167)
    cr = CB.Createblock(H, W)
    bb = BB.Bounce_back(H, W)
    circles = [((8, 7), 4), ((20, 14), 5)]
    bpa, sl, cl, vl = cr.setCirleblock(circles)
    d["circ_circles"] = np.array([[c[0][0], c[0][1], c[1]] for c in circles])
    for k, v in geometry_dict(bpa, sl, cl, vl).items():
        d["circ_" + k] = v
    fb, gb = fs.copy(), gs.copy()
    bb.halfway_bounceback_circle(sl, cl, vl, f, g, fb, gb)
    d["f_bb_circle"], d["g_bb_circle"] = fb, gb

    # rectangle table (bounce_back.py:25-86) + wall rows (fingering.py:432-451,573)
    rects = [((5, 4), (9, 8)), ((17, 12), (23, 17))]
    bpa_r, corner_list = cr.setblock(rects)
    d["rect_rects"] = np.array([[r[0][0], r[0][1], r[1][0], r[1][1]] for r in rects])
    d["rect_block_psi_all"] = bpa_r.astype(np.int64)
    d["rect_corners"] = np.array([[c["top_left"][0], c["top_left"][1], c["bottom_left"][0], c["bottom_left"][1],
                                   c["top_right"][0], c["top_right"][1], c["bottom_right"][0],
                                   c["bottom_right"][1]] for c in corner_list])
    fb, gb = fs.copy(), gs.copy()
    bb.halfway_bounceback_rec(corner_list, f, g, fb, gb)
    d["f_bb_rect"], d["g_bb_rect"] = fb.copy(), gb.copy()
    FG.bottom_top_wall(f[:, 1:-1], g[:, 1:-1], fb[:, 1:-1], gb[:, 1:-1])
    d["f_bb_rect_walls"], d["g_bb_rect_walls"] = fb, gb
    fb, gb = fs.copy(), gs.copy()
    VA.halfway_bounceback(f, g, fb, gb)
    d["f_bb_va"], d["g_bb_va"] = fb, gb
    fb, gb = fs.copy(), gs.copy()
    bb.left_boundary(f, g, fb, gb, 3)
    d["f_left_boundary"], d["g_left_boundary"] = fb, gb

    # stencils in the three flavours on a random psi with a solid mask
    psi = rng.random((H, W)) * 2 - 1
    d["psi_in"] = psi.copy()
    old = (FP.H, FP.W, FP.psi_wall, FG.H, FG.W, FG.psi_wall, VA.H, VA.W, VA.psi_wall)
    FP.H, FP.W, FP.psi_wall = H, W, -0.5
    FG.H, FG.W, FG.psi_wall = H, W, -0.7
    VA.H, VA.W, VA.psi_wall = H, W, 0.3
    mask = np.logical_not(bpa == 1)
    d["stencil_mask"] = mask
    cm = FP.Compute(mask)
    cm.psi = psi.copy()
    d["fp_nabla_psix"], d["fp_nabla_psiy"], d["fp_nabla_psi2"] = cm.getNabla_psix(), cm.getNabla_psiy(), cm.getNabla_psi2()
    np.random.seed(1)
    cm = FG.Compute(mask)
    cm.psi = psi.copy()
    d["fg_nabla_psix"], d["fg_nabla_psiy"], d["fg_nabla_psi2"] = cm.getNabla_psix(), cm.getNabla_psiy(), cm.getNabla_psi2()
    (FP.H, FP.W, FP.psi_wall, FG.H, FG.W, FG.psi_wall) = old[:6]
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **d)
    print("ops", sum(v.nbytes for v in d.values() if hasattr(v, "nbytes")) // 1024, "KiB raw")
    (VA.H, VA.W, VA.psi_wall) = old[6:]


def make_geometry():
    """Createblock outputs (create_block.py:51-407) for a handful of shape lists."""
    rng = np.random.default_rng(99)
    d = {}
    H, W = 64, 80
    cr = CB.Createblock(H, W)
    for case in range(4):
        circles = []
        for cx in range(14, W - 14, 26):
            for cy in range(14, H - 14, 26):
                circles.append(((int(cx + rng.integers(-2, 3)), int(cy + rng.integers(-2, 3))),
                                int(rng.integers(3, 11))))
        bpa, sl, cl, vl = cr.setCirleblock(circles)
        d["circ%d_list" % case] = np.array([[c[0][0], c[0][1], c[1]] for c in circles])
        for k, v in geometry_dict(bpa, sl, cl, vl).items():
            d["circ%d_%s" % (case, k)] = v
    # overlapping circles: doubly covered cells have block_psi_all == 2 (create_block.py:75)
    circles = [((20, 20), 8), ((30, 22), 8), ((60, 40), 10)]
    bpa, sl, cl, vl = cr.setCirleblock(circles)
    d["circ_overlap_list"] = np.array([[c[0][0], c[0][1], c[1]] for c in circles])
    for k, v in geometry_dict(bpa, sl, cl, vl).items():
        d["circ_overlap_%s" % k] = v
    # ellipses (create_block.py:223-393)
    ellipses = [{"c_x": 20, "c_y": 20, "r_x": 16, "r_y": 10, "angle": 0},
                {"c_x": 55, "c_y": 40, "r_x": 12, "r_y": 20, "angle": 0}]
    bpa, sl, cl, vl = cr.setEllipseblock(ellipses)
    d["ell_list"] = np.array([[e["c_x"], e["c_y"], e["r_x"], e["r_y"], e["angle"]] for e in ellipses])
    for k, v in geometry_dict(bpa, sl, cl, vl).items():
        d["ell_%s" % k] = v
    rects = [((5, 4), (9, 8)), ((17, 12), (23, 17)), ((40, 30), (60, 50)), ((62, 5), (70, 6))]
    bpa, corner_list = cr.setblock(rects)
    d["rect_list"] = np.array([[r[0][0], r[0][1], r[1][0], r[1][1]] for r in rects])
    d["rect_block_psi_all"] = bpa.astype(np.int64)
    d["rect_corners"] = np.array([[c["top_left"][0], c["top_left"][1], c["bottom_left"][0], c["bottom_left"][1],
                                   c["top_right"][0], c["top_right"][1], c["bottom_right"][0],
                                   c["bottom_right"][1]] for c in corner_list])
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **d)
    print("geometry", sum(v.nbytes for v in d.values() if hasattr(v, "nbytes")) // 1024, "KiB raw")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run config 1 at 400x400 for 1000 steps (~3 min)")
    ap.add_argument("--only", default="")
    ap.add_argument("--full23", action="store_true", help="configs 2 and 3 at their shipped sizes (~2 min)")
    args = ap.parse_args()
    np.seterr(all="raise")  # the reference's own failure detection (fingering_periodic.py:497)
    todo = {"fp": make_fp_small, "fg": make_fg_small,
            "va": lambda: (make_va_small(0.0, "va_small.npz"), make_va_small(0.3, "va_small_wet.npz")),
            "ops": make_ops, "geometry": make_geometry}
    for k, fn in todo.items():
        if not args.only or k in args.only.split(","):
            fn()
    if args.full:
        make_fp_full()


# ----------------------------------------------------------------------------------------------
# configs 2 and 3 at their shipped sizes: scalars + subsampled fields (run with --full)
# ----------------------------------------------------------------------------------------------
def make_fg_full():
    """fingering.py as shipped (380x380, 40 squares, fingering.py:16-50, 534-550), np.random.seed(0)."""
    H, W = FG.H, FG.W
    rects, count, flag = [], 1, True
    while True:
        if (count + 1) * 20 > 380:
            break
        if flag:
            for i in range(4):
                rects.append(((count * 20, 60 * (i + 1) + i * 20), ((count + 1) * 20, 60 * (i + 1) + (i + 1) * 20)))
            flag = False
            count += 2
            continue
        else:
            for i in range(5):
                rects.append(((count * 20, 60 * i + (i + 1) * 20), ((count + 1) * 20, 60 * i + (i + 2) * 20)))
            flag = True
            count += 2
    cr = CB.Createblock(H, W)
    bb = BB.Bounce_back(H, W)
    block_psi_all, corner_list = cr.setblock(rects)
    mask = np.logical_not(np.where(block_psi_all == 1, True, False))
    np.random.seed(0)
    cm = FG.Compute(mask)
    d = {"H": H, "W": W, "n_rects": len(rects), "n_fluid": int(mask.sum()), "mask_bits": np.packbits(mask),
         "rects": np.array([[r[0][0], r[0][1], r[1][0], r[1][1]] for r in rects])}
    d.update(consts(FG, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "u0", "psi_wall"]))
    for it in range(1, FG.MAX_T + 1):   # the shipped run length (fingering.py:18)
        fg_iteration(cm, mask, bb, corner_list)
        if it in (10, 100, 300, FG.MAX_T):
            tag = "s%d" % it
            d[tag + "_sum_psi"] = np.float64(cm.psi.sum())
            d[tag + "_sum_rho"] = np.float64(cm.rho.sum())
            for k in ("rho", "ux", "uy"):
                d["%s_%s_sub" % (tag, k)] = full(mask, getattr(cm, k))[::5, ::5].copy()
            d[tag + "_psi_sub"] = cm.psi[::5, ::5].copy()
            print("fg_full step", it, d[tag + "_sum_psi"], d[tag + "_sum_rho"], flush=True)
    np.savez_compressed(os.path.join(OUT, "fg_full_scalars.npz"), **d)


def make_va_full():
    """validation.py as shipped (200x250, psi_wall = 0, validation.py:14-40)."""
    with contextlib.redirect_stdout(io.StringIO()):
        cm = VA.Compute()
    d = {"H": VA.H, "W": VA.W}
    d.update(consts(VA, ["tau", "gamma", "a", "kappa", "Eta_n", "M", "psi_wall", "cs", "c"]))
    for it in range(1, VA.MAX_T + 1):   # the shipped run length (validation.py:16)
        va_iteration(cm)
        if it in (10, 100, 500, VA.MAX_T):
            tag = "s%d" % it
            d[tag + "_sum_psi"] = np.float64(cm.psi.sum())
            d[tag + "_sum_rho"] = np.float64(cm.rho.sum())
            for k in ("psi", "rho", "ux", "uy"):
                d["%s_%s_sub" % (tag, k)] = np.asarray(getattr(cm, k))[::5, ::5].copy()
            print("va_full step", it, d[tag + "_sum_psi"], d[tag + "_sum_rho"], flush=True)
    np.savez_compressed(os.path.join(OUT, "va_full_scalars.npz"), **d)


if __name__ == "__main__" and "--full23" in sys.argv:
    np.seterr(all="raise")
    make_fg_full()
    make_va_full()
