"""The drop-in Createblock (local-window templates) against the reference's own outputs
(tests/golden/geometry.npz, produced by create_block.py:51-407 in the build container)."""
import os

import numpy as np
import pytest

from fingering_dynamics_b200.lattice_boltzmann.create_block import Createblock

H, W = 64, 80


def _check(d, pre, res):
    bpa, side, cave, vex = res
    assert np.array_equal(bpa, d[pre + "block_psi_all"])
    for k in range(4):
        assert np.array_equal(side[k], d["%sside_%d" % (pre, k)]), ("side", k)
        assert np.array_equal(cave[k], d["%sconcave_%d" % (pre, k)]), ("concave", k)
        assert np.array_equal(vex[k], d["%sconvex_%d" % (pre, k)]), ("convex", k)
        assert side[k].dtype == bool


@pytest.mark.parametrize("case", ["circ0_", "circ1_", "circ2_", "circ3_", "circ_overlap_"])
def test_circles(golden, case):
    d = golden("geometry")
    lst = [((int(c[0]), int(c[1])), int(c[2])) for c in d[case + "list"]]
    _check(d, case, Createblock(H, W).setCirleblock(lst))


def test_ellipses(golden):
    d = golden("geometry")
    lst = [{"c_x": int(e[0]), "c_y": int(e[1]), "r_x": int(e[2]), "r_y": int(e[3]), "angle": int(e[4])}
           for e in d["ell_list"]]
    _check(d, "ell_", Createblock(H, W).setEllipseblock(lst))


def test_rectangles(golden):
    d = golden("geometry")
    rects = [((int(r[0]), int(r[1])), (int(r[2]), int(r[3]))) for r in d["rect_list"]]
    bpa, corners = Createblock(H, W).setblock(rects)
    assert np.array_equal(bpa, d["rect_block_psi_all"])
    got = np.array([[c["top_left"][0], c["top_left"][1], c["bottom_left"][0], c["bottom_left"][1],
                     c["top_right"][0], c["top_right"][1], c["bottom_right"][0], c["bottom_right"][1]]
                    for c in corners])
    assert np.array_equal(got, d["rect_corners"])


def test_border_shapes_take_the_full_grid_route():
    """a circle whose window leaves the grid: same code path as the reference (full-grid raster);
    out-of-grid class cells raise IndexError like the reference's mask[ori[1], ori[0]] = True."""
    cb = Createblock(40, 40)
    bpa, side, cave, vex = cb.setCirleblock([((8, 8), 6)])
    assert bpa.sum() > 0 and side[0].sum() > 0
    with pytest.raises(IndexError):
        cb.setCirleblock([((36, 20), 6)])


def test_default_config1_geometry_counts(golden):
    """fingering_periodic.py:406-420: 90 circles of r=10 on 400x400 -> 131470 fluid cells."""
    d = golden("fp_full_scalars")
    lst = [((int(c[0]), int(c[1])), int(c[2])) for c in d["circles"]]
    bpa, side, cave, vex = Createblock(400, 400).setCirleblock(lst)
    mask = np.unpackbits(d["mask_bits"])[:160000].reshape(400, 400).astype(bool)
    assert int((bpa != 1).sum()) == 131470
    assert np.array_equal(bpa != 1, mask)
    cls = np.unpackbits(d["class_bits"])[:12 * 160000].reshape(12, 400, 400).astype(bool)
    for k, m in enumerate(list(side) + list(cave) + list(vex)):
        assert np.array_equal(m, cls[k]), k


REF_DIR = "/root/reference/lattice_boltzmann"


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("seed", [0, 1])
def test_random_shape_lists_against_the_reference_itself(seed):
    """Build-container only: random circle / ellipse / rectangle lists (overlaps, shapes on the border, lists that
    make the reference raise) through the reference's Createblock (create_block.py:51-407) and through the twin:
    identical outputs, or the same exception type."""
    import importlib.util
    import warnings
    spec = importlib.util.spec_from_file_location("_ref_create_block", os.path.join(REF_DIR, "create_block.py"))
    ref = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(ref)
    rng = np.random.default_rng(seed)

    def run(cls, Hh, Ww, kind, lst):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                cb = cls(Hh, Ww)
                return getattr(cb, ("setCirleblock", "setEllipseblock", "setblock")[kind])(lst), None
        except Exception as ex:  # noqa: BLE001
            return None, type(ex).__name__

    for it in range(45):
        Hh, Ww = int(rng.integers(40, 90)), int(rng.integers(40, 90))
        kind = it % 3
        n = int(rng.integers(1, 5))
        if kind == 0:
            lst = [((int(rng.integers(0, Ww)), int(rng.integers(0, Hh))), int(rng.integers(1, 13))) for _ in range(n)]
        elif kind == 1:
            lst = [{"c_x": int(rng.integers(5, Ww - 5)), "c_y": int(rng.integers(5, Hh - 5)), "r_x": int(rng.integers(4, 24)),
                    "r_y": int(rng.integers(4, 24)), "angle": int(rng.choice([0, 90, 30, 45, 120, 180]))} for _ in range(n)]
        else:
            lst = []
            for _ in range(n):
                x0, y0 = int(rng.integers(0, Ww - 3)), int(rng.integers(0, Hh - 3))
                lst.append(((x0, y0), (int(rng.integers(x0, min(Ww, x0 + 15))), int(rng.integers(y0, min(Hh, y0 + 15))))))
        want, wex = run(ref.Createblock, Hh, Ww, kind, lst)
        got, gex = run(Createblock, Hh, Ww, kind, lst)
        assert wex == gex, (kind, Hh, Ww, lst, wex, gex)
        if wex:
            continue
        assert np.array_equal(want[0], got[0]), (kind, Hh, Ww, lst)
        if kind == 2:
            assert want[1] == got[1], (Hh, Ww, lst)
        else:
            for li in (1, 2, 3):
                for k in range(4):
                    assert np.array_equal(np.asarray(want[li][k]), np.asarray(got[li][k])), (kind, li, k, Hh, Ww, lst)
