"""The drop-in Createblock (local-window templates) against the reference's own outputs
(tests/golden/geometry.npz, produced by create_block.py:51-407 in the build container)."""
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
