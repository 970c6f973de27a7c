"""Drop-in twin of the reference's lattice_boltzmann/bounce_back.py (class Bounce_back).

Same methods and argument meaning (bounce_back.py:5-167): the populations `f`, `g` (already streamed,
shape (9,H,W)) are updated IN PLACE from the pre-stream copies `f_behind`, `g_behind`.  The class tables
are folded into per-cell reflect bits (fingering_dynamics_b200.geometry) and applied on the GPU by
fdlbm_op_bounce_back; a whole run should use Engine.step, which does the same inside the fused kernel.
"""
try:
    from .. import geometry as geo
    from ..ops import bounce_back as _apply_bits
except ImportError:  # imported by bare name with this directory on sys.path
    from fingering_dynamics_b200 import geometry as geo
    from fingering_dynamics_b200.ops import bounce_back as _apply_bits


def _apply(reflect, f_behind, g_behind, f, g):
    _apply_bits(reflect, f_behind, g_behind, f, g)


class Bounce_back:
    def __init__(self, H, W):
        self.H = H
        self.W = W

    def left_boundary(self, f_behind, g_behind, f, g, hole):
        """bounce_back.py:13-22: column 0 outside the inlet hole reflects directions 1, 5, 8."""
        _apply(geo.reflect_bits_left_boundary(self.H, self.W, hole), f_behind, g_behind, f, g)

    def halfway_bounceback_rec(self, corner_list, f_behind, g_behind, f, g):
        """bounce_back.py:25-86: rectangles given by their corner dicts (Createblock.setblock)."""
        self.H, self.W = f[0].shape
        _apply(geo.reflect_bits_rect(corner_list, self.H, self.W), f_behind, g_behind, f, g)

    def halfway_bounceback_circle(self, side_list, concave_list, convex_list, f_behind, g_behind, f, g):
        """bounce_back.py:89-167: the 12 class masks of Createblock.setCirleblock / setEllipseblock."""
        _apply(geo.reflect_bits_circle(side_list, concave_list, convex_list), f_behind, g_behind, f, g)
