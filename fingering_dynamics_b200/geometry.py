"""Geometry -> the two per-cell inputs of the engine: solid flag and 8-bit "reflect direction i" mask.

The reference's half-way bounce-back is a class table, not geometric link reflection (SURVEY.md
finding 2): Createblock classifies boundary cells into 12 boolean masks (create_block.py:158-218) and
Bounce_back reflects a fixed set of directions per class (bounce_back.py:89-167; rectangles:
bounce_back.py:25-86; wall rows: fingering.py:432-451,573 and validation.py:357-376).  Every one of
those writes is `f_i(c) <- f_behind_opp(i)(c)` at the same cell, so overlapping classes simply OR
together: bit (i-1) of reflect[c] says "direction i of cell c is bounced back".
"""
import numpy as np

# class -> directions, in the order Createblock returns the lists (create_block.py:207-218) and as
# Bounce_back consumes them (bounce_back.py:90-101)
SIDE_DIRS = ((2, 5, 6), (4, 7, 8), (3, 6, 7), (1, 5, 8))      # side_list = [top, bottom, right, left]
CONCAVE_DIRS = ((1, 2, 5), (2, 3, 6), (1, 4, 8), (3, 4, 7))   # concave_list = [tr, tl, br, bl]
CONVEX_DIRS = ((5,), (6,), (8,), (7,))                        # convex_list = [tr, tl, br, bl]
# rectangles (bounce_back.py:28-86): n, s, e, w barriers and nw, ne, sw, se outer corners
RECT_DIRS = {"n": (2, 5, 6), "s": (4, 7, 8), "e": (1, 5, 8), "w": (3, 6, 7),
             "nw": (6,), "ne": (5,), "sw": (7,), "se": (8,)}


def _bits(dirs):
    b = 0
    for i in dirs:
        b |= 1 << (i - 1)
    return np.uint8(b)


def solid_from_block_psi(block_psi_all):
    """fingering_periodic.py:450: only cells covered exactly once are solid."""
    return np.ascontiguousarray((np.asarray(block_psi_all) == 1).astype(np.uint8))


def reflect_bits_circle(side_list, concave_list, convex_list):
    """Fold the 12 class masks of setCirleblock / setEllipseblock into reflect bits."""
    shape = np.asarray(side_list[0]).shape
    r = np.zeros(shape, dtype=np.uint8)
    for masks, table in ((side_list, SIDE_DIRS), (concave_list, CONCAVE_DIRS), (convex_list, CONVEX_DIRS)):
        for m, dirs in zip(masks, table):
            r[np.asarray(m, dtype=bool)] |= _bits(dirs)
    return r


def reflect_bits_rect(corner_list, H, W):
    """Fold halfway_bounceback_rec's per-call masks (bounce_back.py:28-44) into reflect bits."""
    r = np.zeros((H, W), dtype=np.uint8)
    for cor in corner_list:
        tlx, tly = cor["top_left"]
        blx, bly = cor["bottom_left"]
        trx = cor["top_right"][0]
        brx = cor["bottom_right"][0]
        r[tly + 1, tlx:trx + 1] |= _bits(RECT_DIRS["n"])
        r[bly - 1, tlx:trx + 1] |= _bits(RECT_DIRS["s"])
        r[bly:tly + 1, tlx - 1] |= _bits(RECT_DIRS["w"])
        r[bly:tly + 1, trx + 1] |= _bits(RECT_DIRS["e"])
        r[tly + 1, tlx - 1] |= _bits(RECT_DIRS["nw"])
        r[tly + 1, trx + 1] |= _bits(RECT_DIRS["ne"])
        r[bly - 1, blx - 1] |= _bits(RECT_DIRS["sw"])
        r[bly - 1, brx + 1] |= _bits(RECT_DIRS["se"])
    return r


def reflect_bits_wall_rows(H, W, row_lo, row_hi):
    """row_lo reflects {2,5,6}, row_hi reflects {4,7,8}: rows (1, H-2) for fingering.py:573,
    rows (0, H-1) for validation.py:357-376."""
    r = np.zeros((H, W), dtype=np.uint8)
    r[row_lo, :] |= _bits((2, 5, 6))
    r[row_hi, :] |= _bits((4, 7, 8))
    return r


def reflect_bits_left_boundary(H, W, hole):
    """Bounce_back.left_boundary (bounce_back.py:13-22): column 0 outside the inlet hole reflects {1,5,8}."""
    r = np.zeros((H, W), dtype=np.uint8)
    r[:int(H / 2 - hole), 0] |= _bits((1, 5, 8))
    r[int(H / 2 + hole):, 0] |= _bits((1, 5, 8))
    return r
