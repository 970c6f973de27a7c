"""Drop-in twin of the reference's lattice_boltzmann/create_block.py (class Createblock).

Same public surface and the same outputs, cell for cell (tests/test_create_block.py checks them
against the reference's own results in tests/golden/geometry.npz), but O(shape area) per obstacle
instead of O(H*W): the reference rasterises every obstacle on a full-grid image
(create_block.py:19-21,74), which makes ~10^4 obstacles on 8192x2048 take hours.  Here each distinct
shape (radius, or ellipse axes+angle) is rasterised and classified ONCE on a small local window with
the same cv2 / scipy calls, and the resulting cell offsets are stamped at every centre.  Shapes whose
window would leave the grid take the full-grid route, so clipping, negative-index wrap-around and
IndexError behave as in the reference.

The 12 boolean masks returned are the data the engine's reflect bits are folded from
(fingering_dynamics_b200.geometry.reflect_bits_circle).
"""
import numpy as np
import cv2
from scipy.ndimage import binary_fill_holes


def _scan_first_quadrant(outline, cx, cy, j_start, i_count):
    """Boundary-cell classes in the first quadrant of one obstacle (create_block.py:81-112).

    Rows i = 1..i_count above the centre are walked from the outside (x offset j_start) inwards until
    the outline is met; each visited fluid cell is classed by its left / lower / lower-left outline
    neighbours.  Returns offset lists (dx, dy) relative to the centre."""
    side_top, side_left, convex, concave = [], [], [], []
    jj = j_start
    for i in range(1, i_count + 1):
        for j in range(jj, 0, -1):
            if outline[cy + i, cx + j]:
                jj = j + 1
                break
            below = outline[cy + i - 1, cx + j]
            left = outline[cy + i, cx + j - 1]
            diag = outline[cy + i - 1, cx + j - 1]
            if left and below:
                concave.append((j, i))
            elif left:
                side_left.append((j, i))
            elif below:
                side_top.append((j, i))
            elif diag:
                convex.append((j, i))
    return side_top, side_left, concave, convex


def _mirror_classes(side_top, side_left, concave, convex, rx, ry):
    """Offsets of all 12 classes from the first-quadrant ones (create_block.py:114-152), keyed in the
    order the masks are returned: side[top,bottom,right,left], concave[tr,tl,br,bl], convex[tr,tl,br,bl]."""
    top = [(-dx, dy) for dx, dy in side_top] + [(0, ry + 1)] + list(side_top)
    bottom = [p for dx, dy in side_top for p in ((dx, -dy), (-dx, -dy))] + [(0, -ry - 1)]
    right = [p for dx, dy in side_left for p in ((-dx, dy), (-dx, -dy))] + [(-rx - 1, 0)]
    left = [(dx, -dy) for dx, dy in side_left] + list(side_left) + [(rx + 1, 0)]

    def four(q):
        return (list(q), [(-dx, dy) for dx, dy in q], [(dx, -dy) for dx, dy in q], [(-dx, -dy) for dx, dy in q])

    return (top, bottom, right, left) + four(concave) + four(convex)


class _Template:
    """filled-cell offsets and the 12 class offset arrays of one shape, relative to its centre"""

    def __init__(self, fill_dx, fill_dy, classes, reach):
        self.fill_dx, self.fill_dy = fill_dx, fill_dy
        self.classes = [np.array(c, dtype=np.int64).reshape(-1, 2) for c in classes]
        self.reach = reach  # half-size of the window the shape was rasterised on


class Createblock:
    def __init__(self, H, W):
        self.H = H
        self.W = W
        self._templates = {}

    # -- single-shape outline images, as the reference returns them ------------------------------
    def getRectangleblock(self, bottom_left, top_right):
        block = np.zeros((self.H, self.W), dtype=np.uint8)
        cv2.rectangle(block, bottom_left, top_right, (1, 0, 0))
        return block

    def getCicleblock(self, center, radius):
        block = np.zeros((self.H, self.W), dtype=np.uint8)
        cv2.circle(block, center, radius, (1, 0, 0))
        return block

    def getEllipseblock(self, center, axes, angle):
        block = np.zeros((self.H, self.W), dtype=np.uint8)
        cv2.ellipse(block, (center, axes, angle), (1, 0, 0))
        return block

    def getCorner(self, block):
        ys, xs = np.nonzero(block)
        x0, x1, y0, y1 = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
        return {"top_left": (x0, y1), "bottom_left": (x0, y0), "top_right": (x1, y1), "bottom_right": (x1, y0)}

    # -- templates --------------------------------------------------------------------------------
    def _template(self, key, draw, rx, ry, j_start):
        t = self._templates.get(key)
        if t is None:
            reach = max(rx, ry) + 3
            n = 2 * reach + 1
            img = np.zeros((n, n), dtype=np.uint8)
            draw(img, reach, reach)
            fill = binary_fill_holes(img)
            fy, fx = np.nonzero(fill)
            quad = _scan_first_quadrant(img, reach, reach, j_start, ry + 1)
            t = _Template(fx - reach, fy - reach, _mirror_classes(*quad, rx, ry), reach)
            self._templates[key] = t
        return t

    def _stamp_shapes(self, shapes):
        """shapes: iterable of (x, y, rx, ry, j_start, key, draw).  Returns the reference's 4-tuple."""
        H, W = self.H, self.W
        block_psi_all = np.zeros((H, W), dtype=int)
        masks = [np.zeros((H, W), dtype=bool) for _ in range(12)]
        for x, y, rx, ry, j_start, key, draw in shapes:
            reach = max(rx, ry) + 3
            if x - reach >= 0 and y - reach >= 0 and x + reach < W and y + reach < H:
                t = self._template(key, draw, rx, ry, j_start)
                block_psi_all[t.fill_dy + y, t.fill_dx + x] += 1
                for m, off in zip(masks, t.classes):
                    if len(off):
                        m[off[:, 1] + y, off[:, 0] + x] = True
            else:  # window leaves the grid: full-grid raster, numpy index semantics as in the reference
                img = np.zeros((H, W), dtype=np.uint8)
                draw(img, x, y)
                block_psi_all = block_psi_all + binary_fill_holes(img).astype(int)
                quad = _scan_first_quadrant(img, x, y, j_start, ry + 1)
                for m, off in zip(masks, _mirror_classes(*quad, rx, ry)):
                    for dx, dy in off:
                        m[y + dy, x + dx] = True
        return block_psi_all, masks[0:4], masks[4:8], masks[8:12]

    # -- the reference's entry points ---------------------------------------------------------------
    def setCirleblock(self, circle_list):
        """circle_list = [((cx, cy), r), ...] -> (block_psi_all, side_list[top,bottom,right,left],
        concave_list[tr,tl,br,bl], convex_list[tr,tl,br,bl])   (create_block.py:51-220)"""
        def shapes():
            for (cx, cy), r in circle_list:
                r = int(r)
                yield (int(cx), int(cy), r, r, r, ("c", r),
                       lambda img, x, y, r=r: cv2.circle(img, (x, y), r, (1, 0, 0)))
        return self._stamp_shapes(shapes())

    def setEllipseblock(self, ellipse_list):
        """ellipse_list = [{'c_x','c_y','r_x','r_y','angle'}, ...]; r_x, r_y are full axis lengths
        (create_block.py:223-393)."""
        def shapes():
            for el in ellipse_list:
                rx, ry = int(el["r_x"] / 2), int(el["r_y"] / 2)
                ang = el["angle"]
                yield (int(el["c_x"]), int(el["c_y"]), rx, ry, rx + 1, ("e", rx, ry, ang),
                       lambda img, x, y, rx=rx, ry=ry, ang=ang: cv2.ellipse(img, ((x, y), (int(rx * 2), int(ry * 2)), ang),
                                                                            (1, 0, 0)))
        return self._stamp_shapes(shapes())

    def setblock(self, rect_corner_list):
        """rect_corner_list = [((x0, y0), (x1, y1)), ...] -> (block_psi_all, corner_list)
        (create_block.py:395-407).  A rectangle outline is its own bounding box and its fill is the
        closed box, clipped to the grid like cv2.rectangle."""
        H, W = self.H, self.W
        block_psi_all = np.zeros((H, W), dtype=int)
        corner_list = []
        for p0, p1 in rect_corner_list:
            xa, xb = sorted((int(p0[0]), int(p1[0])))
            ya, yb = sorted((int(p0[1]), int(p1[1])))
            inside = xa >= 1 and ya >= 1 and xb < W - 1 and yb < H - 1
            if inside:
                block_psi_all[ya:yb + 1, xa:xb + 1] += 1
                corner_list.append({"top_left": (xa, yb), "bottom_left": (xa, ya), "top_right": (xb, yb),
                                    "bottom_right": (xb, ya)})
            else:
                block = self.getRectangleblock(p0, p1)
                block_psi_all = block_psi_all + binary_fill_holes(block).astype(int)
                corner_list.append(self.getCorner(block))
        return block_psi_all, corner_list
