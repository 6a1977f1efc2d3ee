"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Headless restatement of the reference's overlay drawing:

* vector layer   pathfinder_viewer.py:179-191 -- ``cv2.polylines`` of the kept vectors in (0, 0, 255), ``cv2.circle``
                 (radius 1, thickness 1) in (255, 0, 255) at their starts, then the rejected ones in (255, 255, 0)
* lamp layer     pathfinder_viewer.py:210-222 -- HSV (0, 255, V) at every kept point, ``cv2.cvtColor(HSV2BGR)``, then a
                 filled ``cv2.circle`` of radius 6 in the point's own colour

The rasterisation itself is third-party (opencv imgproc/src/drawing.cpp: ``clipLine``, ``LineIterator`` with
connectivity 8 and left-to-right ordering, the midpoint circle); ``line_pixels`` was checked against ``cv2.line`` on
20 000 random segments with end points up to 40 px outside the frame, the two circle masks against ``cv2.circle``.
"""
import numpy as np

from .pathfinder import hsv2bgr_u8

CIRCLE1 = [(0, -1), (-1, 0), (1, 0), (0, 1)]                                   # radius 1, thickness 1: (dx, dy)
DISC6_HALF_WIDTH = {0: 6, 1: 5, 2: 5, 3: 5, 4: 4, 5: 3, 6: 0}                   # radius 6, filled: |dy| -> max |dx|


def clip_line(w, h, x1, y1, x2, y2):
    """cv::clipLine (64-bit integers, double quotient truncated toward zero)."""
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def line_pixels(w, h, p1, p2):
    """Pixels cv2.line(img, p1, p2, colour, 1) sets (LINE_8)."""
    ok, x1, y1, x2, y2 = clip_line(w, h, int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]))
    if not ok:
        return []
    dx, dy = x2 - x1, y2 - y1
    sy = 1
    if dx < 0:                       # left to right
        dx, dy = -dx, -dy
        x1, y1 = x2, y2
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
    err, plus, minus = dx - 2 * dy, 2 * dx, -2 * dy
    x, y, out = x1, y1, []
    for _ in range(dx + 1):
        out.append((x, y))
        m = err < 0
        err += minus + (plus if m else 0)
        if vert:
            y += sy
            x += 1 if m else 0
        else:
            x += 1
            y += sy if m else 0
    return out


def vector_layer(all_pts, all_next, mask, width, height, draw_bad=True):
    """pathfinder_viewer.py:179-191.  all_pts / all_next int32 (N,2) (every grid point, rounded as :169-170), mask
    bool (N,) = kept.  Returns uint8 (H,W,3) BGR."""
    layer = np.zeros((height, width, 3), np.uint8)

    def draw(sel, line_col, dot_col):
        for p, q in zip(all_pts[sel], all_next[sel]):
            for x, y in line_pixels(width, height, p, q):
                layer[y, x] = line_col
        for (x1, y1) in all_pts[sel]:
            for dx, dy in CIRCLE1:
                x, y = int(x1) + dx, int(y1) + dy
                if 0 <= x < width and 0 <= y < height:
                    layer[y, x] = dot_col
    draw(mask, (0, 0, 255), (255, 0, 255))
    if draw_bad:
        draw(~mask, (255, 255, 0), (255, 255, 0))
    return layer


def lamp_layer(kept_flow, kept_pts, width, height):
    """pathfinder_viewer.py:196-223 (draw_sparse_lamps).  Returns uint8 (H,W,3) BGR."""
    fx, fy = kept_flow[:, 0], kept_flow[:, 1]
    modulus = np.sqrt(fx * fx + fy * fy)
    hsv = np.zeros((height, width, 3), np.uint8)
    for (x, y), m in zip(kept_pts, modulus):
        hsv[y, x] = (0, 255, np.minimum(50 + m * 2, 255))
    bgr = hsv2bgr_u8(hsv)
    for x, y in kept_pts:
        col = bgr[y, x].copy()
        for dy, hw in DISC6_HALF_WIDTH.items():
            for yy in ({y - dy, y + dy}):
                if 0 <= yy < height:
                    x0, x1 = max(x - hw, 0), min(x + hw, width - 1)
                    if x0 <= x1:
                        bgr[yy, x0:x1 + 1] = col
    return bgr
