"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Headless numpy restatement of the reference's own downstream logic -- this part IS
reference code (not cv2), restated without the drawing calls:

* grid generator          pathfinder_viewer.py:255-267 (DenseOF.py:166-180)
* vector filter           pathfinder_viewer.py:159-178 (get_flow_lk, after the LK call)
* danger-point intensity  pathfinder_viewer.py:210-217 (draw_sparse_lamps, before drawing)
* dense-flow HSV picture   pathfinder_viewer.py:124-141 (draw_hsv; its cv2.cvtColor(HSV2BGR) call is third-party:
                          opencv color_hsv.simd.hpp, restated in hsv2bgr_u8 and pinned exhaustively against cv2)
"""
import numpy as np


def grid_points(width, height, step=30):
    """float32 (N,2), x-major flatten -- pathfinder_viewer.py:255-267."""
    if width // step % 2 == 1:
        indent_w = width % step / 2
    else:
        indent_w = (width % step + step) / 2
    if height // step % 2 == 1:
        indent_h = height % step / 2
    else:
        indent_h = (height % step + step) / 2
    g = np.mgrid[indent_w:width:step, indent_h:height:step].astype(int)
    pts = [[x, y] for x, y in zip(g[0].flatten(), g[1].flatten())]
    return np.array(pts).astype(np.float32).reshape(-1, 2)


def vector_filter(next_pts, points_, width, height, rule="viewer"):
    """pathfinder_viewer.py:159-178 (rule="viewer") or the development script's variant DenseOF.py:194-242
    (rule="denseof": the only difference is the mask, :228).
    Returns (flow int32 (M,2), points int32 (M,2), mask bool (N,), modulus f32 (N,))."""
    half_width = int(width / 2)
    half_height = int(height / 2)
    flow_ = next_pts - points_
    fx, fy = flow_[:, 0], flow_[:, 1]
    x, y = points_[:, 0], points_[:, 1]
    ang = np.arctan2(fy, fx)
    modulus = np.sqrt(fx * fx + fy * fy)
    modulus_middle = np.sqrt((half_width - x) ** 2 + (half_height - y) ** 2)
    modulus = modulus / (5 + np.sqrt(modulus_middle)) * 30
    fx = modulus * np.cos(ang)
    fy = modulus * np.sin(ang)
    nxt = np.vstack([x + fx, y + fy]).T
    nxt = np.int32(nxt + 0.5)
    pts = np.int32(points_ + 0.5)
    if rule == "denseof":
        mask = np.greater(modulus, np.median(modulus) * 1.2)                      # DenseOF.py:228
    else:
        mask = (np.median(modulus) * 1.0 < modulus) & (modulus < np.percentile(modulus, 99))
    pts_k, nxt_k = pts[mask], nxt[mask]
    return nxt_k - pts_k, pts_k, mask, modulus


def danger_intensity(flow_, points_):
    """pathfinder_viewer.py:210-217: V channel written at each kept point, uint8 (M,)."""
    fx, fy = flow_[:, 0], flow_[:, 1]
    modulus = np.sqrt(fx * fx + fy * fy)
    v = np.zeros(len(points_), np.uint8)
    for i, m in enumerate(modulus):
        v[i] = np.minimum(50 + m * 2, 255)
    return v


def hsv2bgr_u8(hsv):
    """cv2.cvtColor(hsv, COLOR_HSV2BGR) for uint8 input with H in [0, 180] (third-party: opencv-python, imgproc
    color_hsv.simd.hpp HSV2RGB_b / HSV2RGB_f).  float32 throughout: h * (6/180), s and v * (1/255), the four-entry
    table {v, v(1-s), v(1-s f), v(1-s(1-f))}, times 255 and TRUNCATED.  Pinned bit for bit against cv2 for all
    181 x 256 (h, v) at s = 255 (the only saturation draw_hsv writes) and s = 0 in tests/test_oracle_golden.py;
    other saturations differ from cv2's SIMD rounding by one level in places and are not used."""
    f32 = np.float32
    h = hsv[..., 0].astype(f32) * f32(6.0 / 180.0)
    s = hsv[..., 1].astype(f32) * f32(1.0 / 255.0)
    v = hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    h = np.where(h >= 6, h - f32(6), h)
    sector = np.floor(h)
    f = h - sector
    sector = sector.astype(np.int64) % 6
    one = f32(1)
    tab = np.stack([v, v * (one - s), v * (one - s * f), v * (one - s * (one - f))], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    idx = sd[sector]
    out = np.stack([np.take_along_axis(tab, idx[..., k:k + 1], -1)[..., 0] for k in range(3)], -1) * f32(255.0)
    return np.clip(np.trunc(out), 0, 255).astype(np.uint8)


def draw_hsv(flow_):
    """pathfinder_viewer.py:124-141: hue = direction, value = 4 x length, BGR uint8 (H,W,3).
    Returns (bgr, hsv) -- the reference returns bgr; hsv is kept for the tests."""
    h, w = flow_.shape[:2]
    fx, fy = flow_[:, :, 0], flow_[:, :, 1]
    ang = np.arctan2(fy, fx) + np.pi
    v = np.sqrt(fx * fx + fy * fy)
    hsv = np.zeros((h, w, 3), np.uint8)
    hsv[..., 0] = ang * (180 / np.pi / 2)
    hsv[..., 1] = 255
    hsv[..., 2] = np.minimum(v * 4, 255)
    return hsv2bgr_u8(hsv), hsv
