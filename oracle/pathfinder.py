"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Headless numpy restatement of the reference's own downstream logic -- this part IS
reference code (not cv2), restated without the drawing calls:

* grid generator          pathfinder_viewer.py:255-267 (DenseOF.py:166-180)
* vector filter           pathfinder_viewer.py:159-178 (get_flow_lk, after the LK call)
* danger-point intensity  pathfinder_viewer.py:210-217 (draw_sparse_lamps, before drawing)
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


def vector_filter(next_pts, points_, width, height):
    """pathfinder_viewer.py:159-178.  Returns (flow int32 (M,2), points int32 (M,2), mask bool (N,), modulus f32 (N,))."""
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
