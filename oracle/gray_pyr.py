"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the integer front-end the reference reaches through
``cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)`` (pathfinder_viewer.py:244, :280;
DenseOF.py:481, :510; SparseOF.py:28) and of the ``pyrDown`` chain that
``cv2.calcOpticalFlowPyrLK`` builds internally (pathfinder_viewer.py:156-158;
SparseOF.py:35-36).  Upstream arithmetic lives in opencv ``imgproc/color_rgb``,
``imgproc/pyramids.cpp`` and ``video/lkpyramid.cpp`` (third-party, not vendored;
restated from SURVEY.md App. A.1/A.2, bit-exact against cv2 4.13).
"""
import numpy as np


def reflect101(idx, n):
    """BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba), valid for any offset."""
    idx = np.asarray(idx)
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.mod(idx, period)
    return np.where(idx >= n, period - idx, idx)


def bgr2gray(bgr):
    """uint8 (H,W,3) BGR -> uint8 (H,W); 15-bit fixed point, SURVEY App. A.1."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def pyrdown_u8(img):
    """uint8 (H,W) -> uint8 ((H+1)//2,(W+1)//2); [1 4 6 4 1]^2/256, REFLECT_101, App. A.2."""
    h, w = img.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    k = (1, 4, 6, 4, 1)
    src = img.astype(np.int64)
    xs = 2 * np.arange(dw)
    rows = np.zeros((h, dw), np.int64)
    for t in range(5):
        rows += k[t] * src[:, reflect101(xs + t - 2, w)]
    ys = 2 * np.arange(dh)
    out = np.zeros((dh, dw), np.int64)
    for t in range(5):
        out += k[t] * rows[reflect101(ys + t - 2, h), :]
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img, win, max_level):
    """Level list as buildOpticalFlowPyramid: stops early when a level would be <= winSize.

    Returns (levels, effective_max_level).
    """
    levels = [img]
    ww, wh = win
    for _ in range(max_level):
        h, w = levels[-1].shape
        nh, nw = (h + 1) // 2, (w + 1) // 2
        if nw <= ww or nh <= wh:
            break
        levels.append(pyrdown_u8(levels[-1]))
    return levels, len(levels) - 1


def scharr_s16(img):
    """int16 (H,W,2) = (Ix, Iy): 3x3 Scharr, REFLECT_101, as calcScharrDeriv (App. A.4)."""
    h, w = img.shape
    s = img.astype(np.int32)
    ym = reflect101(np.arange(h) - 1, h)
    yp = reflect101(np.arange(h) + 1, h)
    t0 = (s[ym] + s[yp]) * 3 + s * 10
    t1 = s[yp] - s[ym]
    xm = reflect101(np.arange(w) - 1, w)
    xp = reflect101(np.arange(w) + 1, w)
    ix = t0[:, xp] - t0[:, xm]
    iy = (t1[:, xm] + t1[:, xp]) * 3 + t1 * 10
    return np.stack([ix, iy], -1).astype(np.int16)
