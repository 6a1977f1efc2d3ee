"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of ``cv2.goodFeaturesToTrack`` as the reference calls it
(SparseOF.py:69 with feature_params SparseOF.py:10-13: maxCorners=20,
qualityLevel=0.3, minDistance=10, blockSize=7, plus the disc mask built at
SparseOF.py:61-66).  Arithmetic lives in opencv ``imgproc/src/corner.cpp`` and
``featureselect.cpp`` (third-party, not vendored); restated from SURVEY.md
App. A.5.
"""
import numpy as np

from .gray_pyr import reflect101

f32 = np.float32


def sobel3(img):
    """Integer 3x3 Sobel (dx, dy), REFLECT_101."""
    h, w = img.shape
    s = img.astype(np.int32)
    ym, yp = reflect101(np.arange(h) - 1, h), reflect101(np.arange(h) + 1, h)
    xm, xp = reflect101(np.arange(w) - 1, w), reflect101(np.arange(w) + 1, w)
    sm_y = s[ym] + 2 * s + s[yp]
    df_y = s[yp] - s[ym]
    dx = sm_y[:, xp] - sm_y[:, xm]
    dy = df_y[:, xm] + 2 * df_y + df_y[:, xp]
    return dx, dy


def _fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)


def sobel3_f32(img, scale):
    """cv2.Sobel(CV_32F, ksize=3, scale) as the SIMD body of opencv's separable filter rounds it (probed against
    cv2 4.13 bit for bit): the scaled smoothing kernel [s, 2s, s] is applied with fused multiply-adds,
        Dx = fma(t[y-1] + t[y+1], s, t[y] * 2s),        t = I[x+1] - I[x-1]   (exact)
        Dy = r[y+1] - r[y-1],   r = fma(I[x+1], s, fma(I[x], 2s, I[x-1] * s))
    (cv2's scalar tail columns -- width mod SIMD width -- round differently; that is machine dependent)."""
    h, w = img.shape
    s = img.astype(f32)
    ym, yp = reflect101(np.arange(h) - 1, h), reflect101(np.arange(h) + 1, h)
    xm, xp = reflect101(np.arange(w) - 1, w), reflect101(np.arange(w) + 1, w)
    s2 = f32(2) * scale
    t = s[:, xp] - s[:, xm]
    dx = _fma(t[ym] + t[yp], scale, t * s2)
    r = _fma(s[:, xp], scale, _fma(s, s2, s[:, xm] * scale))
    dy = r[yp] - r[ym]
    return dx, dy


SOBEL_TAPS = {3: ([-1, 0, 1], [1, 2, 1]), 5: ([-1, -2, 0, 2, 1], [1, 4, 6, 4, 1]),
              7: ([-1, -4, -5, 0, 5, 4, 1], [1, 6, 15, 20, 15, 6, 1])}


def sobel_f32(img, ksize, scale):
    """cv2.Sobel(CV_32F, ksize in {3, 5, 7}, scale) as opencv's separable filter computes it (probed bit for bit
    against cv2 4.13 on widths without a scalar tail).  cv::Sobel scales the SMOOTHING kernel (float32 taps times
    float32 scale); sepFilter2D then runs a float32 row filter and a symmetric / antisymmetric column filter:
        Dx: row = derivative taps (integers: exact), column = k[r] * t[y], then fma(t[y+j] + t[y-j], k[r+j], .)
        Dy: row = scaled smoothing taps, fused multiply-adds left to right; column = (r[y+1] - r[y-1]) * d[r+1],
            then fma(r[y+j] - r[y-j], d[r+j], .)"""
    d, sm = SOBEL_TAPS[ksize]
    r = ksize // 2
    h, w = img.shape
    s = img.astype(f32)
    ys, xs = np.arange(h), np.arange(w)
    cols = [reflect101(xs + j - r, w) for j in range(ksize)]
    rows = [reflect101(ys + j - r, h) for j in range(ksize)]
    k = [f32(f32(v) * f32(scale)) for v in sm]
    t = np.zeros((h, w), f32)
    for j in range(ksize):
        t = t + f32(d[j]) * s[:, cols[j]]
    dx = t * k[r]
    for j in range(1, r + 1):
        dx = _fma(t[rows[r + j]] + t[rows[r - j]], k[r + j], dx)
    rr = s[:, cols[0]] * k[0]
    for j in range(1, ksize):
        rr = _fma(s[:, cols[j]], k[j], rr)
    dy = (rr[rows[r + 1]] - rr[rows[r - 1]]) * f32(d[r + 1])
    for j in range(2, r + 1):
        dy = _fma(rr[rows[r + j]] - rr[rows[r - j]], f32(d[r + j]), dy)
    return dx, dy


def min_eig_map(img, block_size=3, gradient_size=3, harris=False, k=0.04):
    """cornerMinEigenVal / cornerHarris response, float32 (H,W)."""
    assert gradient_size in (3, 5, 7)
    h, w = img.shape
    scale = f32(1.0 / ((1 << (gradient_size - 1)) * block_size * 255.0))
    Dx, Dy = sobel3_f32(img, scale) if gradient_size == 3 else sobel_f32(img, gradient_size, scale)
    cov = np.stack([Dx * Dx, Dx * Dy, Dy * Dy], -1).astype(np.float64)
    r = block_size // 2
    ys, xs = np.arange(h), np.arange(w)
    v = np.zeros_like(cov)
    for t in range(block_size):
        v += cov[reflect101(ys + t - r, h)]
    o = np.zeros_like(cov)
    for t in range(block_size):
        o += v[:, reflect101(xs + t - r, w)]
    o = o.astype(f32)
    if harris:
        a, b, c = o[..., 0], o[..., 1], o[..., 2]
        # calcHarris as cv2's SIMD body rounds it (probed bit for bit): float32, k times the SQUARED trace
        return ((a * c - b * b) - f32(k) * ((a + c) * (a + c))).astype(f32)
    a = o[..., 0] * f32(0.5)
    b = o[..., 1]
    c = o[..., 2] * f32(0.5)
    return ((a + c) - np.sqrt((a - c) * (a - c) + b * b)).astype(f32)


def good_features_to_track(img, max_corners, quality, min_dist, mask=None, block_size=3,
                           gradient_size=3, harris=False, k=0.04):
    """Returns float32 (n,1,2) of (x,y) or None."""
    h, w = img.shape
    eig = min_eig_map(img, block_size, gradient_size, harris, k)
    if mask is not None:
        sel = mask != 0
        max_val = eig[sel].max() if sel.any() else f32(0)
    else:
        max_val = eig.max()
    thr = f32(np.float64(max_val) * quality)
    eig = np.where(eig > thr, eig, f32(0))
    pad = np.pad(eig, 1, mode="constant", constant_values=-np.inf)
    dil = np.max(np.stack([pad[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)]), 0)
    cand = (eig != 0) & (eig == dil)
    if mask is not None:
        cand &= mask != 0
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    if len(ys) == 0:
        return None
    vals = eig[ys, xs]
    lin = ys.astype(np.int64) * w + xs
    order = np.lexsort((-lin, -vals.astype(np.float64)))
    ys, xs = ys[order], xs[order]
    out = []
    if min_dist >= 1:
        cell = int(np.rint(min_dist))
        gw, gh = (w + cell - 1) // cell, (h + cell - 1) // cell
        grid = {}
        md2 = min_dist * min_dist
        for y, x in zip(ys, xs):
            cx, cy = x // cell, y // cell
            good = True
            for yy in range(max(cy - 1, 0), min(cy + 1, gh - 1) + 1):
                for xx in range(max(cx - 1, 0), min(cx + 1, gw - 1) + 1):
                    for (qx, qy) in grid.get((xx, yy), ()):
                        ddx, ddy = x - qx, y - qy
                        if ddx * ddx + ddy * ddy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid.setdefault((cx, cy), []).append((x, y))
                out.append((x, y))
                if max_corners > 0 and len(out) == max_corners:
                    break
    else:
        for y, x in zip(ys, xs):
            out.append((x, y))
            if max_corners > 0 and len(out) == max_corners:
                break
    return np.array(out, np.float32).reshape(-1, 1, 2)
