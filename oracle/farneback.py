"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of ``cv2.calcOpticalFlowFarneback`` as the reference calls it
(DenseOF.py:127-157, call site :520; parameters 0.5/3/15/3/5/1.2/0).  The
arithmetic lives in opencv ``video/src/optflowgf.cpp`` plus ``GaussianBlur`` and
``resize`` from ``imgproc`` (third-party, not vendored); restated from SURVEY.md
App. A.3 and pinned against the live cv2 4.13 wheel in
tests/test_oracle_farneback.py.

cv2 keeps R/M/flow in float32 and accumulates the horizontal PolyExp pass, the
box sums and the 2x2 solve in double; so does this restatement.
"""
import numpy as np

from .gray_pyr import reflect101

OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_FARNEBACK_GAUSSIAN = 256
MIN_SIZE = 32


def cv_round(x):
    return int(np.rint(x))


def gaussian_kernel(n, sigma):
    """cv::getGaussianKernel(n, sigma, CV_32F)."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if sigma <= 0 and n in small:
        return np.array(small[n], np.float32)
    s = sigma if sigma > 0 else ((n - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * s * s))
    return (k / k.sum()).astype(np.float32)


def level_plan(width, height, pyr_scale, levels):
    """[(k, scale, sigma, ksize, w, h)] from coarsest to finest (App. A.3 head)."""
    k, scale = 0, 1.0
    while k < levels:
        scale *= pyr_scale
        if width * scale < MIN_SIZE or height * scale < MIN_SIZE:
            break
        k += 1
    plan = []
    for lv in range(k, -1, -1):
        scale = 1.0
        for _ in range(lv):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1.0) * 0.5
        ksz = max(cv_round(sigma * 5) | 1, 3)
        plan.append((lv, scale, sigma, ksz, cv_round(width * scale), cv_round(height * scale)))
    return plan


def gaussian_blur_f32(img, ksz, sigma):
    """float32 separable blur, BORDER_REFLECT_101, horizontal then vertical."""
    k = gaussian_kernel(ksz, sigma)
    h, w = img.shape
    r = ksz // 2
    xs = np.arange(w)
    tmp = np.zeros((h, w), np.float32)
    for t in range(ksz):
        tmp += k[t] * img[:, reflect101(xs + t - r, w)]
    ys = np.arange(h)
    out = np.zeros((h, w), np.float32)
    for t in range(ksz):
        out += k[t] * tmp[reflect101(ys + t - r, h), :]
    return out


def _linear_coords(dst_n, src_n):
    scale = np.float64(src_n) / dst_n
    f = ((np.arange(dst_n) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0
    hi = s >= src_n - 1
    s[hi] = src_n - 1
    f[hi] = 0
    s1 = np.minimum(s + 1, src_n - 1)
    return s, s1, f


def resize_linear_f32(src, w, h):
    """cv2.resize(..., INTER_LINEAR) for float32 (H,W) or (H,W,C)."""
    sh, sw = src.shape[:2]
    if (sw, sh) == (w, h):
        return src.copy()
    x0, x1, fx = _linear_coords(w, sw)
    y0, y1, fy = _linear_coords(h, sh)
    if src.ndim == 3:
        fx = fx[None, :, None]
        fy = fy[:, None, None]
    else:
        fx = fx[None, :]
        fy = fy[:, None]
    top = src[y0][:, x0] * (1 - fx) + src[y0][:, x1] * fx
    bot = src[y1][:, x0] * (1 - fx) + src[y1][:, x1] * fx
    return (top * (1 - fy) + bot * fy).astype(np.float32)


def level_image(img_u8, ksz, sigma, w, h):
    """K3: convertTo(f32) -> GaussianBlur on the FULL-RES frame -> resize(INTER_LINEAR)."""
    f = gaussian_blur_f32(img_u8.astype(np.float32), ksz, sigma)
    return resize_linear_f32(f, w, h)


def polyexp_constants(n, sigma):
    """(g, xg, xxg float32[n+1], ig11, ig03, ig33, ig55 double) -- FarnebackPrepareGaussian."""
    if sigma < np.finfo(np.float32).eps:
        sigma = n * 0.3
    x = np.arange(-n, n + 1)
    g = np.exp(-(x * x) / (2.0 * sigma * sigma)).astype(np.float32)
    s = np.float64(g.astype(np.float64).sum())
    g = (g / s).astype(np.float32)
    xg = (x * g).astype(np.float32)
    xxg = (x * x * g).astype(np.float32)
    G = np.zeros((6, 6), np.float64)
    for yy in range(-n, n + 1):
        for xx in range(-n, n + 1):
            gg = np.float64(g[yy + n]) * np.float64(g[xx + n])
            G[0, 0] += gg
            G[1, 1] += gg * xx * xx
            G[3, 3] += gg * xx * xx * xx * xx
            G[5, 5] += gg * xx * xx * yy * yy
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g[n:], xg[n:], xxg[n:], inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def polyexp(img, n, sigma):
    """K4: FarnebackPolyExp -> float32 (h,w,5); REPLICATE borders."""
    g, xg, xxg, ig11, ig03, ig33, ig55 = polyexp_constants(n, sigma)
    h, w = img.shape
    ys = np.arange(h)
    row0 = img * g[0]
    row1 = np.zeros_like(img)
    row2 = np.zeros_like(img)
    for k in range(1, n + 1):
        s0 = img[np.maximum(ys - k, 0)]
        s1 = img[np.minimum(ys + k, h - 1)]
        row0 = row0 + g[k] * (s0 + s1)
        row1 = row1 + xg[k] * (s1 - s0)
        row2 = row2 + xxg[k] * (s0 + s1)
    xs = np.arange(w)
    g64, xg64, xxg64 = g.astype(np.float64), xg.astype(np.float64), xxg.astype(np.float64)
    b1 = row0.astype(np.float64) * g64[0]
    b3 = row1.astype(np.float64) * g64[0]
    b5 = row2.astype(np.float64) * g64[0]
    b2 = np.zeros((h, w))
    b4 = np.zeros((h, w))
    b6 = np.zeros((h, w))
    for k in range(1, n + 1):
        xp = np.minimum(xs + k, w - 1)
        xm = np.maximum(xs - k, 0)
        tg = (row0[:, xp] + row0[:, xm]).astype(np.float64)
        b1 += tg * g64[k]
        b4 += tg * xxg64[k]
        b2 += (row0[:, xp] - row0[:, xm]).astype(np.float64) * xg64[k]
        b3 += (row1[:, xp] + row1[:, xm]).astype(np.float64) * g64[k]
        b6 += (row1[:, xp] - row1[:, xm]).astype(np.float64) * xg64[k]
        b5 += (row2[:, xp] + row2[:, xm]).astype(np.float64) * g64[k]
    out = np.empty((h, w, 5), np.float32)
    out[..., 0] = b3 * ig11
    out[..., 1] = b2 * ig11
    out[..., 2] = b1 * ig03 + b5 * ig33
    out[..., 3] = b1 * ig03 + b4 * ig33
    out[..., 4] = b6 * ig55
    return out


BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], np.float32)


def update_matrices(R0, R1, flow):
    """K5a: FarnebackUpdateMatrices -> float32 (h,w,5)."""
    h, w = flow.shape[:2]
    f32 = np.float32
    xs = np.arange(w, dtype=np.float32)[None, :]
    ys = np.arange(h, dtype=np.float32)[:, None]
    dx, dy = flow[..., 0], flow[..., 1]
    fx = xs + dx
    fy = ys + dy
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1).astype(f32)
    fy = (fy - y1).astype(f32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc = np.clip(x1, 0, max(w - 2, 0))
    yc = np.clip(y1, 0, max(h - 2, 0))
    a00 = ((1 - fx) * (1 - fy))[..., None]
    a01 = (fx * (1 - fy))[..., None]
    a10 = ((1 - fx) * fy)[..., None]
    a11 = (fx * fy)[..., None]
    xc1 = np.minimum(xc + 1, w - 1)
    yc1 = np.minimum(yc + 1, h - 1)
    r = a00 * R1[yc, xc] + a01 * R1[yc, xc1] + a10 * R1[yc1, xc] + a11 * R1[yc1, xc1]
    r = r.astype(f32)
    r2 = np.where(inside, r[..., 0], 0).astype(f32)
    r3 = np.where(inside, r[..., 1], 0).astype(f32)
    r4 = np.where(inside, (R0[..., 2] + r[..., 2]) * f32(0.5), R0[..., 2]).astype(f32)
    r5 = np.where(inside, (R0[..., 3] + r[..., 3]) * f32(0.5), R0[..., 3]).astype(f32)
    r6 = np.where(inside, (R0[..., 4] + r[..., 4]) * f32(0.25), R0[..., 4] * f32(0.5)).astype(f32)
    r2 = (R0[..., 0] - r2) * f32(0.5)
    r3 = (R0[..., 1] - r3) * f32(0.5)
    r2 = r2 + r4 * dy + r6 * dx
    r3 = r3 + r6 * dy + r5 * dx
    sx = np.ones(w, f32)
    sy = np.ones(h, f32)
    for i in range(min(5, w)):
        sx[i] *= BORDER[i]
        sx[w - 1 - i] *= BORDER[i]
    for i in range(min(5, h)):
        sy[i] *= BORDER[i]
        sy[h - 1 - i] *= BORDER[i]
    s = sy[:, None] * sx[None, :]
    # optflowgf.cpp gates the attenuation with unsigned comparisons that wrap when a dimension is below 10 px
    u32 = lambda v: np.asarray(v, np.int64) & 0xFFFFFFFF
    gate = (u32(np.arange(w) - 5) >= u32(w - 10))[None, :] | (u32(np.arange(h) - 5) >= u32(h - 10))[:, None]
    s = np.where(gate, s, f32(1))
    r2, r3, r4, r5, r6 = r2 * s, r3 * s, r4 * s, r5 * s, r6 * s
    M = np.empty((h, w, 5), f32)
    M[..., 0] = r4 * r4 + r6 * r6
    M[..., 1] = (r4 + r5) * r6
    M[..., 2] = r5 * r5 + r6 * r6
    M[..., 3] = r4 * r2 + r6 * r3
    M[..., 4] = r6 * r2 + r5 * r3
    return M


def _window_sum_replicate(M, m, weights=None, dtype=np.float64):
    """Separable (2m+1)^2 window sum with BORDER_REPLICATE (vertical then horizontal)."""
    h, w = M.shape[:2]
    ys = np.arange(h)
    xs = np.arange(w)
    Md = M.astype(dtype)
    v = np.zeros_like(Md)
    for t in range(-m, m + 1):
        wt = 1.0 if weights is None else weights[abs(t)]
        v += dtype(wt) * Md[np.clip(ys + t, 0, h - 1)]
    o = np.zeros_like(Md)
    for t in range(-m, m + 1):
        wt = 1.0 if weights is None else weights[abs(t)]
        o += dtype(wt) * v[:, np.clip(xs + t, 0, w - 1)]
    return o


def blur_solve(M, winsize, gaussian=False):
    """K5b: FarnebackUpdateFlow_Blur / _GaussianBlur -> float32 (h,w,2)."""
    m = winsize // 2
    if not gaussian:
        b = _window_sum_replicate(M, m) * (1.0 / (winsize * winsize))
        g11, g12, g22, h1, h2 = (b[..., i] for i in range(5))
        idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
        fx = (g11 * h2 - g12 * h1) * idet
        fy = (g22 * h1 - g12 * h2) * idet
    else:
        sigma = m * 0.3
        k = np.exp(-(np.arange(m + 1, dtype=np.float32) ** 2) / np.float32(2 * sigma * sigma)).astype(np.float32)
        k[0] = 1.0
        s = np.float32(1.0) / (np.float32(1.0) + 2 * k[1:].sum(dtype=np.float32))
        k = (k * s).astype(np.float32)
        b = _window_sum_replicate(M, m, weights=k, dtype=np.float32)
        g11, g12, g22, h1, h2 = (b[..., i] for i in range(5))
        idet = np.float32(1.0) / (g11 * g22 - g12 * g12 + np.float32(1e-3))
        fx = (g11 * h2 - g12 * h1) * idet
        fy = (g22 * h1 - g12 * h2) * idet
    return np.stack([fx, fy], -1).astype(np.float32)


def farneback(prev, nxt, flow=None, pyr_scale=0.5, levels=3, winsize=15, iterations=3,
              poly_n=5, poly_sigma=1.2, flags=0, trace=None):
    """Full restatement; ``trace`` (a dict) receives per-level intermediates when given."""
    H, W = prev.shape
    gaussian = bool(flags & OPTFLOW_FARNEBACK_GAUSSIAN)
    prev_flow = None
    cur = None
    for (lv, scale, sigma, ksz, w, h) in level_plan(W, H, pyr_scale, levels):
        if prev_flow is None:
            if flags & OPTFLOW_USE_INITIAL_FLOW:
                import cv2  # INTER_AREA of the caller's flow: delegated, not on the reference's path
                cur = (cv2.resize(flow, (w, h), interpolation=cv2.INTER_AREA) * np.float32(scale)).astype(np.float32)
                cur = cur.reshape(h, w, 2)
            else:
                cur = np.zeros((h, w, 2), np.float32)
        else:
            cur = (resize_linear_f32(prev_flow, w, h) * np.float32(1.0 / pyr_scale)).astype(np.float32)
        I = [level_image(img, ksz, sigma, w, h) for img in (prev, nxt)]
        R = [polyexp(i, poly_n, poly_sigma) for i in I]
        if trace is not None:
            trace[lv] = {"I": I, "R": R, "flow_in": cur.copy(), "flow_it": []}
        M = update_matrices(R[0], R[1], cur)
        for it in range(iterations):
            cur = blur_solve(M, winsize, gaussian)
            if trace is not None:
                trace[lv]["flow_it"].append(cur.copy())
            if it < iterations - 1:
                M = update_matrices(R[0], R[1], cur)
        prev_flow = cur
    return cur
