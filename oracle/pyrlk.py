"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of ``cv2.calcOpticalFlowPyrLK`` as the reference calls it:
grid form pathfinder_viewer.py:153-158 / DenseOF.py:181-185 (winSize 45x45,
maxLevel 2, criteria (EPS|COUNT, 10, 0.03), prev = current frame, next = previous
frame) and track form SparseOF.py:35-36 (winSize 15x15, forward + backward).
Arithmetic lives in opencv ``video/src/lkpyramid.cpp`` (third-party, not
vendored); restated from SURVEY.md App. A.4.  Python loop over points, numpy over
the window: use on a few hundred points at most.

Window sums are taken exactly (int64) where cv2 accumulates float lanes; the
survey measured <= 7e-4 px position difference from that on well-conditioned
windows.  ``accum="cv2_simd128"`` restates cv2's accumulation order instead (the
CV_SIMD128 branch of LKTrackerInvoker: four float lanes over groups of eight
columns, a scalar float tail, lanes folded at the end) -- used by the tests to
show that where the exact sums and cv2 part ways on real footage (diverging
tracks through near-singular windows) it is this float rounding that does it.
"""
import numpy as np

from .gray_pyr import build_pyramid, scharr_s16

COUNT, EPS = 1, 2
USE_INITIAL_FLOW = 4
GET_MIN_EIGENVALS = 8
f32 = np.float32
FLT_SCALE = f32(1.0 / (1 << 20))
FLT_EPSILON = np.finfo(np.float32).eps


def _weights(a, b):
    iw00 = int(np.rint((f32(1) - a) * (f32(1) - b) * f32(16384)))
    iw01 = int(np.rint(a * (f32(1) - b) * f32(16384)))
    iw10 = int(np.rint((f32(1) - a) * b * f32(16384)))
    return iw00, iw01, iw10, 16384 - iw00 - iw01 - iw10


def _seq_f32(v):
    """Sequential float32 accumulation of v (np.add.accumulate adds left to right)."""
    v = np.asarray(v, np.float32).ravel()
    return np.add.accumulate(v, dtype=np.float32)[-1] if len(v) else f32(0)


def _lane_sums_A(p):
    """sum of the exact per-pixel products p (wh, ww) in cv2's order: lane l of a float32x4 takes columns 8g + l and
    8g + 4 + l of every group of eight, row after row; the columns past the last full group go to a scalar float."""
    wh, ww = p.shape
    g8 = ww // 8 * 8
    lanes = [_seq_f32(p[:, :g8].reshape(wh, g8 // 8, 2, 4)[:, :, :, l]) for l in range(4)]
    tail = _seq_f32(p[:, g8:])
    return f32(tail + f32(f32(f32(lanes[0] + lanes[1]) + lanes[2]) + lanes[3]))


def _lane_sums_b(p):
    """same for the mismatch vector: a lane adds float(int32 p[8g + l] + p[8g + 4 + l]); lanes 0 / 2 of qb0 and qb1 are
    folded as (qb0 + qb1)[0] + (qb0 + qb1)[2]."""
    wh, ww = p.shape
    g8 = ww // 8 * 8
    q = p[:, :g8].reshape(wh, g8 // 8, 2, 4)
    lanes = [_seq_f32((q[:, :, 0, l] + q[:, :, 1, l]).astype(np.float32)) for l in range(4)]
    tail = _seq_f32(p[:, g8:])
    return f32(tail + f32(f32(lanes[0] + lanes[2]) + f32(lanes[1] + lanes[3])))


def _bilin(P, x0, y0, ww, wh, w4):
    """Integer bilinear over a ww x wh window whose top-left integer corner is (x0,y0) in P's frame."""
    a = P[y0:y0 + wh + 1, x0:x0 + ww + 1].astype(np.int64)
    return a[:-1, :-1] * w4[0] + a[:-1, 1:] * w4[1] + a[1:, :-1] * w4[2] + a[1:, 1:] * w4[3]


def _descale(v, n):
    return (v + (1 << (n - 1))) >> n


def pyrlk(prev_img, next_img, prev_pts, next_pts=None, win=(21, 21), max_level=3,
          criteria=(COUNT | EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4, accum="exact"):
    """Returns (nextPts shaped like prev_pts, status uint8 (N,1), err float32 (N,1))."""
    ww, wh = win
    ctype, max_count, eps = criteria
    max_count = min(max(int(max_count), 0), 100) if ctype & COUNT else 30
    eps = min(max(float(eps), 0.0), 10.0) if ctype & EPS else 0.01   # probed: cv2 COUNT-only == (COUNT|EPS, .., 0.01)
    eps2 = eps * eps
    pts = np.asarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(pts)
    if flags & USE_INITIAL_FLOW:
        nxt = np.asarray(next_pts, np.float32).reshape(-1, 2).copy()
    else:
        nxt = np.zeros((n, 2), np.float32)
    status = np.ones(n, np.uint8)
    err = np.zeros(n, np.float32)
    pyr_i, lv_i = build_pyramid(prev_img, win, max_level)
    pyr_j, lv_j = build_pyramid(next_img, win, max_level)
    max_level = min(lv_i, lv_j)
    half = np.array([(ww - 1) * 0.5, (wh - 1) * 0.5], np.float32)
    px, py = ww + 2, wh + 2  # padding used by this restatement (cv2 pads by winSize)
    for level in range(max_level, -1, -1):
        I, J = pyr_i[level], pyr_j[level]
        H, W = I.shape
        Ip = np.pad(I, ((py, py), (px, px)), mode="reflect")
        Jp = np.pad(J, ((py, py), (px, px)), mode="reflect")
        D = scharr_s16(I)
        Dx = np.pad(D[..., 0], ((py, py), (px, px)))
        Dy = np.pad(D[..., 1], ((py, py), (px, px)))
        inv_scale = f32(1.0 / (1 << level))
        for i in range(n):
            prev_pt = pts[i] * inv_scale
            if level == max_level:
                next_pt = nxt[i] * inv_scale if flags & USE_INITIAL_FLOW else prev_pt.copy()
            else:
                next_pt = nxt[i] * f32(2)
            nxt[i] = next_pt
            prev_pt = prev_pt - half
            ip = np.floor(prev_pt).astype(np.int64)
            if ip[0] < -ww or ip[0] >= W or ip[1] < -wh or ip[1] >= H:
                if level == 0:
                    status[i] = 0
                    err[i] = 0
                continue
            w4 = _weights(f32(prev_pt[0] - ip[0]), f32(prev_pt[1] - ip[1]))
            x0, y0 = int(ip[0]) + px, int(ip[1]) + py
            Iw = _descale(_bilin(Ip, x0, y0, ww, wh, w4), 9)
            Ixw = _descale(_bilin(Dx, x0, y0, ww, wh, w4), 14)
            Iyw = _descale(_bilin(Dy, x0, y0, ww, wh, w4), 14)
            sumA = _lane_sums_A if accum == "cv2_simd128" else (lambda p: f32(p.sum()))
            sumb = _lane_sums_b if accum == "cv2_simd128" else (lambda p: f32(p.sum()))
            A11 = f32(sumA(Ixw * Ixw) * FLT_SCALE)
            A12 = f32(sumA(Ixw * Iyw) * FLT_SCALE)
            A22 = f32(sumA(Iyw * Iyw) * FLT_SCALE)
            Dt = f32(A11 * A22 - A12 * A12)
            min_eig = f32((A22 + A11 - np.sqrt(f32((A11 - A22) * (A11 - A22) + f32(4) * A12 * A12))) / f32(2 * ww * wh))
            if flags & GET_MIN_EIGENVALS:
                err[i] = min_eig
            if min_eig < min_eig_threshold or Dt < FLT_EPSILON:
                if level == 0:
                    status[i] = 0
                continue
            Dt = f32(1.0) / Dt
            next_pt = next_pt - half
            prev_delta = np.zeros(2, np.float32)
            for j in range(max_count):
                inx = np.floor(next_pt).astype(np.int64)
                if inx[0] < -ww or inx[0] >= W or inx[1] < -wh or inx[1] >= H:
                    if level == 0:
                        status[i] = 0
                    break
                w4j = _weights(f32(next_pt[0] - inx[0]), f32(next_pt[1] - inx[1]))
                diff = _descale(_bilin(Jp, int(inx[0]) + px, int(inx[1]) + py, ww, wh, w4j), 9) - Iw
                b1 = f32(sumb(diff * Ixw) * FLT_SCALE)
                b2 = f32(sumb(diff * Iyw) * FLT_SCALE)
                delta = np.array([f32(f32(A12 * b2 - A22 * b1) * Dt), f32(f32(A12 * b1 - A11 * b2) * Dt)], np.float32)
                next_pt = next_pt + delta
                nxt[i] = next_pt + half
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps2:
                    break
                if j > 0 and abs(delta[0] + prev_delta[0]) < 0.01 and abs(delta[1] + prev_delta[1]) < 0.01:
                    nxt[i] = nxt[i] - delta * f32(0.5)
                    break
                prev_delta = delta
            if status[i] and level == 0 and not (flags & GET_MIN_EIGENVALS):
                p = nxt[i] - half
                ipt = np.floor(p).astype(np.int64)
                if ipt[0] < -ww or ipt[0] >= W or ipt[1] < -wh or ipt[1] >= H:
                    status[i] = 0
                    continue
                w4e = _weights(f32(p[0] - ipt[0]), f32(p[1] - ipt[1]))
                diff = _descale(_bilin(Jp, int(ipt[0]) + px, int(ipt[1]) + py, ww, wh, w4e), 9) - Iw
                err[i] = f32(f32(np.abs(diff).sum()) / f32(32 * ww * wh))
    return nxt.reshape(np.asarray(prev_pts).shape), status.reshape(-1, 1), err.reshape(-1, 1)
