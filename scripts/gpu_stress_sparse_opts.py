"""Sparse path options against LIVE cv2 on random crops of the real pairs: LK with random window sizes, pyramid depths,
criteria forms, flags (initial flow, min-eigenvalue error) and thresholds; corner detector with random masks, block
sizes, Sobel apertures, Harris.  LK outliers are re-run on the oracle (see gpu_stress_vs_cv2.py)."""
import os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2
from oracle import pyrlk as olk
from bench import decode_png
z = np.load(os.path.join(ROOT, "tests/golden/real_sweep.npz"))
f = np.load(os.path.join(ROOT, "tests/golden/real_1080p.npz"))
pairs = [(decode_png(z[f"png0_{i}"]), decode_png(z[f"png1_{i}"])) for i in range(3)] + [(decode_png(f["png0"]), decode_png(f["png1"]))]
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
bad = 0
for c in range(N):
    g0, g1 = pairs[c % 4]
    h, w = int(rng.integers(40, 1080)), int(rng.integers(40, 1920))
    y0, x0 = int(rng.integers(0, 1080 - h + 1)), int(rng.integers(0, 1920 - w + 1))
    a, b = np.ascontiguousarray(g0[y0:y0 + h, x0:x0 + w]), np.ascontiguousarray(g1[y0:y0 + h, x0:x0 + w])
    if c % 3 == 0:                      # strided views, as cv2 accepts them
        a, b = g0[y0:y0 + h, x0:x0 + w], g1[y0:y0 + h, x0:x0 + w]
    n = int(rng.integers(1, 3000))
    pts = np.float32(np.stack([rng.uniform(-5, w + 5, n), rng.uniform(-5, h + 5, n)], 1))
    win = (int(rng.integers(3, 60)), int(rng.integers(3, 60)))
    lvl = int(rng.integers(0, 6))
    ctype = int(rng.choice([1, 2, 3]))
    crit = (ctype, int(rng.integers(1, 40)), float(rng.choice([0.001, 0.01, 0.03, 0.3])))
    flags = int(rng.choice([0, 0, 4, 8, 12]))
    thr = float(rng.choice([1e-4, 1e-4, 1e-3, 1e-2]))
    init = np.float32(pts + rng.normal(0, 3, pts.shape)) if flags & 4 else None
    kw = dict(winSize=win, maxLevel=lvl, criteria=crit, flags=flags, minEigThreshold=thr)
    try:
        wn, ws, we = cv2.calcOpticalFlowPyrLK(b, a, pts, None if init is None else init.copy(), **kw)
    except cv2.error as e:
        print("cv2 rejects", h, w, kw); continue
    try:
        gn, gs, ge = b2.calcOpticalFlowPyrLK(b, a, pts, None if init is None else init.copy(), **kw)
    except Exception as e:
        bad += 1; print("B200 RAISES", h, w, n, kw, repr(e)[:300], flush=True); continue
    d = np.abs(gn - wn).max(-1)
    both = (gs.ravel() == 1) & (ws.ravel() == 1)
    de = np.abs(ge - we).ravel()[both & (d <= 0.05)]
    off = np.where((d > 0.05) | (gs.ravel() != ws.ravel()))[0]
    line = "LK %4dx%-4d n %4d %s: status %.4f, pos within 0.05: %.4f, err max %.3g" % (h, w, n, kw, (gs == ws).mean(), (d <= 0.05).mean(), de.max() if len(de) else 0)
    if len(off):
        sel = off[:12]
        on, os_, oe = olk.pyrlk(b, a, pts[sel], None if init is None else init[sel].copy(), win, lvl, crit, flags, thr)
        do = np.abs(on - gn[sel]).max()
        same = np.array_equal(os_.ravel(), gs[sel].ravel())
        line += " | %d outliers, first %d vs oracle: max %.2e status %s" % (len(off), len(sel), do, same)
        if do > 1e-3 or not same:
            bad += 1; line = "DEFECT? " + line
    print(line, flush=True)
for c in range(N):
    g0 = pairs[c % 4][c // 4 % 2]
    h, w = int(rng.integers(16, 1080)), int(rng.integers(16, 1920))
    y0, x0 = int(rng.integers(0, 1080 - h + 1)), int(rng.integers(0, 1920 - w + 1))
    img = g0[y0:y0 + h, x0:x0 + w] if c % 2 else np.ascontiguousarray(g0[y0:y0 + h, x0:x0 + w])
    kw = dict(maxCorners=int(rng.choice([0, 1, 20, 100, 500, 3000])), qualityLevel=float(rng.choice([0.3, 0.1, 0.03, 0.01])),
              minDistance=float(rng.choice([0, 1, 3.5, 10, 10.5, 40])), blockSize=int(rng.choice([3, 5, 7, 9, 15])))
    extra = {}
    if rng.random() < 0.3:
        extra["useHarrisDetector"] = True; extra["k"] = float(rng.choice([0.04, 0.06]))
    gs_ = int(rng.choice([3, 3, 5, 7]))
    mask = None
    if rng.random() < 0.5:
        mask = (rng.random((h, w)) < rng.random()).astype(np.uint8) * 255
    try:
        want = cv2.goodFeaturesToTrack(img, mask=mask, gradientSize=gs_, **kw, **extra) if gs_ != 3 else cv2.goodFeaturesToTrack(img, mask=mask, **kw, **extra)
    except cv2.error:
        print("cv2 rejects gftt", h, w, kw, extra, gs_); continue
    try:
        got = b2.goodFeaturesToTrack(img, mask=mask, gradientSize=gs_, **kw, **extra)
    except Exception as e:
        bad += 1; print("B200 RAISES gftt", h, w, kw, extra, gs_, repr(e)[:300], flush=True); continue
    if want is None or got is None:
        ok = want is None and got is None
        msg = "none" if ok else "ONE IS NONE"
    else:
        A, C = set(map(tuple, got.reshape(-1, 2))), set(map(tuple, want.reshape(-1, 2)))
        exact = got.shape == want.shape and np.array_equal(got, want)
        ok = exact or (got.shape == want.shape and len(A ^ C) <= max(2, len(C) // 100))
        msg = "exact" if exact else "n %d/%d, set difference %d" % (len(got), len(want), len(A ^ C))
    if not ok:
        bad += 1
    print("%sGFTT %4dx%-4d %s %s grad %d mask %s: %s" % ("" if ok else "DEFECT? ", h, w, kw, extra, gs_, mask is not None, msg), flush=True)
print("suspect cases", bad)
