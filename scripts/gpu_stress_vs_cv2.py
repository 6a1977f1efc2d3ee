"""Stress of the sparse path against LIVE cv2 on the GPU box, on the committed full-resolution real pairs: tens of
thousands of random sub-pixel points per pair and window size (positions, statuses), corner lists for several parameter
sets.  Outliers are re-run on the numpy oracle to tell a CUDA-path defect (differs from the oracle) from cv2's float
accumulation noise in near-singular windows (equals the oracle)."""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(123)
crit = (3, 10, 0.03)
for ci, (g0, g1) in enumerate(pairs):
    pts = np.float32(np.stack([rng.uniform(-20, 1940, N), rng.uniform(-20, 1100, N)], 1))
    for win, lvl in [((45, 45), 2), ((15, 15), 2), ((21, 21), 3)]:
        wn, ws, we = cv2.calcOpticalFlowPyrLK(g1, g0, pts, None, winSize=win, maxLevel=lvl, criteria=crit)
        gn, gs, ge = b2.calcOpticalFlowPyrLK(g1, g0, pts, None, winSize=win, maxLevel=lvl, criteria=crit)
        d = np.abs(gn - wn).max(-1)
        off = np.where((d > 0.05) | (gs.ravel() != ws.ravel()))[0]
        msg = "clip %d win %s: status agree %.5f, within 0.05 px %.5f, outliers %d" % (ci, win, (gs == ws).mean(), (d <= 0.05).mean(), len(off))
        if len(off):
            sel = off[:int(os.environ.get('STRESS_ORACLE', 6))]
            on, os_, oe = olk.pyrlk(g1, g0, pts[sel], None, win, lvl, crit)
            do = np.abs(on - gn[sel]).max(-1)
            msg += "; first %d vs oracle: max %.2e, status equal %s, median flow length %.0f" % (
                len(sel), do.max(), np.array_equal(os_.ravel(), gs[sel].ravel()), np.median(np.linalg.norm(wn[sel] - pts[sel], axis=1)))
        print(msg, flush=True)
    for name, kw in {"sparse": dict(maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7),
                     "dense": dict(maxCorners=500, qualityLevel=0.01, minDistance=5, blockSize=3),
                     "bs5": dict(maxCorners=200, qualityLevel=0.05, minDistance=8, blockSize=5),
                     "harris": dict(maxCorners=100, qualityLevel=0.05, minDistance=7, blockSize=3, useHarrisDetector=True, k=0.04)}.items():
        for img in (g0, g1):
            w = cv2.goodFeaturesToTrack(img, mask=None, **kw)
            g = b2.goodFeaturesToTrack(img, mask=None, **kw)
            same = (w is None and g is None) or (w is not None and g is not None and np.array_equal(w, g))
            if not same:
                print("clip", ci, "GFTT", name, "DIFFERS", None if w is None else w.shape, None if g is None else g.shape, flush=True)
    print("clip", ci, "gftt sets done", flush=True)
