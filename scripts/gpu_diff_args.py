"""Differential check of argument handling: the same odd calls made on live cv2 and on the drop-in; prints where one
raises and the other does not, or where both succeed with different results."""
import os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2, synth
fr = synth.sequence(120, 160, 2, seed=3)
a, b = fr[0], fr[1]
bgr = synth.sequence(120, 160, 1, seed=3, gray=False)[0]
pts = np.float32([[20, 20], [80.5, 60.25], [150, 110], [5, 5]])
FB = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def run(mod, f):
    try:
        return "ok", f(mod)
    except Exception as e:
        return "error", type(e).__name__ + ": " + str(e)[:80]


def close(x, y):
    if x is None or y is None:
        return x is None and y is None
    if isinstance(x, tuple):
        return all(close(i, j) for i, j in zip(x, y))
    x, y = np.asarray(x), np.asarray(y)
    return x.shape == y.shape and x.dtype == y.dtype and np.allclose(x.astype(np.float64), y.astype(np.float64), atol=0.05, equal_nan=True)


cases = {
    "fb 3-channel input": lambda m: m.calcOpticalFlowFarneback(bgr, bgr, None, **FB),
    "fb float32 input": lambda m: m.calcOpticalFlowFarneback(a.astype(np.float32), b.astype(np.float32), None, **FB),
    "fb size mismatch": lambda m: m.calcOpticalFlowFarneback(a, b[:-1], None, **FB),
    "fb levels 0": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "levels": 0}),
    "fb levels 20": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "levels": 20}),
    "fb winsize 1": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "winsize": 1}),
    "fb winsize 2": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "winsize": 2}),
    "fb winsize 3": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "winsize": 3}),
    "fb iterations 0": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "iterations": 0}),
    "fb poly_n 3": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "poly_n": 3}),
    "fb poly_n 6": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "poly_n": 6}),
    "fb poly_n 9": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "poly_n": 9, "poly_sigma": 1.7}),
    "fb pyr_scale 0.95": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "pyr_scale": 0.95}),
    "fb pyr_scale 0.3": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "pyr_scale": 0.3}),
    "fb poly_sigma 0": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "poly_sigma": 0.0}),
    "fb tiny 20x20": lambda m: m.calcOpticalFlowFarneback(a[:20, :20].copy(), b[:20, :20].copy(), None, **FB),
    "fb positional": lambda m: m.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0),
    "fb init flow wrong shape": lambda m: m.calcOpticalFlowFarneback(a, b, np.zeros((10, 10, 2), np.float32), **{**FB, "flags": 4}),
    "fb init flow None with flag": lambda m: m.calcOpticalFlowFarneback(a, b, None, **{**FB, "flags": 4}),
    "lk pts (N,1,2)": lambda m: m.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None),
    "lk pts (1,N,2)": lambda m: m.calcOpticalFlowPyrLK(a, b, pts.reshape(1, -1, 2), None),
    "lk pts float64": lambda m: m.calcOpticalFlowPyrLK(a, b, pts.astype(np.float64), None),
    "lk pts int32": lambda m: m.calcOpticalFlowPyrLK(a, b, pts.astype(np.int32), None),
    "lk pts (N,3)": lambda m: m.calcOpticalFlowPyrLK(a, b, np.zeros((4, 3), np.float32), None),
    "lk empty pts": lambda m: m.calcOpticalFlowPyrLK(a, b, np.zeros((0, 2), np.float32), None),
    "lk maxLevel -1": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, maxLevel=-1),
    "lk maxLevel 10": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, maxLevel=10),
    "lk win (2,2)": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(2, 2)),
    "lk win (3,3)": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(3, 3)),
    "lk win larger than image": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(201, 201)),
    "lk criteria count 0": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, criteria=(3, 0, 0.03)),
    "lk criteria type 0": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, criteria=(0, 10, 0.03)),
    "lk size mismatch": lambda m: m.calcOpticalFlowPyrLK(a, b[:-2], pts, None),
    "lk 3-channel": lambda m: m.calcOpticalFlowPyrLK(bgr, bgr, pts, None),
    "lk init flow": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, pts + 1, flags=4),
    "lk init flow missing": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, flags=4),
    "lk minEig 0": lambda m: m.calcOpticalFlowPyrLK(a, b, pts, None, minEigThreshold=0.0),
    "gftt q 0": lambda m: m.goodFeaturesToTrack(a, 10, 0.0, 5),
    "gftt q 1.5": lambda m: m.goodFeaturesToTrack(a, 10, 1.5, 5),
    "gftt minDistance -1": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, -1),
    "gftt maxCorners -5": lambda m: m.goodFeaturesToTrack(a, -5, 0.1, 5),
    "gftt blockSize 4": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, 5, blockSize=4),
    "gftt blockSize 1": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, 5, blockSize=1),
    "gftt blockSize 0": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, 5, blockSize=0),
    "gftt mask wrong size": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, 5, mask=np.ones((10, 10), np.uint8)),
    "gftt mask float": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, 5, mask=np.ones(a.shape, np.float32)),
    "gftt 3-channel": lambda m: m.goodFeaturesToTrack(bgr, 10, 0.1, 5),
    "gftt float32 image": lambda m: m.goodFeaturesToTrack(a.astype(np.float32), 10, 0.1, 5),
    "gftt gradientSize 1": lambda m: m.goodFeaturesToTrack(a, 10, 0.1, 5, gradientSize=1),
    "gftt tiny 3x3": lambda m: m.goodFeaturesToTrack(a[:3, :3].copy(), 10, 0.1, 1),
    "cvt bgra": lambda m: m.cvtColor(np.dstack([bgr, bgr[..., :1]]), m.COLOR_BGR2GRAY),
    "cvt gray input": lambda m: m.cvtColor(a, m.COLOR_BGR2GRAY),
    "cvt float32": lambda m: m.cvtColor(bgr.astype(np.float32), m.COLOR_BGR2GRAY),
    "cvt uint16": lambda m: m.cvtColor(bgr.astype(np.uint16), m.COLOR_BGR2GRAY),
    "cvt empty": lambda m: m.cvtColor(np.zeros((0, 0, 3), np.uint8), m.COLOR_BGR2GRAY),
    "cvt rgb2gray": lambda m: m.cvtColor(bgr, cv2.COLOR_RGB2GRAY),
    "pyrDown odd": lambda m: m.pyrDown(a[:119, :159].copy()),
    "pyrDown 1x1": lambda m: m.pyrDown(a[:1, :1].copy()),
    "pyrDown 3-channel": lambda m: m.pyrDown(bgr),
    "pyrDown dstsize": lambda m: m.pyrDown(a, dstsize=(80, 60)),
}
n_diff = 0
for name, f in cases.items():
    if not hasattr(b2, "pyrDown") and name.startswith("pyrDown"):
        continue
    sc, rc = run(cv2, f)
    sb, rb = run(b2, f)
    if sc != sb:
        n_diff += 1
        print("DIFF  %-28s cv2 %s (%s) | b200 %s (%s)" % (name, sc, rc if sc == "error" else "", sb, rb if sb == "error" else ""), flush=True)
    elif sc == "ok" and not close(rc, rb):
        n_diff += 1
        def d(r):
            return [None if x is None else (np.asarray(x).shape, str(np.asarray(x).dtype)) for x in (r if isinstance(r, tuple) else (r,))]
        def mx(x, y):
            try:
                return float(np.abs(np.asarray(x, np.float64) - np.asarray(y, np.float64)).max())
            except Exception:
                return None
        md = [mx(x, y) for x, y in zip(rc if isinstance(rc, tuple) else (rc,), rb if isinstance(rb, tuple) else (rb,)) if x is not None and y is not None]
        print("VALUE %-28s cv2 %s | b200 %s | max abs differences %s" % (name, d(rc), d(rb), md), flush=True)
    else:
        print("same  %-28s %s" % (name, sc), flush=True)
print("differences", n_diff, "of", len(cases))
