import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from hackathonopticalflow_b200 import cv2compat as b2
rng = np.random.default_rng(1)
def smooth(h, w):
    a = rng.random((h // 4 + 3, w // 4 + 3)).astype(np.float32)
    return (cv2.resize(a, (w, h), interpolation=cv2.INTER_CUBIC).clip(0, 1) * 255).astype(np.uint8)
for (h, w) in [(8, 8), (6, 40), (20, 30), (33, 65), (64, 48), (32, 32), (31, 200), (57, 57), (113, 71)]:
    a = smooth(h, w); b = np.roll(a, 1, axis=1)
    for args in [(0.5, 3, 15, 3, 5, 1.2, 0), (0.5, 1, 5, 1, 7, 1.5, 0), (0.7, 4, 9, 2, 5, 1.1, 256)]:
        try:
            ref = cv2.calcOpticalFlowFarneback(a, b, None, *args)
        except cv2.error as e:
            print(h, w, args, "cv2 error", str(e)[:60]); continue
        try:
            got = b2.calcOpticalFlowFarneback(a, b, None, *args)
            d = np.sqrt(((ref - got) ** 2).sum(-1))
            print(h, w, args, f"mean {d.mean():.2e} max {d.max():.2e}")
        except Exception as e:
            print(h, w, args, "ERR", str(e)[:100])
    pts = np.float32([[1, 1], [w / 2, h / 2], [w - 2, h - 2], [0.5, 0.5]])
    for win in [(5, 5), (15, 15), (45, 45)]:
        try:
            r = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=win, maxLevel=2)
            m = b2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=win, maxLevel=2)
            print(h, w, "LK", win, "status eq", (r[1] == m[1]).all(), "dpos", np.abs(r[0] - m[0]).max())
        except Exception as e:
            print(h, w, "LK", win, "ERR", type(e).__name__, str(e)[:80])
    try:
        r = cv2.goodFeaturesToTrack(a, 10, 0.1, 3, blockSize=3); m = b2.goodFeaturesToTrack(a, 10, 0.1, 3, blockSize=3)
        print(h, w, "GFTT", None if r is None else len(r), None if m is None else len(m), (r is None and m is None) or (r is not None and m is not None and r.shape == m.shape and np.array_equal(r, m)))
    except Exception as e:
        print(h, w, "GFTT ERR", str(e)[:80])
