"""Per-call latency of the cv2 drop-in calls (host numpy in, host numpy out), steady state, vs live cv2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from hackathonopticalflow_b200 import cv2compat as b2, pathfinder, synth
def bench(fn, n=20, warm=3):
    for _ in range(warm): fn()
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n * 1e3
for (h, w) in [(1080, 1920), (720, 1280)]:
    bgr = synth.sequence(h, w, 2, seed=3, gray=False)
    g0, g1 = cv2.cvtColor(bgr[0], cv2.COLOR_BGR2GRAY), cv2.cvtColor(bgr[1], cv2.COLOR_BGR2GRAY)
    pts = pathfinder.grid_points(w, h, 30)
    buf = np.empty((h, w, 2), np.float32)
    rows = [
        ("cvtColor", lambda: b2.cvtColor(bgr[0], b2.COLOR_BGR2GRAY), lambda: cv2.cvtColor(bgr[0], cv2.COLOR_BGR2GRAY)),
        ("Farneback (flow=None)", lambda: b2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0), lambda: cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)),
        ("Farneback (flow=buf)", lambda: b2.calcOpticalFlowFarneback(g0, g1, buf, 0.5, 3, 15, 3, 5, 1.2, 0), None),
        ("PyrLK grid 45x45", lambda: b2.calcOpticalFlowPyrLK(g1, g0, pts, None, winSize=(45, 45), maxLevel=2, criteria=(3, 10, 0.03)), lambda: cv2.calcOpticalFlowPyrLK(g1, g0, pts, None, winSize=(45, 45), maxLevel=2, criteria=(3, 10, 0.03))),
        ("GFTT", lambda: b2.goodFeaturesToTrack(g0, 20, 0.3, 10, blockSize=7), lambda: cv2.goodFeaturesToTrack(g0, 20, 0.3, 10, blockSize=7)),
        ("get_flow_lk (LK+filter)", lambda: pathfinder.get_flow_lk(g0, g1, pts), None),
    ]
    for name, mine, ref in rows:
        a = bench(mine)
        c = bench(ref, n=3, warm=1) if ref else float("nan")
        print(f"{h}x{w} {name:28s} b200 {a:8.3f} ms   cv2 {c:8.3f} ms   x{c / a:6.1f}")
