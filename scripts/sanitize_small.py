"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once, small sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hackathonopticalflow_b200 import cv2compat as b2, pathfinder, synth
bgr = synth.sequence(135, 241, 3, seed=5, gray=False)
g = [b2.cvtColor(f, b2.COLOR_BGR2GRAY) for f in bgr]
b2.pyrDown(g[0])
for args in [(0.5, 3, 15, 3, 5, 1.2, 0), (0.5, 3, 15, 3, 5, 1.2, 256), (0.8, 5, 13, 2, 7, 1.5, 0), (0.5, 2, 16, 3, 5, 1.1, 0)]:
    f = b2.calcOpticalFlowFarneback(g[0], g[1], None, *args)
    assert np.isfinite(f).all()
b2.calcOpticalFlowFarnebackSequence(np.stack(g))
pts = pathfinder.grid_points(241, 135, 30)
b2.calcOpticalFlowPyrLK(g[1], g[0], pts, None, winSize=(45, 45), maxLevel=2, criteria=(3, 10, 0.03))
b2.calcOpticalFlowPyrLK(g[1], g[0], pts, None, winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))
b2.goodFeaturesToTrack(g[0], 20, 0.3, 10, blockSize=7)
b2.goodFeaturesToTrack(g[0], 0, 0.01, 0, blockSize=3)
pathfinder.get_flow_lk(g[0], g[1], pts)
print("sanitize run ok")
