"""One grid point of the real-footage sweep (clip 0, (1185, 1065)) on the CUDA path and the numpy oracle: single window
passes on the level-2 images (maxLevel 0, one iteration, caller-supplied start positions)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2
from oracle import pyrlk as olk, gray_pyr as ogp
from bench import decode_png
z = np.load(os.path.join(ROOT, "tests/golden/real_sweep.npz"))
g0, g1 = decode_png(z["png0_0"]), decode_png(z["png1_0"])
P = np.float32([[1185, 1065]])
for cnt in (5, 6, 7, 8, 9, 10):
    c = (3, cnt, 0.03)
    got = b2.calcOpticalFlowPyrLK(g1, g0, P, None, winSize=(45, 45), maxLevel=2, criteria=c)
    want = olk.pyrlk(g1, g0, P, None, (45, 45), 2, c)
    print("count", cnt, "gpu", got[0].ravel(), "oracle", want[0].ravel())
a2, b2_ = ogp.pyrdown_u8(ogp.pyrdown_u8(g1)), ogp.pyrdown_u8(ogp.pyrdown_u8(g0))
rng = np.random.default_rng(0)
n = 400
prev = np.tile(np.float32([[296.25, 266.25]]), (n, 1))
start = np.float32(np.stack([rng.uniform(255, 300, n), rng.uniform(255, 275, n)], 1))
c = (3, 1, 0.03)
got = b2.calcOpticalFlowPyrLK(a2, b2_, prev, start.copy(), winSize=(45, 45), maxLevel=0, criteria=c, flags=4)
want = olk.pyrlk(a2, b2_, prev, start.copy(), (45, 45), 0, c, flags=4)
d = np.abs(got[0] - want[0]).max(-1)
bad = np.where(d > 1e-3)[0]
print("single passes: bad", len(bad), "of", n)
for k in bad[:40]:
    q = start[k] - 22
    print("  start", start[k], "ix,iy", np.floor(q).astype(int), "gpu", got[0][k], "oracle", want[0][k])
ok = np.where(d <= 1e-3)[0]
iy_ok = np.floor(start[ok, 1] - 22).astype(int); iy_bad = np.floor(start[bad, 1] - 22).astype(int)
print("iy ok range", iy_ok.min(), iy_ok.max(), "iy bad", sorted(set(iy_bad.tolist())))
ix_bad = np.floor(start[bad, 0] - 22).astype(int)
print("ix bad", sorted(set(ix_bad.tolist())))
