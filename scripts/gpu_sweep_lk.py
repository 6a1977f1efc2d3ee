"""LK / GFTT / track on the real-footage sweep pairs against the committed cv2 results: per-clip disagreement report."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2, pathfinder
from bench import decode_png
z = np.load(os.path.join(ROOT, "tests/golden/real_sweep.npz"))
pts = pathfinder.grid_points(1920, 1080, 30)
out = {}
for i in range(3):
    g0, g1 = decode_png(z[f"png0_{i}"]), decode_png(z[f"png1_{i}"])
    nxt, st, err = b2.calcOpticalFlowPyrLK(g1, g0, pts, None, winSize=(45, 45), maxLevel=2, criteria=(3, 10, 0.03))
    out[f"nxt_{i}"], out[f"st_{i}"], out[f"err_{i}"] = nxt, st, err
    d = np.abs(nxt.reshape(-1, 2) - z[f"lk_next_{i}"].reshape(-1, 2)).max(-1)
    ws = z[f"lk_status_{i}"].ravel()
    print(i, "status agree", (st.ravel() == ws).mean(), "pos max", d.max(), "n>0.05", (d > 0.05).sum(),
          "of which cv2 status 0:", ((d > 0.05) & (ws == 0)).sum(), "max over cv2-status-1", d[ws == 1].max())
    ok = (st.ravel() == 1) & (ws == 1)
    print("   err max diff on both-ok", np.abs(err.ravel() - z[f"lk_err_{i}"].ravel())[ok].max())
    p0 = b2.goodFeaturesToTrack(g0, mask=None, maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7)
    print("   gftt equal", np.array_equal(p0, z[f"gftt_{i}"]))
    p1, s1, _ = b2.calcOpticalFlowPyrLK(g0, g1, z[f"gftt_{i}"], None, winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))
    print("   track st equal", np.array_equal(s1, z[f"trk_st_f_{i}"]), "pos max", np.abs(p1 - z[f"trk_p1_{i}"]).max())
np.savez_compressed(os.path.join(ROOT, "gpurun_out/sweep_lk.npz"), **out)
