"""Dumps the CUDA path's Farneback flow (every 8th pixel) on the committed full-resolution real pairs to
gpurun_out/sweep_flow.npz, for the conditioning analysis of tests/golden/make_golden_sweep.py."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2
from bench import decode_png
z = np.load(os.path.join(ROOT, "tests/golden/real_sweep.npz"))
f = np.load(os.path.join(ROOT, "tests/golden/real_1080p.npz"))
out = {}
pairs = [(decode_png(z[f"png0_{i}"]), decode_png(z[f"png1_{i}"])) for i in range(3)] + [(decode_png(f["png0"]), decode_png(f["png1"]))]
for i, (g0, g1) in enumerate(pairs):
    fl = b2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    out[f"flow_{i}"] = fl[::8, ::8].copy()
    out[f"flow_full_{i}"] = fl.astype(np.float16) if i == 2 else np.zeros(1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out/sweep_flow.npz"), **out)
print("ok")
