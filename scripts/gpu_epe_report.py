"""EPE of the CUDA Farneback against the committed cv2 goldens (real footage) -- the numbers DESIGN.md quotes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from io import BytesIO
from PIL import Image
from hackathonopticalflow_b200 import cv2compat as b2
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
def epe(a, b):
    d = np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum(-1))
    return d.mean(), d.max()
c = np.load(os.path.join(G, "real_crops.npz")); f = np.load(os.path.join(G, "real_1080p.npz"))
for i in range(4):
    fl = b2.calcOpticalFlowFarneback(c[f"gray0_{i}"], c[f"gray1_{i}"], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    print("crop", i, "mean %.3e max %.3e" % epe(fl[::4, ::4], c[f"flow_s4_{i}"]), "flow max %.1f" % np.abs(c[f"flow_s4_{i}"]).max())
dec = lambda b: np.array(Image.open(BytesIO(b.tobytes())))
fl = b2.calcOpticalFlowFarneback(dec(f["png0"]), dec(f["png1"]), None, 0.5, 3, 15, 3, 5, 1.2, 0)
print("full 1080p mean %.3e max %.3e" % epe(fl[::8, ::8], f["flow_s8"]))
