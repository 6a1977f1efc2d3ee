"""GFTT (SparseOF.py parameters) timing split on the GPU box: 16 frames at 1080p and 720p."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hackathonopticalflow_b200 import batch, synth, _lib
for (h, w) in [(1080, 1920), (720, 1280)]:
    fr = torch.from_numpy(synth.sequence(h, w, 16, seed=1002)).cuda()
    for kw in [dict(), dict(maxCorners=500, qualityLevel=0.01)]:
        for _ in range(3): batch.gftt(fr, **kw)
        torch.cuda.synchronize()
        _lib.profile(True, reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): c, n = batch.gftt(fr, **kw)
        e1.record(); torch.cuda.synchronize()
        print(h, w, kw, f"{e0.elapsed_time(e1)/10:.3f} ms per 16 frames", {k: round(v['ms'] / 10, 3) for k, v in _lib.profile().items()}, int(n[0]))
        _lib.profile(False, reset=True)

# the reference's footage: the committed real 1080p frame, 16 copies (candidate counts of real content)
import numpy as np
from io import BytesIO
from PIL import Image
f = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "real_1080p.npz"))
g = np.array(Image.open(BytesIO(f["png0"].tobytes())))
fr = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(g, (16,) + g.shape))).cuda()
for kw in [dict(), dict(maxCorners=500, qualityLevel=0.01)]:
    for _ in range(3): batch.gftt(fr, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c, n = batch.gftt(fr, **kw)
    e1.record(); torch.cuda.synchronize()
    print("real 1080p", kw, f"{e0.elapsed_time(e1)/10:.3f} ms per 16 frames", int(n[0]))
