"""LK grid (viewer form) timing split on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hackathonopticalflow_b200 import batch, synth, pathfinder as pf, _lib
h, w = 1080, 1920
B = 16
fr = torch.from_numpy(synth.sequence(h, w, B + 1, seed=1002)).cuda()
pts = torch.from_numpy(pf.grid_points(w, h, 30)).cuda()
ws = torch.empty(_lib.lib().b2of_pyrlk_workspace_bytes(h, w, batch._lk_params((45, 45), 2, (3, 10, 0.03), 0, 1e-4), B), dtype=torch.uint8, device="cuda")
cur, prev = fr[1:].contiguous(), fr[:-1].contiguous()
for _ in range(3): batch.pyrlk(cur, prev, pts, workspace=ws, **batch.LK_GRID_DEFAULTS)
torch.cuda.synchronize()
_lib.profile(True, reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): batch.pyrlk(cur, prev, pts, workspace=ws, **batch.LK_GRID_DEFAULTS)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"LK grid 45x45 2304 pts: {ms:.3f} ms per {B} pairs -> {B/ms*1e3:.1f} pairs/s")
print({k: round(v['ms'] / 10, 3) for k, v in _lib.profile().items()})
_lib.profile(False, reset=True)
for win in [(15, 15), (21, 21)]:
    kw = dict(winSize=win, maxLevel=2, criteria=(3, 10, 0.03))
    for _ in range(2): batch.pyrlk(cur, prev, pts, workspace=ws, **kw)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): batch.pyrlk(cur, prev, pts, workspace=ws, **kw)
    e1.record(); torch.cuda.synchronize()
    print(win, f"{e0.elapsed_time(e1)/10:.3f} ms per {B} pairs")
