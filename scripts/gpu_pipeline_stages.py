"""Stage-by-stage wall time of PathfinderPipeline.run (host + device, synchronised after every call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hackathonopticalflow_b200 import batch, pathfinder, synth
h, w, P = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1080, 1920, 16)
base = synth.sequence(h, w, 3, seed=7, gray=False)
bgr = torch.from_numpy(np.ascontiguousarray(base[([0, 1, 2, 1] * (P // 4 + 1))[:P + 1]])).cuda()
pipe = pathfinder.PathfinderPipeline(h, w, dense=True, chunk_pairs=P)
def T(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    print(f"{name:28s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms")
    return r
gray = T("bgr2gray", lambda: batch.bgr2gray(bgr))
prev, cur = gray[:-1].contiguous(), gray[1:].contiguous()
nxt, st, err = T("pyrlk", lambda: batch.pyrlk(cur, prev, pipe.points, **batch.LK_GRID_DEFAULTS))
T("filter", lambda: batch.pathfinder_filter(pipe.points, nxt, w, h))
flow = T("flow_sequence", lambda: pipe.dense.flow_sequence(gray))
T("flow_stats", lambda: batch.flow_stats(flow))
T("flow_sample", lambda: batch.flow_sample(flow, pipe.points))
T("run (all)", lambda: pipe.run(bgr))
