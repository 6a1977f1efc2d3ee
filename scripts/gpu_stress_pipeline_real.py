"""PathfinderPipeline on the four real 1080p pairs played as one BGR sequence (a0 a1 b0 b1 ...), several chunk sizes:
every per-pair output against the single-call path (cv2compat) and the oracle restatement of the reference's filter, and
against live cv2 / the reference-style chain for the proper pairs."""
import os, sys
import numpy as np
import torch
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2, pathfinder, batch
from oracle import pathfinder as opf
from bench import decode_png
z = np.load(os.path.join(ROOT, "tests/golden/real_sweep.npz"))
f = np.load(os.path.join(ROOT, "tests/golden/real_1080p.npz"))
pairs = [(decode_png(z[f"png0_{i}"]), decode_png(z[f"png1_{i}"])) for i in range(3)] + [(decode_png(f["png0"]), decode_png(f["png1"]))]
seq = np.stack([g for p in pairs for g in p])                       # 8 frames, 7 pairs (odd pairs are scene cuts)
bgr = np.ascontiguousarray(np.repeat(seq[..., None], 3, -1))
assert np.array_equal(b2.cvtColor(bgr[0], b2.COLOR_BGR2GRAY), seq[0])
pts = pathfinder.grid_points(1920, 1080, 30)
LK = dict(winSize=(45, 45), maxLevel=2, criteria=(3, 10, 0.03))
FB = (0.5, 3, 15, 3, 5, 1.2, 0)
bad = 0
single = []
for k in range(7):
    nxt, st, err = b2.calcOpticalFlowPyrLK(seq[k + 1], seq[k], pts, None, **LK)
    flow = b2.calcOpticalFlowFarneback(seq[k], seq[k + 1], None, *FB)
    single.append((nxt, st, err, flow))
for chunk in (1, 2, 3, 4, 7, 8):
    pipe = pathfinder.PathfinderPipeline(1080, 1920, dense=True, chunk_pairs=chunk)
    out = pipe.run(torch.from_numpy(bgr).cuda())
    torch.cuda.synchronize()
    for k in range(7):
        nxt, st, err, flow = single[k]
        ok = [np.array_equal(out["gray"][k].cpu().numpy(), seq[k]),
              np.array_equal(out["next_pts"][k].cpu().numpy(), nxt.reshape(-1, 2)),
              np.array_equal(out["status"][k].cpu().numpy().ravel(), st.ravel()),
              np.array_equal(out["err"][k].cpu().numpy().ravel(), err.ravel()),
              np.array_equal(out["flow"][k].cpu().numpy(), flow)]
        fo, po, mo, _ = opf.vector_filter(nxt.reshape(-1, 2), pts, 1920, 1080)
        n = int(out["n_kept"][k])
        ok += [n == len(po), np.array_equal(out["kept_pts"][k, :n].cpu().numpy(), po),
               np.array_equal(out["kept_flow"][k, :n].cpu().numpy(), fo),
               np.array_equal(out["danger_v"][k, :n].cpu().numpy(), opf.danger_intensity(fo, po))]
        mag = np.sqrt((flow.astype(np.float64) ** 2).sum(-1))
        s = out["flow_stats"][k].cpu().numpy()
        ok += [abs(s[0] - mag.mean()) <= 1e-4 * max(1, mag.mean()), abs(s[1] - mag.max()) <= 1e-4 * max(1, mag.max())]
        if not all(ok):
            bad += 1
            print("MISMATCH chunk", chunk, "pair", k, ok, s[:4], mag.mean(), mag.max(), flush=True)
    print("chunk", chunk, "done", flush=True)
# the proper pairs against live cv2 through the reference-style chain
for c in range(4):
    k = 2 * c
    wn, ws, we = cv2.calcOpticalFlowPyrLK(seq[k + 1], seq[k], pts, None, **LK)
    fo, po, mo, _ = opf.vector_filter(wn.reshape(-1, 2), pts, 1920, 1080)
    n = int(out["n_kept"][k])
    mask = out["mask"][k].cpu().numpy().astype(bool)
    print("clip", c, "filter mask agreement with the cv2 chain %.5f, kept %d vs %d" % ((mask == mo).mean(), n, len(po)))
print("mismatching", bad)
