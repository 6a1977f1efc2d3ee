"""Vector filter + danger intensity (K12) against the oracle restatement of the reference's functions on inputs the
fixtures do not hold: the sweep's grid LK results, random flows of several scales, tiny and large point counts, ties,
zero flow, far outliers; both mask rules."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import batch, pathfinder
from oracle import pathfinder as opf
z = np.load(os.path.join(ROOT, "tests/golden/real_sweep.npz"))
rng = np.random.default_rng(5)
cases = []
g1080 = pathfinder.grid_points(1920, 1080, 30)
for i in range(3):
    cases.append((f"sweep{i}", g1080, z[f"lk_next_{i}"].reshape(-1, 2), 1920, 1080))
g4k = pathfinder.grid_points(3840, 2160, 30)
for sc in (0.01, 0.5, 3.0, 40.0, 1000.0):
    cases.append((f"rand1080 s={sc}", g1080, g1080 + np.float32(rng.normal(0, sc, g1080.shape)), 1920, 1080))
    cases.append((f"rand4k s={sc}", g4k, g4k + np.float32(rng.normal(0, sc, g4k.shape)), 3840, 2160))
for n in (1, 2, 3, 4, 5, 7, 33, 100, 101):
    p = np.float32(np.stack([rng.uniform(0, 640, n), rng.uniform(0, 360, n)], 1))
    cases.append((f"n={n}", p, p + np.float32(rng.normal(0, 2, p.shape)), 640, 360))
cases.append(("zero flow", g1080, g1080.copy(), 1920, 1080))
q = g1080 + np.float32(np.round(rng.normal(0, 2, g1080.shape)))           # many exact ties
cases.append(("ties", g1080, q, 1920, 1080))
q = g1080 + np.float32(rng.normal(0, 1, g1080.shape)); q[::97] += 1e4
cases.append(("outliers", g1080, q, 1920, 1080))
q = g1080.copy(); q[:1200] += 3.0
cases.append(("half constant", g1080, q, 1920, 1080))
bad = 0
for name, pts, nxt, w, h in cases:
    for mode, rule in ((batch.FILTER_VIEWER, "viewer"), (batch.FILTER_DENSEOF, "denseof")):
        flow_o, pts_o, mask_o, mod_o = opf.vector_filter(nxt, pts, w, h, rule=rule)
        out = batch.pathfinder_filter(torch.from_numpy(np.ascontiguousarray(pts)).cuda(), torch.from_numpy(np.ascontiguousarray(nxt)).cuda()[None], w, h, mode=mode)
        mask = out["mask"][0].cpu().numpy().astype(bool)
        k = int(out["n_kept"][0])
        ok_mask = np.array_equal(mask, mask_o)
        ok = ok_mask and k == len(pts_o) and np.array_equal(out["kept_pts"][0, :k].cpu().numpy(), pts_o)
        flow_same = ok and np.array_equal(out["kept_flow"][0, :k].cpu().numpy(), flow_o)
        v_same = ok and np.array_equal(out["danger_v"][0, :k].cpu().numpy(), opf.danger_intensity(flow_o, pts_o))
        if not (ok and flow_same and v_same):
            bad += 1
            print("MISMATCH", name, rule, "mask agree %.5f" % (mask == mask_o).mean(), "k", k, len(pts_o), "flow", flow_same, "v", v_same,
                  "flow diff count", int((out["kept_flow"][0, :k].cpu().numpy() != flow_o).any(1).sum()) if ok else -1, flush=True)
print("cases", 2 * len(cases), "mismatching", bad)
