"""Developer check run on the GPU box: PyrLK / GFTT / viewer filter against live cv2 and the oracle."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch
from hackathonopticalflow_b200 import cv2compat as b2, batch, synth, pathfinder as pf, _lib
from oracle import pathfinder as opf, cv2_reference as ref

def cmp_lk(name, r, m):
    st_eq = (r[1] == m[1]).mean()
    ok = (r[1].ravel() == 1) & (m[1].ravel() == 1)
    d = np.abs(r[0].reshape(-1, 2) - m[0].reshape(-1, 2)).max(axis=1)
    e = np.abs(r[2] - m[2]).ravel()
    print(f"{name}: n={len(d)} status match {st_eq*100:.2f}% fails {int((r[1]==0).sum())} | max dpos(all) {d.max():.2e} max dpos(ok) {d[ok].max() if ok.any() else 0:.2e} | max derr(ok) {e[ok].max() if ok.any() else 0:.2e}")

for (h, w) in [(360, 640), (720, 1280), (1080, 1920)]:
    fr = synth.sequence(h, w, 3, seed=1000)
    pts = pf.grid_points(w, h, 30)
    extra = np.float32([[0, 0], [w - 1, h - 1], [-50, 10], [w + 60, h + 40], [3.5, 100.25], [w - 2.5, 7.75]])
    pts_all = np.vstack([pts, extra])
    for win, ml in [((45, 45), 2), ((15, 15), 2), ((21, 21), 3), ((9, 31), 1)]:
        kw = dict(winSize=win, maxLevel=ml, criteria=(3, 10, 0.03))
        r = cv2.calcOpticalFlowPyrLK(fr[1], fr[0], pts_all, None, **kw)
        t = time.time(); m = b2.calcOpticalFlowPyrLK(fr[1], fr[0], pts_all, None, **kw); dt = time.time() - t
        cmp_lk(f"LK {h}x{w} win{win} L{ml} ({dt*1e3:.1f} ms)", r, m)
    # track form, (N,1,2) points, forward-backward
    p0 = cv2.goodFeaturesToTrack(fr[0], mask=None, **ref.FEATURE_PARAMS)
    m0 = b2.goodFeaturesToTrack(fr[0], mask=None, **ref.FEATURE_PARAMS)
    print(f"GFTT {h}x{w} ref params: cv2 {None if p0 is None else p0.shape} mine {None if m0 is None else m0.shape} equal {p0 is not None and m0 is not None and p0.shape == m0.shape and np.array_equal(p0, m0)}")
    if p0 is not None:
        mask = ref.track_mask(fr[0].shape, p0.reshape(-1, 2))
        for kw in [ref.FEATURE_PARAMS, dict(maxCorners=500, qualityLevel=0.01, minDistance=5, blockSize=3), dict(maxCorners=0, qualityLevel=0.05, minDistance=0, blockSize=5), dict(maxCorners=100, qualityLevel=0.02, minDistance=7.5, blockSize=4, useHarrisDetector=True)]:
            for mk in [None, mask]:
                a = cv2.goodFeaturesToTrack(fr[1], mask=mk, **kw)
                t = time.time(); c = b2.goodFeaturesToTrack(fr[1], mask=mk, **kw); dt = time.time() - t
                same = (a is None and c is None) or (a is not None and c is not None and a.shape == c.shape and np.array_equal(a, c))
                nmatch = 0 if (a is None or c is None) else len(set(map(tuple, a.reshape(-1, 2))) & set(map(tuple, c.reshape(-1, 2))))
                print(f"  GFTT {kw.get('maxCorners')}/{kw.get('qualityLevel')}/{kw.get('minDistance')}/{kw.get('blockSize')} mask={mk is not None}: cv2 {None if a is None else len(a)} mine {None if c is None else len(c)} exact {same} common {nmatch} ({dt*1e3:.1f} ms)")
        r = cv2.calcOpticalFlowPyrLK(fr[0], fr[1], p0, None, **ref.LK_TRACK_PARAMS)
        m = b2.calcOpticalFlowPyrLK(fr[0], fr[1], p0, None, **ref.LK_TRACK_PARAMS)
        cmp_lk(f"LK track fwd {h}x{w}", r, m); assert m[0].shape == p0.shape
    # viewer filter
    nxt, st, err = cv2.calcOpticalFlowPyrLK(fr[1], fr[0], pts, None, **ref.LK_GRID_PARAMS)
    flow_o, pts_o, mask_o, mod_o = opf.vector_filter(nxt, pts, w, h)
    _layer, flow_m, pts_m = pf.get_flow_lk(fr[0], fr[1], pts)
    out = batch.pathfinder_filter(torch.from_numpy(pts).cuda(), torch.from_numpy(nxt).cuda()[None], w, h)
    mask_same = (out["mask"][0].cpu().numpy().astype(bool) == mask_o).mean()
    k = int(out["n_kept"][0])
    same_filter = k == len(pts_o) and np.array_equal(out["kept_pts"][0, :k].cpu().numpy(), pts_o) and np.array_equal(out["kept_flow"][0, :k].cpu().numpy(), flow_o)
    v_o = opf.danger_intensity(flow_o, pts_o)
    print(f"filter {h}x{w}: oracle kept {len(pts_o)} mine {k}; mask agreement {mask_same*100:.2f}% ; filter-on-cv2-LK exact {same_filter}; danger V exact {np.array_equal(out['danger_v'][0,:k].cpu().numpy(), v_o)}; end-to-end kept {len(pts_m)} pts-equal {len(pts_m)==len(pts_o) and np.array_equal(pts_m, pts_o)} flow-equal {len(pts_m)==len(pts_o) and (flow_m==flow_o).all(axis=1).mean() if len(pts_m)==len(pts_o) else 'n/a'}")
    print("  stats", out["stats"][0].cpu().numpy(), "median/p99 numpy", np.median(mod_o), np.percentile(mod_o, 99))

# throughput of the batched LK grid at 1080p
h, w = 1080, 1920
fr = torch.from_numpy(synth.sequence(h, w, 17, seed=1002)).cuda()
pts = torch.from_numpy(pf.grid_points(w, h, 30)).cuda()
ws = torch.empty(_lib.lib().b2of_pyrlk_workspace_bytes(h, w, batch._lk_params((45, 45), 2, (3, 10, 0.03), 0, 1e-4), 16), dtype=torch.uint8, device="cuda")
for _ in range(3): batch.pyrlk(fr[1:], fr[:-1], pts, workspace=ws, **batch.LK_GRID_DEFAULTS)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): batch.pyrlk(fr[1:], fr[:-1], pts, workspace=ws, **batch.LK_GRID_DEFAULTS)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"LK grid 45x45 2304 pts: {ms:.3f} ms per 16 pairs -> {16/ms*1e3:.1f} pairs/s")
pipe = pf.PathfinderPipeline(h, w, dense=True)
bgr = torch.from_numpy(synth.sequence(h, w, 9, seed=1003, gray=False)).cuda()
o = pipe.run(bgr); torch.cuda.synchronize()
print("pipeline ok; n_kept", o["n_kept"].cpu().numpy(), "flow stats", o["flow_stats"][0].cpu().numpy())
