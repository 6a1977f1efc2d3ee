"""Generates reference_funcs.npz and extras.npz (run HERE, where /root/reference and the cv2 wheel are present):

    python tests/golden/make_golden_ref.py

reference_funcs.npz -- outputs of the REFERENCE'S OWN FUNCTIONS, not of a restatement.  The reference's scripts cannot
be imported (module-level GUI loops), so the function definitions are cut out of the source files with
``ast`` and executed unmodified in a namespace that holds what their module would have held (cv2, numpy, logging and
the module globals they read: width, height, half_width, half_height, draw_bad_flow):
  * pathfinder_viewer.py  get_flow_lk (:144-193), draw_sparse_lamps (:196-223), draw_hsv (:124-141)
  * DenseOF.py            get_flow_lk (:160-266, the m > 1.2 * median rule at :228)
on the committed real 1080p pair (tests/golden/real_1080p.npz) and the four real crops (real_crops.npz).

extras.npz -- cv2 outputs for the options the reference leaves at their defaults: Harris response,
gradientSize 5 / 7, OPTFLOW_LK_GET_MIN_EIGENVALS, Farneback with iterations = 0 (with and without an initial flow).
"""
import ast
import glob
import logging
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
VIEWER = "/root/reference/pathfinder_viewer.py"
DENSEOF = glob.glob("/root/reference/*/DenseOF.py")[0]


def extract(path, names, extra_globals):
    """exec the named top-level function definitions of `path`, unmodified, in a fresh namespace."""
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in picked} == set(names), (path, names)
    ns = {"np": np, "cv2": cv2, "logging": logging, "__name__": "reference"}
    ns.update(extra_globals)
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), ns)
    return ns


def run_viewer(prev_gray, gray, pts):
    h, w = gray.shape
    ns = extract(VIEWER, ["get_flow_lk", "draw_sparse_lamps", "draw_hsv"],
                 dict(width=w, height=h, half_width=int(w / 2), half_height=int(h / 2), draw_bad_flow=True))
    layer, flow, kept = ns["get_flow_lk"](prev_gray, gray, pts.copy())
    lamps = ns["draw_sparse_lamps"](flow, kept)
    return ns, layer, flow, kept, lamps


def main():
    cv2.setNumThreads(1)
    from oracle import pathfinder as opf
    full = np.load(os.path.join(HERE, "real_1080p.npz"))
    crops = np.load(os.path.join(HERE, "real_crops.npz"))
    g0 = cv2.imdecode(full["png0"], cv2.IMREAD_GRAYSCALE)
    g1 = cv2.imdecode(full["png1"], cv2.IMREAD_GRAYSCALE)
    out = {"cv2_version": np.array(cv2.__version__)}

    cases = [("full", g0, g1, 1920, 1080)] + [(f"crop{i}", crops[f"gray0_{i}"], crops[f"gray1_{i}"], 640, 360)
                                              for i in range(4)]
    for name, a, b, w, h in cases:
        pts = opf.grid_points(w, h, 30)
        ns, layer, flow, kept, lamps = run_viewer(a, b, pts)
        out[f"{name}_kept_flow"] = flow.astype(np.int32)
        out[f"{name}_kept_pts"] = kept.astype(np.int32)
        out[f"{name}_danger_v"] = lamps[kept[:, 1], kept[:, 0], 2].copy()      # V of the lamp at its own centre
        # the drawn layers, losslessly compressed (a few KB each: mostly black)
        out[f"{name}_layer_png"] = np.frombuffer(cv2.imencode(".png", layer)[1].tobytes(), np.uint8)
        out[f"{name}_lamps_png"] = np.frombuffer(cv2.imencode(".png", lamps)[1].tobytes(), np.uint8)
        # DenseOF.py's variant builds its own grid from `step`
        nd = extract(DENSEOF, ["get_flow_lk"], dict(width=w, height=h, half_width=int(w / 2), half_height=int(h / 2)))
        _layer, dflow, dkept = nd["get_flow_lk"](a, b, step=30)
        out[f"{name}_denseof_kept_flow"] = dflow.astype(np.int32)
        out[f"{name}_denseof_kept_pts"] = dkept.astype(np.int32)
    # draw_hsv on a real dense field (every 8th pixel of the committed golden flow: what the GPU test can rebuild)
    hsv_in = np.ascontiguousarray(full["flow_s8"])
    out["hsv_of_flow_s8"] = ns["draw_hsv"](hsv_in)
    np.savez_compressed(os.path.join(HERE, "reference_funcs.npz"), **out)

    ex = {"cv2_version": np.array(cv2.__version__)}
    for i in range(4):
        a, b = crops[f"gray0_{i}"], crops[f"gray1_{i}"]
        empty = np.zeros((0, 1, 2), np.float32)
        for tag, kw in {"harris": dict(useHarrisDetector=True, k=0.04),
                        "harris_dense": dict(maxCorners=500, qualityLevel=0.01, minDistance=5, blockSize=3,
                                             useHarrisDetector=True, k=0.06),
                        "grad5": dict(gradientSize=5), "grad7": dict(gradientSize=7),
                        "grad5_harris": dict(gradientSize=5, useHarrisDetector=True, k=0.04)}.items():
            p = dict(maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7)
            p.update(kw)
            c = cv2.goodFeaturesToTrack(a, mask=None, **p)
            ex[f"gftt_{tag}_{i}"] = c if c is not None else empty
        pts = opf.grid_points(640, 360, 30)
        nxt, st, err = cv2.calcOpticalFlowPyrLK(b, a, pts, None, winSize=(45, 45), maxLevel=2,
                                                criteria=(3, 10, 0.03), flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS)
        ex[f"lk_mineig_next_{i}"], ex[f"lk_mineig_status_{i}"], ex[f"lk_mineig_err_{i}"] = nxt, st, err
        nxt, st, err = cv2.calcOpticalFlowPyrLK(b, a, pts, None, winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03),
                                                flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS, minEigThreshold=1e-3)
        ex[f"lk_mineig15_next_{i}"], ex[f"lk_mineig15_status_{i}"], ex[f"lk_mineig15_err_{i}"] = nxt, st, err
    syn = np.load(os.path.join(HERE, "synth_small.npz"))
    f0, f1 = syn["f0"], syn["f1"]
    ex["fb_iter0"] = cv2.calcOpticalFlowFarneback(f0, f1, None, 0.5, 3, 15, 0, 5, 1.2, 0)
    init = syn["flow_ref"].copy()
    ex["fb_iter0_init_in"] = init.copy()
    ex["fb_iter0_init"] = cv2.calcOpticalFlowFarneback(f0, f1, init.copy(), 0.5, 3, 15, 0, 5, 1.2,
                                                       cv2.OPTFLOW_USE_INITIAL_FLOW)
    np.savez_compressed(os.path.join(HERE, "extras.npz"), **ex)
    for f in ("reference_funcs.npz", "extras.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
