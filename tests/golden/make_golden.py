"""Generates the committed golden fixtures in this directory.

Run HERE (the build container), where /root/reference and the cv2 wheel are both present:
    python tests/golden/make_golden.py
Inputs are real footage from the reference's own sample clips (/root/reference/videos/*.mp4, all 1920x1080);
outputs are what the reference's cv2 calls return on them, called with the reference's call-site arguments
(oracle/cv2_reference.py).  The GPU box has no /root/reference: tests only read the .npz files.

Files written:
  real_crops.npz   4 clips x (BGR crop pair 360x640) + cv2 gray, pyrDown chain, Farneback flow (every 4th px),
                   grid LK (viewer form), Shi-Tomasi corners (+mask form), track LK forward/backward
  real_1080p.npz   one full-resolution gray pair (PNG bytes) + cv2 Farneback flow sampled every 8th px,
                   + the viewer's 2304-point grid LK result and GFTT corners
  synth_small.npz  seeded synthetic pair 135x241 (odd size) with cv2 outputs for non-default parameter sets
"""
import glob
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv2_reference as ref  # noqa: E402
from oracle import pathfinder as opf  # noqa: E402
from hackathonopticalflow_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
VIDEOS = sorted(glob.glob("/root/reference/videos/*.mp4"))
CROP = (360, 640)


def read_pair(path, frame_no):
    cap = cv2.VideoCapture(path)
    cap.set(cv2.CAP_PROP_POS_FRAMES, frame_no)
    ok0, f0 = cap.read()
    ok1, f1 = cap.read()
    assert ok0 and ok1, path
    return f0, f1


def main():
    cv2.setNumThreads(1)
    out = {"cv2_version": np.array(cv2.__version__), "clips": np.array([os.path.basename(v) for v in VIDEOS])}
    starts = [40, 100, 60, 150]
    offs = [(360, 640), (300, 200), (500, 1000), (200, 900)]
    for i, (v, s, (oy, ox)) in enumerate(zip(VIDEOS, starts, offs)):
        f0, f1 = read_pair(v, s)
        c0 = np.ascontiguousarray(f0[oy:oy + CROP[0], ox:ox + CROP[1]])
        c1 = np.ascontiguousarray(f1[oy:oy + CROP[0], ox:ox + CROP[1]])
        g0, g1 = ref.gray(c0), ref.gray(c1)
        out[f"bgr0_{i}"], out[f"bgr1_{i}"] = c0, c1
        out[f"gray0_{i}"], out[f"gray1_{i}"] = g0, g1
        p1 = cv2.pyrDown(g0)
        p2 = cv2.pyrDown(p1)
        out[f"pyr1_{i}"], out[f"pyr2_{i}"] = p1, p2
        flow = ref.farneback(g0, g1)
        out[f"flow_s4_{i}"] = flow[::4, ::4].copy()
        pts = opf.grid_points(CROP[1], CROP[0], 30)
        nxt, st, err = ref.lk_grid(g0, g1, pts)
        out[f"lk_next_{i}"], out[f"lk_status_{i}"], out[f"lk_err_{i}"] = nxt, st, err
        corners = ref.features(g0)
        out[f"gftt_{i}"] = corners if corners is not None else np.zeros((0, 1, 2), np.float32)
        if corners is not None:
            mask = ref.track_mask(g1.shape, corners.reshape(-1, 2))
            cm = ref.features(g1, mask=mask)
            out[f"gftt_mask_{i}"] = mask
            out[f"gftt_masked_{i}"] = cm if cm is not None else np.zeros((0, 1, 2), np.float32)
            p1_, p0r, good, st_f, st_b = ref.lk_track(g0, g1, corners)
            out[f"trk_p1_{i}"], out[f"trk_p0r_{i}"], out[f"trk_good_{i}"] = p1_, p0r, good
            out[f"trk_st_f_{i}"], out[f"trk_st_b_{i}"] = st_f, st_b
        dense = ref.features(g0, maxCorners=500, qualityLevel=0.01, minDistance=5, blockSize=3)
        out[f"gftt_dense_{i}"] = dense
    np.savez_compressed(os.path.join(HERE, "real_crops.npz"), **out)

    # full-resolution pair
    f0, f1 = read_pair(VIDEOS[3], 120)
    g0, g1 = ref.gray(f0), ref.gray(f1)
    flow = ref.farneback(g0, g1)
    pts = opf.grid_points(1920, 1080, 30)
    nxt, st, err = ref.lk_grid(g0, g1, pts)
    full = {
        "png0": np.frombuffer(cv2.imencode(".png", g0)[1].tobytes(), np.uint8),
        "png1": np.frombuffer(cv2.imencode(".png", g1)[1].tobytes(), np.uint8),
        "flow_s8": flow[::8, ::8].copy(),
        "lk_next": nxt, "lk_status": st, "lk_err": err,
        "gftt": ref.features(g0),
        "cv2_version": np.array(cv2.__version__),
    }
    np.savez_compressed(os.path.join(HERE, "real_1080p.npz"), **full)

    # synthetic odd-size pair, non-default parameter sets
    fr = synth.sequence(135, 241, 2, seed=7)
    syn = {"f0": fr[0], "f1": fr[1], "cv2_version": np.array(cv2.__version__)}
    for name, args in {"ref": (0.5, 3, 15, 3, 5, 1.2, 0), "gauss": (0.5, 3, 15, 3, 5, 1.2, 256),
                       "p08": (0.8, 5, 13, 2, 7, 1.5, 0), "even": (0.5, 2, 16, 3, 5, 1.1, 0),
                       "sig0": (0.6, 4, 21, 1, 5, 0.0, 0)}.items():
        syn[f"flow_{name}"] = cv2.calcOpticalFlowFarneback(fr[0], fr[1], None, *args)
        syn[f"args_{name}"] = np.array(args, np.float64)
    np.savez_compressed(os.path.join(HERE, "synth_small.npz"), **syn)
    for f in ("real_crops.npz", "real_1080p.npz", "synth_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
