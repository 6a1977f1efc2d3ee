"""Overlay layers and the dense-flow helpers against LIVE cv2 drawing calls made the way the reference makes them
(pathfinder_viewer.py:179-191, :196-223, draw_hsv :124-141) on random inputs: vectors far out of the frame (clipped
lines), negative end points, zero-length vectors, several frame sizes."""
import os, sys
import numpy as np
import torch
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import batch, pathfinder
from oracle import pathfinder as opf
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for case in range(24):
    w, h = [(1920, 1080), (640, 360), (1280, 720), (331, 257), (3840, 2160), (64, 48)][case % 6]
    pts = pathfinder.grid_points(w, h, 30)
    sc = [0.5, 3, 30, 300, 3000][case % 5]
    nxt = pts + np.float32(rng.normal(0, sc, pts.shape))
    if case % 4 == 0:
        nxt[::7] = pts[::7]                                    # zero-length vectors
    out = batch.pathfinder_filter(torch.from_numpy(pts).cuda(), torch.from_numpy(nxt).cuda()[None], w, h, all_points=True)
    ap, an = out["all_pts"][0].cpu().numpy(), out["all_next"][0].cpu().numpy()
    mask = out["mask"][0].cpu().numpy().astype(bool)
    # the reference's drawing, with live cv2
    layer = np.zeros((h, w, 3), np.uint8)
    lines = np.concatenate((ap[mask], an[mask]), axis=1)
    cv2.polylines(layer, lines.reshape(-1, 2, 2), False, (0, 0, 255))
    for x1, y1, _x2, _y2 in lines:
        cv2.circle(layer, center=(int(x1), int(y1)), radius=1, color=(255, 0, 255), thickness=1)
    lines_bad = np.concatenate((ap[~mask], an[~mask]), axis=1)
    cv2.polylines(layer, lines_bad.reshape(-1, 2, 2), False, (255, 255, 0))
    for x1, y1, _x2, _y2 in lines_bad:
        cv2.circle(layer, center=(int(x1), int(y1)), radius=1, color=(255, 255, 0), thickness=1)
    got = batch.overlay_vectors(out, h, w)[0].cpu().numpy()
    nd = int((got != layer).any(-1).sum())
    # lamps
    k = int(out["n_kept"][0])
    kp, kf = out["kept_pts"][0, :k].cpu().numpy(), out["kept_flow"][0, :k].cpu().numpy()
    fx, fy = kf[:, 0], kf[:, 1]
    modulus = np.sqrt(fx * fx + fy * fy)
    hsv = np.zeros((h, w, 3), np.uint8)
    for (x, y), m in zip(kp, modulus):
        hsv[y, x, 0] = 0; hsv[y, x, 1] = 255; hsv[y, x, 2] = np.minimum(50 + m * 2, 255)
    bgr = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)
    for x, y in kp:
        cv2.circle(bgr, center=(int(x), int(y)), radius=6, color=(int(bgr[y, x, 0]), int(bgr[y, x, 1]), int(bgr[y, x, 2])), thickness=-1)
    gl = batch.overlay_lamps(out, h, w)[0].cpu().numpy()
    nl = int((gl != bgr).any(-1).sum())
    # dense helpers on a random field
    fl = np.float32(rng.normal(0, [0.01, 1, 20, 200][case % 4], (h, w, 2)))
    if case % 3 == 0:
        fl[::5, ::3] = 0
    fh = batch.flow_hsv(torch.from_numpy(fl).cuda()[None])[0].cpu().numpy()
    want_h, _ = opf.draw_hsv(fl)
    nh = int((fh != want_h).any(-1).sum())
    fs = batch.flow_sample(torch.from_numpy(fl).cuda()[None], torch.from_numpy(pts).cuda())[0].cpu().numpy()
    ip = pts.astype(np.int64)
    want_s = pts + fl[ip[:, 1], ip[:, 0]]
    ns = int((fs != want_s).any(-1).sum())
    flag = "" if nd == nl == nh == ns == 0 else "MISMATCH "
    bad += flag != ""
    print("%s%dx%d scale %g kept %d: vector layer differing px %d, lamp layer %d, hsv picture %d, sampled points %d" % (flag, w, h, sc, k, nd, nl, nh, ns), flush=True)
print("mismatching", bad)
