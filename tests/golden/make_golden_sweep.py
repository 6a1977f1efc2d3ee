"""Real-footage sweep at the clips' native 1920x1080 (BASELINE configs[0]): one full-resolution pair from each of the
reference's sample clips that real_1080p.npz does not already cover.

Run HERE (the build container: /root/reference and the cv2 wheel are present; the GPU box has neither):
    python tests/golden/make_golden_sweep.py
Writes real_sweep.npz: per clip i the gray pair (PNG bytes), cv2's Farneback flow with the reference's arguments
sampled every 8th pixel, the viewer's 2304-point grid LK (current -> previous), the Shi-Tomasi corners with
SparseOF.py's parameters, and the forward / backward 15x15 track of those corners.

Conditioning of the dense flow.  On this footage cv2's flow is not reproducible at every pixel: where the 2x2 system
of FarnebackUpdateFlow is near-singular (flat walls, dark corridor) rounding noise is amplified through
det = g11*g22 - g12^2 + 1e-3 and fed back through the warp of the next iteration.  Measured here: cv2 with and without
its SIMD paths (cv2.setUseOptimized) differs from itself by up to 0.40 / 0.05 / 2.4 px on the three pairs, and one
grey level added to 0.1 % of the pixels of the first frame moves cv2's own flow by up to 35 / 14 / 55 px.  So the
file also holds, per pair (clip 3 = real_1080p.npz included), at the same every-8th-pixel sampling:
  flow_s8_noopt_i   cv2's flow with setUseOptimized(False)
  stable_i          packed bits: 1 where, within +-8 px, cv2's flow moves by less than 0.05 px under each of three
                    such one-grey-level perturbations (seeds 0..2) AND between its optimised / plain builds
The parity tests hold the north_star's max bound on the stable pixels (90-95 % of a frame) and bound the count of
outliers elsewhere; the mean bound is held over all pixels.
"""
import os
import sys

import cv2
import numpy as np
from scipy import ndimage

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv2_reference as ref  # noqa: E402
from oracle import pathfinder as opf  # noqa: E402
from make_golden import VIDEOS, read_pair  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = [90, 60, 110]          # clips 0..2 (clip 3, frames 120-121, is real_1080p.npz)


FB_ARGS = (0.5, 3, 15, 3, 5, 1.2, 0)
STABLE_PX, NOISE_FRAC, NOISE_SEEDS, DILATE = 0.05, 0.001, 3, 17


def epe_map(a, b):
    return np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum(-1))


def conditioning(g0, g1, flow):
    """(cv2 flow with its SIMD paths off sampled every 8th px, stable mask at the same sampling, stats)."""
    cv2.setUseOptimized(False)
    plain = cv2.calcOpticalFlowFarneback(g0, g1, None, *FB_ARGS)
    cv2.setUseOptimized(True)
    move = epe_map(flow, plain)
    self_max = float(move.max())
    noise_max = 0.0
    for seed in range(NOISE_SEEDS):
        rng = np.random.default_rng(seed)
        m = rng.random(g0.shape) < NOISE_FRAC
        gn = g0.copy()
        gn[m] = np.clip(gn[m].astype(np.int32) + 1, 0, 255).astype(np.uint8)
        d = epe_map(flow, cv2.calcOpticalFlowFarneback(gn, g1, None, *FB_ARGS))
        noise_max = max(noise_max, float(d.max()))
        move = np.maximum(move, d)
    unstable = ndimage.maximum_filter(move > STABLE_PX, size=DILATE)[::8, ::8]
    return plain[::8, ::8].copy(), ~unstable, self_max, noise_max


def main():
    cv2.setNumThreads(1)
    out = {"cv2_version": np.array(cv2.__version__), "clips": np.array([os.path.basename(v) for v in VIDEOS[:3]]),
           "frame_no": np.array(FRAMES)}
    pts = opf.grid_points(1920, 1080, 30)
    for i, (v, s) in enumerate(zip(VIDEOS[:3], FRAMES)):
        f0, f1 = read_pair(v, s)
        g0, g1 = ref.gray(f0), ref.gray(f1)
        assert g0.shape == (1080, 1920)
        out[f"png0_{i}"] = np.frombuffer(cv2.imencode(".png", g0)[1].tobytes(), np.uint8)
        out[f"png1_{i}"] = np.frombuffer(cv2.imencode(".png", g1)[1].tobytes(), np.uint8)
        flow = ref.farneback(g0, g1)
        out[f"flow_s8_{i}"] = flow[::8, ::8].copy()
        plain, stable, self_max, noise_max = conditioning(g0, g1, flow)
        out[f"flow_s8_noopt_{i}"], out[f"stable_{i}"] = plain, np.packbits(stable)
        print("   cv2 vs itself (SIMD off) max %.3f px, under 1-level noise max %.1f px, stable %.1f %%"
              % (self_max, noise_max, 100 * stable.mean()))
        nxt, st, err = ref.lk_grid(g0, g1, pts)
        out[f"lk_next_{i}"], out[f"lk_status_{i}"], out[f"lk_err_{i}"] = nxt, st, err
        corners = ref.features(g0)
        out[f"gftt_{i}"] = corners if corners is not None else np.zeros((0, 1, 2), np.float32)
        if corners is not None:
            p1, p0r, good, st_f, st_b = ref.lk_track(g0, g1, corners)
            out[f"trk_p1_{i}"], out[f"trk_p0r_{i}"], out[f"trk_good_{i}"] = p1, p0r, good
            out[f"trk_st_f_{i}"], out[f"trk_st_b_{i}"] = st_f, st_b
        mag = np.sqrt((flow ** 2).sum(-1))
        print(i, os.path.basename(v), "frame", s, "flow mean %.2f max %.1f px" % (mag.mean(), mag.max()),
              "corners", 0 if corners is None else len(corners), "lk ok", int(st.sum()))
    z = np.load(os.path.join(HERE, "real_1080p.npz"))       # clip 3: the pair of real_1080p.npz, mask only
    g0, g1 = cv2.imdecode(z["png0"], cv2.IMREAD_GRAYSCALE), cv2.imdecode(z["png1"], cv2.IMREAD_GRAYSCALE)
    flow = ref.farneback(g0, g1)
    assert np.array_equal(flow[::8, ::8], z["flow_s8"])
    plain, stable, self_max, noise_max = conditioning(g0, g1, flow)
    out["flow_s8_noopt_3"], out["stable_3"] = plain, np.packbits(stable)
    print("3 (real_1080p.npz) cv2 vs itself max %.3f px, under noise max %.1f px, stable %.1f %%"
          % (self_max, noise_max, 100 * stable.mean()))
    path = os.path.join(HERE, "real_sweep.npz")
    np.savez_compressed(path, **out)
    print("real_sweep.npz", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
