"""Host-side mirror of the reference's downstream interface, running on the device.

Same names and argument meaning as the reference's functions, minus the drawing (out of scope):

* ``grid_points``      -- the measurement grid, pathfinder_viewer.py:255-267 (DenseOF.py:166-180)
* ``get_flow_lk``      -- pathfinder_viewer.py:144-193: LK from the current frame back to the previous one
                          (note the argument order at :156), modulus normalisation, median/p99 filter, int rounding
* ``draw_sparse_lamps``-- pathfinder_viewer.py:196-223: the lamp layer (filled discs of radius 6 whose brightness is
                          the danger intensity), drawn on the device; ``lamp_intensity`` is the V value alone
* ``draw_hsv``         -- pathfinder_viewer.py:124-141: dense flow -> BGR picture (hue = direction, value = length)
"""
import numpy as np
import torch

from . import batch


def grid_points(width, height, step=30):
    """float32 (N,2), x-major: all y for the first x, then the next x (pathfinder_viewer.py:255-267)."""
    def axis(dim):
        indent = dim % step / 2 if dim // step % 2 == 1 else (dim % step + step) / 2
        return np.arange(indent, dim, step).astype(int)
    xs, ys = axis(width), axis(height)
    pts = np.empty((len(xs), len(ys), 2), np.float32)
    pts[..., 0] = xs[:, None]
    pts[..., 1] = ys[None, :]
    return pts.reshape(-1, 2)


draw_bad_flow = True          # the reference's module-level switch (pathfinder_viewer.py:14)


def get_flow_lk(img1, img2, points_, device="cuda", rule=batch.FILTER_VIEWER):
    """img1: previous gray, img2: current gray, points_: float32 (N,2) (all numpy, as the reference passes them).

    Returns the reference's 3-tuple ``(frame_layer, flow, points_)`` (pathfinder_viewer.py:144-193, unpacked at
    :285): frame_layer uint8 (H,W,3), flow int32 (M,2), points int32 (M,2).  ``rule=batch.FILTER_DENSEOF`` applies the
    development script's mask (DenseOF.py:228) instead of the viewer's.
    """
    height, width = img1.shape
    prev = torch.from_numpy(np.ascontiguousarray(img1)).to(device)[None]
    cur = torch.from_numpy(np.ascontiguousarray(img2)).to(device)[None]
    pts = torch.from_numpy(np.ascontiguousarray(points_, dtype=np.float32)).to(device)
    nxt, _status, _err = batch.pyrlk(cur, prev, pts, **batch.LK_GRID_DEFAULTS)
    out = batch.pathfinder_filter(pts, nxt, width, height, mode=rule, all_points=True)
    m = int(out["n_kept"][0])
    # frame_layer: the kept vectors in red with magenta start points, the rejected ones in cyan when draw_bad_flow is
    # set (:179-191), rasterised on the device exactly as cv2.polylines / cv2.circle do
    frame_layer = batch.overlay_vectors(out, height, width, draw_bad=draw_bad_flow)[0].cpu().numpy()
    return frame_layer, out["kept_flow"][0, :m].cpu().numpy(), out["kept_pts"][0, :m].cpu().numpy()


def lamp_intensity(flow_):
    """Danger intensity per kept point: uint8 V = min(50 + 2*|flow|, 255) (pathfinder_viewer.py:210-217)."""
    f = torch.from_numpy(np.ascontiguousarray(flow_)).to(torch.float64)
    m = torch.sqrt(f[:, 0] * f[:, 0] + f[:, 1] * f[:, 1])
    return torch.clamp(50 + m * 2, max=255).to(torch.uint8).numpy()


def draw_sparse_lamps(flow_, points_, height, width, device="cuda"):
    """flow_ int32 (M,2), points_ int32 (M,2) as get_flow_lk returns them -> uint8 (H,W,3) BGR lamp layer
    (pathfinder_viewer.py:196-223; the reference reads height / width from module globals)."""
    m = len(points_)
    filt = dict(kept_pts=torch.from_numpy(np.ascontiguousarray(points_, dtype=np.int32)).to(device).reshape(1, m, 2),
                danger_v=torch.from_numpy(lamp_intensity(flow_)).to(device).reshape(1, m),
                n_kept=torch.tensor([m], dtype=torch.int32, device=device))
    return batch.overlay_lamps(filt, height, width)[0].cpu().numpy()


def draw_hsv(flow_, device="cuda"):
    """flow_: float32 (H,W,2) numpy, as the reference passes it -> uint8 (H,W,3) BGR (pathfinder_viewer.py:124-141)."""
    f = torch.from_numpy(np.ascontiguousarray(flow_, dtype=np.float32)).to(device)[None]
    return batch.flow_hsv(f)[0].cpu().numpy()


class PathfinderPipeline:
    """Full per-frame pipeline on device-resident BGR frames (config 5): gray -> grid LK (current -> previous)
    -> vector filter + danger points [+ dense Farneback flow and its statistics]."""

    def __init__(self, height, width, step=30, dense=False, chunk_pairs=4, device=None, side_stream=False):
        self.h, self.w = int(height), int(width)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.points = torch.from_numpy(grid_points(self.w, self.h, step)).to(self.device)
        self.dense = batch.FarnebackEngine(self.h, self.w, chunk_pairs=chunk_pairs, device=self.device) if dense else None
        # The sparse branch (grid LK + filter) and the dense branch only share the gray frames; side_stream=True runs
        # the sparse branch on a side stream next to the dense one.  Off by default: at 4K x 8 pairs it measured 1008
        # pairs/s in one run and 746 in the next against 981 on one stream -- lk_track's thousands of small CTAs keep
        # landing on SMs that a 160 KB fb_iter_ws CTA needs empty, and which kernel starves is a matter of timing.
        self._side = torch.cuda.Stream(device=self.device) if dense and side_stream else None

    def _sparse(self, gray):
        prev, cur = gray[:-1], gray[1:]
        nxt, status, err = batch.pyrlk(cur, prev, self.points, **batch.LK_GRID_DEFAULTS)
        out = batch.pathfinder_filter(self.points, nxt, self.w, self.h)
        out.update(next_pts=nxt, status=status, err=err)
        return out

    def run(self, bgr_frames):
        """uint8 (F,H,W,3) -> dict with per-pair outputs for the F-1 consecutive pairs."""
        gray = batch.bgr2gray(bgr_frames)
        main = torch.cuda.current_stream(self.device)
        if self._side is not None:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                out = self._sparse(gray)
        else:
            out = self._sparse(gray)
        out["gray"] = gray
        if self.dense is not None:
            stats = torch.empty((gray.shape[0] - 1, 8), dtype=torch.float32, device=self.device)
            out["flow"] = self.dense.flow_sequence(gray, stats=stats)      # statistics reduced inside the last kernel
            out["flow_stats"] = stats
            if self._side is not None:
                main.wait_stream(self._side)
                for v in out.values():                                      # allocated on the side stream, used here
                    if torch.is_tensor(v):
                        v.record_stream(main)
            # the same vector filter / danger points driven by the dense field sampled on the grid.
            # The viewer tracks current -> previous; the dense field is previous -> current, hence the sign.
            back = self.points[None] - (batch.flow_sample(out["flow"], self.points) - self.points[None])
            dense_out = batch.pathfinder_filter(self.points, back.contiguous(), self.w, self.h)
            out.update({"dense_" + k: v for k, v in dense_out.items()})
        return out
