"""Device-resident batched API: torch tensors are the buffers, the C-ABI does the work.

No host copies happen here; inputs and outputs stay in HBM (this is what ``bench.py``'s ``value`` times).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FarnebackParams, LKParams, GFTTParams

# reference call-site parameters
FARNEBACK_DEFAULTS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2,
                          flags=0)  # DenseOF.py:127-128
LK_GRID_DEFAULTS = dict(winSize=(45, 45), maxLevel=2, criteria=(3, 10, 0.03))  # pathfinder_viewer.py:154-158
LK_TRACK_DEFAULTS = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))  # SparseOF.py:6-8
GFTT_DEFAULTS = dict(maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7)  # SparseOF.py:10-13


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_u8_frames(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 3
            and t.is_contiguous()):
        raise _lib.error(-215, f"{name} must be a contiguous CUDA uint8 tensor (N,H,W)")


def _check_out(t, shape, dtype, device, name):
    """A caller-supplied output tensor must be exactly what the kernel will write (the C-ABI takes bare pointers)."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == device and t.dtype == dtype
            and tuple(t.shape) == tuple(shape) and t.is_contiguous()):
        raise _lib.error(-215, f"{name} must be a contiguous {dtype} CUDA tensor of shape {tuple(shape)} on {device}")


def bgr2gray(bgr, out=None):
    """uint8 (N,H,W,3) or (H,W,3) CUDA tensor -> uint8 (N,H,W) / (H,W).  Bit-exact with cv2.cvtColor(BGR2GRAY)."""
    squeeze = bgr.dim() == 3
    b = bgr.unsqueeze(0) if squeeze else bgr
    if not (b.is_cuda and b.dtype == torch.uint8 and b.dim() == 4 and b.shape[-1] == 3 and b.is_contiguous()):
        raise _lib.error(-215, "bgr must be a contiguous CUDA uint8 tensor (N,H,W,3)")
    n, h, w, _ = b.shape
    if out is None:
        out = torch.empty((n, h, w), dtype=torch.uint8, device=b.device)
    _check_out(out, (n, h, w), torch.uint8, b.device, "out")
    with torch.cuda.device(b.device):
        _lib.check(_lib.lib().b2of_bgr2gray_u8_dev(_p(b), h, w, w * 3, h * w * 3, _p(out), w, h * w, n, _stream()))
    return out[0] if squeeze else out


def pyrdown(img, out=None):
    """uint8 (N,H,W) -> uint8 (N,(H+1)//2,(W+1)//2).  Bit-exact with cv2.pyrDown."""
    _check_u8_frames(img, "img")
    n, h, w = img.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    if out is None:
        out = torch.empty((n, dh, dw), dtype=torch.uint8, device=img.device)
    _check_out(out, (n, dh, dw), torch.uint8, img.device, "out")
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().b2of_pyrdown_u8_dev(_p(img), h, w, w, h * w, _p(out), dw, dh * dw, n, _stream()))
    return out


class FarnebackEngine:
    """Batched dense flow on device-resident frames.

    ``chunk_pairs`` pairs are in flight per pass; the scratch (level images, polynomial expansions, flow
    ping/pong) is allocated once for that many.
    """

    def __init__(self, rows, cols, chunk_pairs=8, device=None, **params):
        p = dict(FARNEBACK_DEFAULTS)
        p.update(params)
        self.rows, self.cols = int(rows), int(cols)
        self.params = FarnebackParams(float(p["pyr_scale"]), int(p["levels"]), int(p["winsize"]),
                                      int(p["iterations"]), int(p["poly_n"]), float(p["poly_sigma"]), int(p["flags"]))
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.chunk_pairs = int(chunk_pairs)
        l = _lib.lib()
        with torch.cuda.device(self.device):
            need = max(l.b2of_farneback_workspace_bytes(self.rows, self.cols, C.byref(self.params), self.chunk_pairs, 0),
                       l.b2of_farneback_workspace_bytes(self.rows, self.cols, C.byref(self.params), self.chunk_pairs, 1))
        if need == 0:
            _lib.check(-215)
        self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)

    def flow_sequence(self, frames, out=None, stats=None):
        """uint8 (F,H,W) consecutive frames -> float32 (F-1,H,W,2); per-frame work is done once per frame.

        ``stats``: optional float32 (F-1,8) CUDA tensor that receives the per-pair flow statistics of
        :func:`flow_stats`, reduced inside the last iteration kernel (no second pass over the flow fields)."""
        _check_u8_frames(frames, "frames")
        f, h, w = frames.shape
        if (h, w) != (self.rows, self.cols):
            raise _lib.error(-215, f"frames are {h}x{w}, the engine was built for {self.rows}x{self.cols}")
        shape = (max(f - 1, 0), h, w, 2)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=frames.device)
        _check_out(out, shape, torch.float32, frames.device, "out")
        if stats is not None:
            _check_out(stats, (shape[0], 8), torch.float32, frames.device, "stats")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b2of_farneback_sequence_stats_dev(_p(frames), w, h * w, f, h, w,
                                                                    C.byref(self.params), _p(out), _p(stats),
                                                                    _p(self.workspace), self.workspace.numel(),
                                                                    _stream()))
        return out

    def flow_pairs(self, prev, next, out=None):
        """uint8 (B,H,W) prev and next stacks (independent pairs) -> float32 (B,H,W,2)."""
        _check_u8_frames(prev, "prev")
        _check_u8_frames(next, "next")
        assert prev.shape == next.shape
        b, h, w = prev.shape
        if (h, w) != (self.rows, self.cols):
            raise _lib.error(-215, f"frames are {h}x{w}, the engine was built for {self.rows}x{self.cols}")
        if out is None:
            out = torch.empty((b, h, w, 2), dtype=torch.float32, device=prev.device)
        _check_out(out, (b, h, w, 2), torch.float32, prev.device, "out")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b2of_farneback_pairs_dev(_p(prev), _p(next), w, h * w, b, h, w,
                                                           C.byref(self.params), _p(out), _p(self.workspace),
                                                           self.workspace.numel(), _stream()))
        return out


def _lk_params(winSize, maxLevel, criteria, flags, minEigThreshold):
    return LKParams(int(winSize[0]), int(winSize[1]), int(maxLevel), int(criteria[0]), int(criteria[1]),
                    float(criteria[2]), int(flags), float(minEigThreshold))


def pyrlk(prev, next, prev_pts, next_pts=None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01), flags=0,
          minEigThreshold=1e-4, workspace=None):
    """Batched cv2.calcOpticalFlowPyrLK on device tensors.

    prev/next: uint8 (B,H,W); prev_pts: float32 (N,2) shared by every pair, or (B,N,2).
    Returns next_pts float32 (B,N,2), status uint8 (B,N), err float32 (B,N).
    """
    _check_u8_frames(prev, "prev")
    _check_u8_frames(next, "next")
    assert prev.shape == next.shape
    b, h, w = prev.shape
    shared = prev_pts.dim() == 2
    n = prev_pts.shape[-2]
    if not (prev_pts.is_cuda and prev_pts.dtype == torch.float32 and prev_pts.is_contiguous()
            and prev_pts.shape[-1] == 2 and (shared or prev_pts.shape[0] == b)):
        raise _lib.error(-215, "(npoints = prevPtsMat.checkVector(2, CV_32F, true)) >= 0")
    p = _lk_params(winSize, maxLevel, criteria, flags, minEigThreshold)
    l = _lib.lib()
    if next_pts is None:
        if flags & 4:
            raise _lib.error(-215, "nextPtsMat.checkVector(2, CV_32F, true) == npoints")
        next_pts = torch.empty((b, n, 2), dtype=torch.float32, device=prev.device)
    _check_out(next_pts, (b, n, 2), torch.float32, prev.device, "next_pts")
    status = torch.empty((b, n), dtype=torch.uint8, device=prev.device)
    err = torch.empty((b, n), dtype=torch.float32, device=prev.device)
    with torch.cuda.device(prev.device):
        need = l.b2of_pyrlk_workspace_bytes(h, w, C.byref(p), b)
        if need == 0:
            _lib.check(-215)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=prev.device)
        _lib.check(l.b2of_pyrlk_dev(_p(prev), _p(next), w, h * w, b, h, w, _p(prev_pts), 0 if shared else n, n,
                                    _p(next_pts), _p(status), _p(err), C.byref(p), _p(workspace), workspace.numel(),
                                    _stream()))
    return next_pts, status, err


def gftt(img, mask=None, maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7, useHarrisDetector=False,
         k=0.04, cap=None, workspace=None, gradientSize=3):
    """Batched cv2.goodFeaturesToTrack: uint8 (B,H,W) [+ mask (B,H,W)] -> corners float32 (B,cap,2), count int32 (B)."""
    _check_u8_frames(img, "img")
    if mask is not None:
        _check_u8_frames(mask, "mask")
        assert mask.shape == img.shape
    b, h, w = img.shape
    p = GFTTParams(int(maxCorners), float(qualityLevel), float(minDistance), int(blockSize), int(gradientSize),
                   int(bool(useHarrisDetector)), float(k))
    if cap is None:
        cap = int(maxCorners) if maxCorners > 0 else 4096
    corners = torch.zeros((b, cap, 2), dtype=torch.float32, device=img.device)
    count = torch.zeros((b,), dtype=torch.int32, device=img.device)
    l = _lib.lib()
    with torch.cuda.device(img.device):
        need = l.b2of_gftt_workspace_bytes(h, w, C.byref(p), b)
        if need == 0:
            _lib.check(-215)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=img.device)
        _lib.check(l.b2of_gftt_dev(_p(img), _p(mask), w, h * w, b, h, w, C.byref(p), _p(corners), cap, _p(count),
                                   _p(workspace), workspace.numel(), _stream()))
    return corners, count


FILTER_VIEWER, FILTER_DENSEOF = 0, 1


def pathfinder_filter(pts, next_pts, width, height, mode=FILTER_VIEWER, all_points=False):
    """The viewer's vector filter + danger intensity (pathfinder_viewer.py:159-178, :210-217) per frame.

    pts float32 (N,2) shared grid or (B,N,2); next_pts float32 (B,N,2).
    mode: FILTER_VIEWER keeps median < m < p99 (pathfinder_viewer.py:171); FILTER_DENSEOF keeps m > 1.2 * median
    (the development script's rule, DenseOF.py:228).
    all_points=True adds all_pts / all_next int32 (B,N,2): every point and its normalised end point as the reference
    rounds them (what :func:`overlay_vectors` draws).
    Returns dict(kept_pts int32 (B,N,2), kept_flow int32 (B,N,2), danger_v uint8 (B,N), mask uint8 (B,N),
                 n_kept int32 (B), stats float32 (B,8)); rows [0, n_kept[b]) of the kept_* arrays are valid.
    """
    if not (isinstance(next_pts, torch.Tensor) and next_pts.is_cuda and next_pts.dtype == torch.float32
            and next_pts.dim() == 3 and next_pts.shape[-1] == 2 and next_pts.is_contiguous()):
        raise _lib.error(-215, "next_pts must be a contiguous float32 CUDA tensor (B,N,2)")
    b, n, _ = next_pts.shape
    shared = pts.dim() == 2
    dev = next_pts.device
    _check_out(pts, (n, 2) if shared else (b, n, 2), torch.float32, dev, "pts")
    out = dict(kept_pts=torch.zeros((b, n, 2), dtype=torch.int32, device=dev),
               kept_flow=torch.zeros((b, n, 2), dtype=torch.int32, device=dev),
               danger_v=torch.zeros((b, n), dtype=torch.uint8, device=dev),
               mask=torch.zeros((b, n), dtype=torch.uint8, device=dev),
               n_kept=torch.zeros((b,), dtype=torch.int32, device=dev),
               stats=torch.zeros((b, 8), dtype=torch.float32, device=dev))
    if all_points:
        out["all_pts"] = torch.empty((b, n, 2), dtype=torch.int32, device=dev)
        out["all_next"] = torch.empty((b, n, 2), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2of_pathfinder_filter_dev(_p(pts), 0 if shared else n, _p(next_pts), n, b, int(width),
                                                         int(height), int(mode), _p(out["kept_pts"]), _p(out["kept_flow"]),
                                                         _p(out["danger_v"]), _p(out["mask"]), _p(out["n_kept"]),
                                                         _p(out["stats"]), _p(out.get("all_pts")),
                                                         _p(out.get("all_next")), _stream()))
    return out


def overlay_vectors(filt, height, width, draw_bad=True):
    """The reference's vector layer (pathfinder_viewer.py:179-191) from a :func:`pathfinder_filter` result computed with
    ``all_points=True``: uint8 (B,H,W,3) BGR, pixel for pixel what cv2.polylines / cv2.circle draw."""
    b, n, _ = filt["all_pts"].shape
    layer = torch.empty((b, int(height), int(width), 3), dtype=torch.uint8, device=filt["all_pts"].device)
    with torch.cuda.device(layer.device):
        _lib.check(_lib.lib().b2of_overlay_vectors_dev(_p(filt["all_pts"]), _p(filt["all_next"]), _p(filt["mask"]), n, b,
                                                       int(height), int(width), int(bool(draw_bad)), _p(layer),
                                                       _stream()))
    return layer


def overlay_lamps(filt, height, width):
    """The reference's danger-lamp layer (draw_sparse_lamps, pathfinder_viewer.py:196-223): uint8 (B,H,W,3) BGR."""
    b, n, _ = filt["kept_pts"].shape
    bgr = torch.empty((b, int(height), int(width), 3), dtype=torch.uint8, device=filt["kept_pts"].device)
    with torch.cuda.device(bgr.device):
        _lib.check(_lib.lib().b2of_overlay_lamps_dev(_p(filt["kept_pts"]), _p(filt["danger_v"]), _p(filt["n_kept"]), n, b,
                                                     int(height), int(width), _p(bgr), _stream()))
    return bgr


def flow_sample(flow, pts):
    """float32 (B,H,W,2) dense flow + grid float32 (N,2) or (B,N,2) -> next_pts float32 (B,N,2) = pts + flow at pts."""
    b, h, w_, _ = flow.shape
    shared = pts.dim() == 2
    n = pts.shape[-2]
    nxt = torch.empty((b, n, 2), dtype=torch.float32, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.lib().b2of_flow_sample_dev(_p(flow), b, h, w_, _p(pts), 0 if shared else n, n, _p(nxt),
                                                   _stream()))
    return nxt


def flow_hsv(flow):
    """float32 (B,H,W,2) dense flow -> uint8 (B,H,W,3) BGR picture, the reference's draw_hsv (pathfinder_viewer.py:124-141)."""
    b, h, w, _ = flow.shape
    bgr = torch.empty((b, h, w, 3), dtype=torch.uint8, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.lib().b2of_flow_hsv_dev(_p(flow), b, h, w, _p(bgr), _stream()))
    return bgr


def flow_stats(flow):
    """float32 (B,H,W,2) -> float32 (B,8): mean|flow|, max|flow|, mean dx, mean dy, 0, 0, 0, 0 (deterministic)."""
    b, h, w, _ = flow.shape
    stats = torch.empty((b, 8), dtype=torch.float32, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.lib().b2of_flow_stats_dev(_p(flow), b, h, w, _p(stats), _stream()))
    return stats
