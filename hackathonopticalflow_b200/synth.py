"""Seeded synthetic FPV-like frame sequences: a multi-scale smooth random texture seen by
a camera that flies forward (zoom 1.5 %/frame about the centre) while drifting.  Test/bench input only.

Pure torch-on-CPU (no cv2): deterministic for a given seed.
"""
import numpy as np
import torch
import torch.nn.functional as F

PAD = 64
ZOOM = 1.015
DRIFT = (1.7, -0.9)


def canvas(h, w, seed):
    """float32 (3, h+2*PAD, w+2*PAD) in [0,1]."""
    g = torch.Generator().manual_seed(int(seed))
    H, W = h + 2 * PAD, w + 2 * PAD
    acc = torch.zeros(1, 3, H, W)
    for cell, amp in ((2, 0.15), (8, 0.55), (32, 0.30)):
        noise = torch.rand(1, 3, H // cell + 3, W // cell + 3, generator=g)
        acc += amp * F.interpolate(noise, size=(H, W), mode="bicubic", align_corners=False)
    acc -= acc.amin()
    acc /= acc.amax()
    return acc[0]


def frame(cv, h, w, t):
    """uint8 BGR (h,w,3): frame t of the flight over canvas ``cv``."""
    z = ZOOM ** t
    H, W = cv.shape[1:]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    sx = (xs - (w - 1) / 2) / z + (W - 1) / 2 + DRIFT[0] * t
    sy = (ys - (h - 1) / 2) / z + (H - 1) / 2 + DRIFT[1] * t
    grid = torch.stack([sx / (W - 1) * 2 - 1, sy / (H - 1) * 2 - 1], -1)[None]
    img = F.grid_sample(cv[None], grid, mode="bilinear", padding_mode="border", align_corners=True)[0]
    return (img.clamp(0, 1) * 255 + 0.5).to(torch.uint8).permute(1, 2, 0).contiguous().numpy()


def to_gray(bgr):
    """Same integer luma as cv2.cvtColor(BGR2GRAY) (numpy; for building inputs only)."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def sequence(h, w, n_frames, seed=1000, gray=True):
    """uint8 (n_frames,h,w) gray, or (n_frames,h,w,3) BGR."""
    cv = canvas(h, w, seed)
    fr = [frame(cv, h, w, t) for t in range(n_frames)]
    if gray:
        return np.stack([to_gray(f) for f in fr])
    return np.stack(fr)
