"""Dense flow against LIVE cv2 over random frame sizes and parameter sets (well-conditioned synthetic texture, so that
the comparison is not limited by cv2's own reproducibility): mean / max EPE per case, sorted by max."""
import os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hackathonopticalflow_b200 import cv2compat as b2, synth
cv2.setNumThreads(8)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
INIT = len(sys.argv) > 3
res = []
for c in range(N):
    h, w = int(rng.integers(33, 700)), int(rng.integers(33, 900))
    if c % 7 == 0:
        h, w = [(1080, 1920), (720, 1280), (33, 33), (64, 2000), (1500, 40), (481, 641)][(c // 7) % 6]
    args = dict(pyr_scale=float(rng.choice([0.5, 0.5, 0.6, 0.75, 0.8, 0.9])), levels=int(rng.integers(1, 7)),
                winsize=int(rng.choice([5, 9, 15, 15, 16, 21, 31, 45])), iterations=int(rng.integers(1, 5)),
                poly_n=int(rng.choice([5, 5, 7])), poly_sigma=float(rng.choice([1.1, 1.2, 1.5])),
                flags=int(rng.choice([0, 0, 256])))
    if c % 5 == 0:
        args.update(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)   # the reference's set
    fr = synth.sequence(h, w, 2, seed=500 + c)
    init = None
    if INIT and c % 3 == 1:                     # caller-supplied initial field (strided every other time)
        args["flags"] |= 4
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        init = np.stack([3 * np.sin(yy / 37) + rng.normal(0, 0.3, (h, w)), 2 * np.cos(xx / 51) + rng.normal(0, 0.3, (h, w))], -1).astype(np.float32)
        if c % 2:
            big = np.zeros((h, w, 4), np.float32); big[..., :2] = init; init = big[..., :2]
    try:
        want = cv2.calcOpticalFlowFarneback(fr[0], fr[1], None if init is None else init.copy(), **args)
    except cv2.error as e:
        print("cv2 rejects", h, w, args); continue
    try:
        got = b2.calcOpticalFlowFarneback(fr[0], fr[1], None if init is None else init.copy(), **args)
    except Exception as e:
        print("b200 raises", h, w, args, repr(e)[:200]); continue
    d = np.sqrt(((got.astype(np.float64) - want) ** 2).sum(-1))
    mag = np.sqrt((want.astype(np.float64) ** 2).sum(-1))
    res.append((d.max(), d.mean(), h, w, args, mag.max()))
    if not np.isfinite(d).all() or d.max() > 0.05:
        y, x = np.unravel_index(np.nanargmax(d), d.shape)
        print("LARGE", "%.4f" % d.max(), "mean %.2e" % d.mean(), h, w, args, "at", (y, x), "flow max %.1f" % mag.max(), flush=True)
res.sort(key=lambda r: -r[0])
print("cases", len(res), "worst max %.5f" % res[0][0], "worst mean %.2e" % max(r[1] for r in res))
for r in res[:8]:
    print("  max %.5f mean %.2e  %dx%d %s flow max %.1f" % (r[0], r[1], r[2], r[3], r[4], r[5]))
