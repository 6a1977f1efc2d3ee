"""Developer check run on the GPU box: gray / pyrDown / Farneback against live cv2 + first timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch
from hackathonopticalflow_b200 import cv2compat as b2, batch, synth, _lib

rng = np.random.default_rng(0)
for (h, w) in [(1080, 1920), (101, 77), (1, 17), (33, 1919)]:
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    print("gray", h, w, np.array_equal(b2.cvtColor(img, b2.COLOR_BGR2GRAY), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)))
for (h, w) in [(1080, 1920), (101, 77), (135, 241), (540, 960), (3, 5), (720, 1280)]:
    g = rng.integers(0, 256, (h, w), dtype=np.uint8)
    print("pyrdown", h, w, np.array_equal(b2.pyrDown(g), cv2.pyrDown(g)))

def epe(a, b):
    d = np.sqrt(((a - b) ** 2).sum(-1))
    return d.mean(), d.max()

for (h, w) in [(270, 480), (135, 241), (384, 683), (720, 1280), (1080, 1920)]:
    fr = synth.sequence(h, w, 3, seed=1000)
    for args in [(0.5, 3, 15, 3, 5, 1.2, 0), (0.5, 3, 15, 3, 5, 1.2, 256), (0.8, 5, 13, 2, 7, 1.5, 0), (0.5, 2, 16, 3, 5, 1.1, 0)]:
        if h >= 720 and args[0] != 0.5: continue
        ref = cv2.calcOpticalFlowFarneback(fr[0], fr[1], None, *args)
        t = time.time(); mine = b2.calcOpticalFlowFarneback(fr[0], fr[1], None, *args); dt = time.time() - t
        m, mx = epe(ref, mine)
        print(f"farneback {h}x{w} {args}: mean EPE {m:.3e} max {mx:.3e} |flow| {np.abs(ref).mean():.2f}  call {dt*1e3:.1f} ms")

# batched device path timing at 1080p
h, w = 1080, 1920
fr = synth.sequence(h, w, 9, seed=1001)
frames = torch.from_numpy(fr).cuda()
eng = batch.FarnebackEngine(h, w, chunk_pairs=8)
out = eng.flow_sequence(frames)
torch.cuda.synchronize()
ref = cv2.calcOpticalFlowFarneback(fr[3], fr[4], None, 0.5, 3, 15, 3, 5, 1.2, 0)
print("seq parity pair 3:", epe(ref, out[3].cpu().numpy()))
outp = eng.flow_pairs(frames[:-1].contiguous(), frames[1:].contiguous())
print("pairs == sequence:", torch.equal(out, outp))
for name, fn in [("sequence", lambda: eng.flow_sequence(frames, out)), ("pairs", lambda: eng.flow_pairs(frames[:-1], frames[1:], outp))]:
    fr0 = frames[:-1].contiguous(); fr1 = frames[1:].contiguous()
    if name == "pairs": fn = lambda: eng.flow_pairs(fr0, fr1, outp)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms per 8 pairs -> {8/ms*1e3:.1f} pairs/s")
print("launches", _lib.lib().b2of_launch_count())
