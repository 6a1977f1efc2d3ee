"""Front-end kernels (K1 bgr2gray, K2 pyrDown) on device-resident batches: time per call and algorithmic GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hackathonopticalflow_b200 import batch
torch.manual_seed(0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (h, w, b) in [(1080, 1920, 64), (2160, 3840, 16), (540, 960, 64)]:
    bgr = torch.randint(0, 256, (b, h, w, 3), dtype=torch.uint8, device="cuda")
    gray = batch.bgr2gray(bgr)
    out = torch.empty_like(gray)
    ms = timeit(lambda: batch.bgr2gray(bgr, out=out))
    print(f"bgr2gray {b}x{h}x{w}: {ms*1e3:8.1f} us  {b*h*w*4/ms/1e6:8.1f} GB/s")
    d = batch.pyrdown(gray)
    ms = timeit(lambda: batch.pyrdown(gray, out=d))
    print(f"pyrdown  {b}x{h}x{w}: {ms*1e3:8.1f} us  {b*h*w*1.25/ms/1e6:8.1f} GB/s")
