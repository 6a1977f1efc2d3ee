"""Raw pinned H2D / D2H bandwidth on this box (ceiling for the end-to-end number)."""
import time, torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(name, f"{10 * n / (time.perf_counter() - t) / 1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("bidirectional: each direction", f"{10 * n / dt / 1e9:.1f} GB/s")
