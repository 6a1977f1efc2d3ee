"""HBM bandwidth by read : write mix (torch ops, CUDA events, best of 10): what a write-heavy kernel can reach on this
GPU compared with the copy figure in MEASURED_PEAKS.json (1 : 1)."""
import torch, json
dev = torch.device("cuda:0")
N = 1 << 28                       # 1 GiB of float32 per operand
a = torch.empty(N, dtype=torch.float32, device=dev).normal_()
b = torch.empty(N, dtype=torch.float32, device=dev)
big = torch.empty(5, N // 4, dtype=torch.float32, device=dev)
small = a[: N // 4]
def t(fn, bytes_):
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return bytes_ / best / 1e6
out = {
    "write_only_fill": t(lambda: b.fill_(1.0), 4 * N),
    "copy_1r_1w": t(lambda: b.copy_(a), 8 * N),
    "read_only_sum": t(lambda: a.sum(), 4 * N),
    "bcast_1r_5w": t(lambda: big.copy_(small.expand(5, -1)), 6 * N),
    "add_2r_1w": t(lambda: torch.add(a, a.flip(0) if False else b, out=b), 12 * N),
}
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
