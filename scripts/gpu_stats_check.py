import torch, numpy as np, sys
sys.path.insert(0, '/root/repo')
from hackathonopticalflow_b200 import batch, synth
for (h, w, n) in [(1080, 1920, 4), (270, 480, 3), (135, 241, 2), (37, 53, 2)]:
    fr = torch.from_numpy(synth.sequence(h, w, n, seed=5)).cuda()
    eng = batch.FarnebackEngine(h, w, chunk_pairs=2)
    st = torch.empty((n - 1, 8), device='cuda')
    flow = eng.flow_sequence(fr, stats=st)
    flow2 = eng.flow_sequence(fr)
    ref = batch.flow_stats(flow)
    print(h, w, 'flow same', torch.equal(flow, flow2), 'stats maxrel', ((st - ref).abs() / ref.abs().clamp_min(1e-6)).max().item(), st[0, :4].tolist(), ref[0, :4].tolist())
