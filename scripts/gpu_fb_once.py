"""Profiling target: one warm-up + one dense-flow pass over PAIRS consecutive 1080p pairs (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hackathonopticalflow_b200 import batch, synth
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
h, w = 1080, 1920
fr = synth.sequence(h, w, min(pairs + 1, 17), seed=1001)
import numpy as np
idx = [i % fr.shape[0] for i in range(pairs + 1)]
frames = torch.from_numpy(np.ascontiguousarray(fr[idx])).cuda()
eng = batch.FarnebackEngine(h, w, chunk_pairs=pairs)
out = eng.flow_sequence(frames)
torch.cuda.synchronize()
out = eng.flow_sequence(frames, out)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
