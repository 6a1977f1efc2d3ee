"""CPU: host-side logic -- grid generator, shard arithmetic, cv2-style argument errors, 2-rank gloo gather."""
import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import have_cv2

from conftest import ROOT
from hackathonopticalflow_b200 import cv2compat as b2
from hackathonopticalflow_b200 import dist as b2dist
from hackathonopticalflow_b200 import error, pathfinder
from oracle import pathfinder as opf


@pytest.mark.parametrize("w,h,step", [(1920, 1080, 30), (1280, 720, 30), (3840, 2160, 30), (640, 360, 30),
                                      (641, 363, 14), (1920, 1080, 100), (100, 60, 30)])
def test_grid_points_match_reference_restatement(w, h, step):
    a, b = pathfinder.grid_points(w, h, step), opf.grid_points(w, h, step)
    assert a.dtype == np.float32 and np.array_equal(a, b)


def test_grid_1080p_is_the_viewers_2304_points():
    g = pathfinder.grid_points(1920, 1080)
    assert g.shape == (2304, 2) and g[0].tolist() == [15, 15] and g[1].tolist() == [15, 45]  # x-major


@pytest.mark.parametrize("frames,world", [(17, 1), (17, 2), (17, 4), (17, 8), (257, 8), (5, 8), (2, 2), (1, 4)])
def test_shard_covers_all_pairs_once_with_one_frame_halo(frames, world):
    seen = []
    for r in range(world):
        lo, hi, flo, fhi = b2dist.shard(frames, r, world)
        assert 0 <= lo <= hi <= max(frames - 1, 0)
        if hi > lo:
            assert (flo, fhi) == (lo, hi + 1)  # frames [lo, hi] : halo frame hi
        seen += list(range(lo, hi))
    assert seen == list(range(max(frames - 1, 0)))


def test_cv2_style_argument_errors():
    g = np.zeros((40, 50), np.uint8)
    with pytest.raises(error, match="pyrScale_ < 1"):
        b2.calcOpticalFlowFarneback(g, g, None, 1.0, 3, 15, 3, 5, 1.2, 0)
    with pytest.raises(error, match="prev0.size"):
        b2.calcOpticalFlowFarneback(g, g[:-1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    with pytest.raises(error, match="CV_8UC1"):
        b2.calcOpticalFlowFarneback(g.astype(np.float32), g, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    with pytest.raises(error, match="checkVector"):
        b2.calcOpticalFlowPyrLK(g, g, np.zeros((4, 2), np.float64), None)
    with pytest.raises(error, match="maxLevel >= 0"):
        b2.calcOpticalFlowPyrLK(g, g, np.zeros((4, 2), np.float32), None, winSize=(2, 2))
    with pytest.raises(error, match="qualityLevel > 0"):
        b2.goodFeaturesToTrack(g, 10, 0.0, 5)
    with pytest.raises(error, match="_mask"):
        b2.goodFeaturesToTrack(g, 10, 0.1, 5, mask=np.zeros((3, 3), np.uint8))
    with pytest.raises(error):
        b2.cvtColor(g, b2.COLOR_BGR2GRAY)  # not 3-channel


def test_empty_point_set_returns_what_cv2_returns_without_gpu():
    g = np.zeros((40, 50), np.uint8)
    for shape in ((0, 1, 2), (0, 2)):
        got = b2.calcOpticalFlowPyrLK(g, g, np.zeros(shape, np.float32), None)
        assert got == (None, None, None)
        if have_cv2():
            import cv2
            assert cv2.calcOpticalFlowPyrLK(g, g, np.zeros(shape, np.float32), None) == got


def test_gather_stats_two_ranks_gloo(tmp_path):
    """world_size 2 on CPU (gloo): shard 9 frames, gather the per-pair rows to rank 0 in pair order."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import torch
from hackathonopticalflow_b200 import dist as d
rank, world, _ = d.init_from_env("gloo")
F = 9
lo, hi, flo, fhi = d.shard(F, rank, world)
local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, d.STATS_WIDTH)
out = d.gather_stats(local, F, rank, world)
if rank == 0:
    assert out.shape == (F - 1, d.STATS_WIDTH), out.shape
    assert out[:, 0].tolist() == [float(i) for i in range(F - 1)], out[:, 0]
    print("GATHER_OK")
else:
    assert out is None
m = d.max_over_ranks(float(rank + 1), "cpu")
assert m == 2.0
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    port = 29500 + os.getpid() % 2000
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GATHER_OK" in res.stdout


def test_sharded_pairs_with_halo_equal_serial_two_ranks_gloo(tmp_path):
    """world_size 2 on CPU: each rank takes its frame shard (one-frame halo), computes a per-pair record with the
    numpy oracle standing in for the GPU engine, rank 0 gathers; the result must equal the serial pass."""
    script = tmp_path / "w2.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch
from hackathonopticalflow_b200 import dist as d
from oracle import farneback as ofb
rank, world, _ = d.init_from_env("gloo")
rng = np.random.default_rng(0)
base = (np.kron(rng.random((12, 16)), np.ones((4, 4))) * 200 + 20).astype(np.uint8)     # 48 x 64
frames = np.stack([np.roll(base, (t, 2 * t), axis=(0, 1)) for t in range(6)])             # 6 frames, 5 pairs
def record(a, b):
    f = ofb.farneback(a, b, None, 0.5, 1, 9, 1, 5, 1.1, 0)
    m = np.sqrt((f ** 2).sum(-1))
    return [m.mean(), m.max(), f[..., 0].mean(), f[..., 1].mean(), 0, 0, 0, 0]
lo, hi, flo, fhi = d.shard(len(frames), rank, world)
mine = frames[flo:fhi]                                     # includes the halo frame
assert len(mine) == (hi - lo) + 1
local = torch.tensor([record(mine[i], mine[i + 1]) for i in range(hi - lo)], dtype=torch.float32).reshape(-1, 8)
out = d.gather_stats(local, len(frames), rank, world)
if rank == 0:
    serial = torch.tensor([record(frames[i], frames[i + 1]) for i in range(len(frames) - 1)], dtype=torch.float32)
    assert out.shape == serial.shape and torch.equal(out, serial), (out, serial)
    print("SHARD_OK")
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    port = 31500 + os.getpid() % 2000
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "SHARD_OK" in res.stdout
