"""hackathonopticalflow_b200 -- B200-native (sm_100a) optical-flow engine for the per-frame-pair hot path of
spirinis/HackathonOpticalFlow: BGR->gray, Farneback dense flow, Shi-Tomasi corners, pyramidal Lucas-Kanade and
the viewer's vector filter / danger points.

* ``cv2compat`` -- the drop-in: same call signatures and return arrays as the four cv2 calls the reference makes.
* ``batch``     -- device-resident batched API (torch tensors as buffers; no host copies).
* ``pathfinder``-- the reference's own downstream logic (grid, vector filter, danger points) on the device.
* ``dist``      -- frame-sharded data parallelism (one process per GPU, NCCL gather of per-frame stats).

Everything computes in hand-written CUDA behind the C-ABI of ``csrc/libb2of.so`` (``include/b2of.h``).
"""
from ._lib import error, build  # noqa: F401

__all__ = ["error", "build", "cv2compat", "batch", "pathfinder", "dist", "synth"]
