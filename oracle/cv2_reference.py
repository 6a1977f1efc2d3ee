"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's own CPU path: the live ``cv2`` wheel called with exactly the
arguments the reference's call sites pass.  This is the authority the CUDA path
and the numpy restatements are both judged against, and the CPU baseline that
``bench.py`` times (``cpu_baseline.kind == "reference"``).
"""
import os
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np

# DenseOF.py:127-128 defaults, passed through at DenseOF.py:147-156
FARNEBACK_PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
# pathfinder_viewer.py:154-158
LK_GRID_PARAMS = dict(winSize=(45, 45), maxLevel=2,
                      criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
# SparseOF.py:6-8
LK_TRACK_PARAMS = dict(winSize=(15, 15), maxLevel=2,
                       criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
# SparseOF.py:10-13
FEATURE_PARAMS = dict(maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7)


def gray(bgr):
    """pathfinder_viewer.py:280."""
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)


def farneback(prev_gray, gray_, flow=None, **over):
    """DenseOF.py:520 -> :147-156."""
    p = dict(FARNEBACK_PARAMS)
    p.update(over)
    return cv2.calcOpticalFlowFarneback(prev=prev_gray, next=gray_, flow=flow, **p)


def lk_grid(prev_gray, gray_, points):
    """pathfinder_viewer.py:156-158: note the argument order (current frame first)."""
    return cv2.calcOpticalFlowPyrLK(gray_, prev_gray, points, None, **LK_GRID_PARAMS)


def lk_track(img0, img1, p0):
    """SparseOF.py:35-38: forward, backward, forward-backward check."""
    p1, st, err = cv2.calcOpticalFlowPyrLK(img0, img1, p0, None, **LK_TRACK_PARAMS)
    p0r, st_b, err_b = cv2.calcOpticalFlowPyrLK(img1, img0, p1, None, **LK_TRACK_PARAMS)
    d = abs(p0 - p0r).reshape(-1, 2).max(-1)
    return p1, p0r, d < 1, st, st_b


def track_mask(shape, live_points):
    """SparseOF.py:61-66."""
    mask = np.zeros(shape, np.uint8)
    mask[:] = 255
    for x, y in [np.int32(p) for p in live_points]:
        cv2.circle(mask, (int(x), int(y)), 5, 0, -1)
    return mask


def features(gray_, mask=None, **over):
    """SparseOF.py:69."""
    p = dict(FEATURE_PARAMS)
    p.update(over)
    return cv2.goodFeaturesToTrack(gray_, mask=mask, **p)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def farneback_pool(frames, workers=None):
    """Pair-parallel throughput mode: cv2 Farneback is serial inside,
    so one cv2 thread per pair across all host cores is the fair multi-core figure."""
    workers = workers or host_cores()
    prev_threads = cv2.getNumThreads()
    cv2.setNumThreads(1)
    try:
        with ThreadPoolExecutor(workers) as ex:
            out = list(ex.map(lambda i: farneback(frames[i], frames[i + 1]), range(len(frames) - 1)))
    finally:
        cv2.setNumThreads(prev_threads)
    return out
