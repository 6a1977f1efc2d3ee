"""CPU: the drop-in accepts exactly the argument forms the reference's call sites use (no GPU, no compute:
``inspect.signature(...).bind``), and rejects what cv2 rejects before anything reaches the device."""
import inspect

import numpy as np
import pytest

from hackathonopticalflow_b200 import cv2compat as b2

G = np.zeros((8, 8), np.uint8)
PTS = np.zeros((4, 2), np.float32)


def _binds(fn, *a, **k):
    inspect.signature(fn).bind(*a, **k)


def test_reference_call_forms_bind():
    # pathfinder_viewer.py:244, :280 / DenseOF.py:481, :510 / SparseOF.py:28
    _binds(b2.cvtColor, np.zeros((8, 8, 3), np.uint8), b2.COLOR_BGR2GRAY)
    # cv2's full positional form, hint included (cv2.cvtColor(src, code[, dst[, dstCn[, hint]]]))
    _binds(b2.cvtColor, np.zeros((8, 8, 3), np.uint8), b2.COLOR_BGR2GRAY, None, 0, b2.ALGO_HINT_DEFAULT)
    _binds(b2.cvtColor, src=np.zeros((8, 8, 3), np.uint8), code=6, hint=0)
    # DenseOF.py:147-156: all ten by keyword
    _binds(b2.calcOpticalFlowFarneback, prev=G, next=G, flow=None, pyr_scale=0.5, levels=3, winsize=15, iterations=3,
           poly_n=5, poly_sigma=1.2, flags=0)
    # pathfinder_viewer.py:156-158 / DenseOF.py:183-185
    _binds(b2.calcOpticalFlowPyrLK, G, G, PTS, None, winSize=(45, 45), maxLevel=2,
           criteria=(b2.TERM_CRITERIA_EPS | b2.TERM_CRITERIA_COUNT, 10, 0.03))
    # SparseOF.py:35-36 with lk_params (:6-8)
    lk_params = dict(winSize=(15, 15), maxLevel=2, criteria=(b2.TERM_CRITERIA_EPS | b2.TERM_CRITERIA_COUNT, 10, 0.03))
    _binds(b2.calcOpticalFlowPyrLK, G, G, PTS.reshape(-1, 1, 2), None, **lk_params)
    # SparseOF.py:69 with feature_params (:10-13)
    feature_params = dict(maxCorners=20, qualityLevel=0.3, minDistance=10, blockSize=7)
    _binds(b2.goodFeaturesToTrack, G, mask=G, **feature_params)
    # cv2's two positional overloads
    _binds(b2.goodFeaturesToTrack, G, 20, 0.3, 10, None, None, 7, False, 0.04)
    _binds(b2.goodFeaturesToTrack, G, 20, 0.3, 10, None, None, 7, 3, False, 0.04)


def test_constants_match_cv2_values():
    assert (b2.COLOR_BGR2GRAY, b2.TERM_CRITERIA_COUNT, b2.TERM_CRITERIA_EPS) == (6, 1, 2)
    assert (b2.OPTFLOW_USE_INITIAL_FLOW, b2.OPTFLOW_LK_GET_MIN_EIGENVALS, b2.OPTFLOW_FARNEBACK_GAUSSIAN) == (4, 8, 256)
    try:
        import cv2
    except Exception:
        return
    for name in ("COLOR_BGR2GRAY", "TERM_CRITERIA_COUNT", "TERM_CRITERIA_EPS", "OPTFLOW_USE_INITIAL_FLOW",
                 "OPTFLOW_LK_GET_MIN_EIGENVALS", "OPTFLOW_FARNEBACK_GAUSSIAN"):
        assert getattr(b2, name) == getattr(cv2, name), name


def test_argument_errors_are_raised_before_the_device_is_touched():
    with pytest.raises(b2.error):
        b2.cvtColor(np.zeros((8, 8), np.uint8), b2.COLOR_BGR2GRAY)                 # not 3 channels
    with pytest.raises(b2.error):
        b2.cvtColor(np.zeros((8, 8, 3), np.uint8), 7)                              # another conversion code
    with pytest.raises(b2.error):
        b2.calcOpticalFlowFarneback(G, np.zeros((8, 9), np.uint8), None, 0.5, 3, 15, 3, 5, 1.2, 0)
    with pytest.raises(b2.error):
        b2.calcOpticalFlowFarneback(G, G, None, 1.0, 3, 15, 3, 5, 1.2, 0)          # pyr_scale < 1
    with pytest.raises(b2.error):
        b2.calcOpticalFlowFarneback(G, G, None, 0.5, 3, 15, 3, 5, 1.2, 4)          # initial flow flag, no flow
    with pytest.raises(b2.error):
        b2.calcOpticalFlowPyrLK(G, G, PTS.astype(np.float64), None)                # cv2: checkVector(2, CV_32F)
    with pytest.raises(b2.error):
        b2.calcOpticalFlowPyrLK(G, G, PTS, None, winSize=(2, 2))
    with pytest.raises(b2.error):
        b2.goodFeaturesToTrack(G, 20, 0.0, 10)                                     # qualityLevel > 0
    with pytest.raises(b2.error):
        b2.goodFeaturesToTrack(G, 20, 0.3, 10, mask=np.zeros((4, 4), np.uint8))    # mask size
    with pytest.raises(b2.error):
        b2.goodFeaturesToTrack(G, 20, 0.3, 10, gradientSize=4)                     # Sobel aperture
