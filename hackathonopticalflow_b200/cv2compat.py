"""Drop-in for the four ``cv2`` calls on the reference's hot path.

Use as ``from hackathonopticalflow_b200 import cv2compat as cv2`` under the
reference's own vector filtering / danger-point code.  Same arguments, same
return arrays (shape, dtype, ``None`` conventions, same-object return when an
output buffer is passed) as:

* ``cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)``      pathfinder_viewer.py:244, :280; DenseOF.py:481, :510; SparseOF.py:28
* ``cv2.calcOpticalFlowFarneback``               DenseOF.py:147-156 (call site :520)
* ``cv2.calcOpticalFlowPyrLK``                   pathfinder_viewer.py:156-158; DenseOF.py:183-185; SparseOF.py:35-36
* ``cv2.goodFeaturesToTrack``                    SparseOF.py:69

Host numpy arrays in, host numpy arrays out; every call goes through the C-ABI
of libb2of.so (ctypes) to hand-written sm_100a kernels.  No OpenCV, no Triton,
no CPU fallback: a missing library or GPU raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import error, FarnebackParams, LKParams, GFTTParams

FARNEBACK_BLOCKED_SUMS = 0x10000   # library extension, include/b2of.h: B2OF_FARNEBACK_BLOCKED_SUMS
COLOR_BGR2GRAY = 6
COLOR_RGB2GRAY = 7
COLOR_BGRA2GRAY = 10
COLOR_RGBA2GRAY = 11
BORDER_DEFAULT = 4
TERM_CRITERIA_COUNT = TERM_CRITERIA_MAX_ITER = 1
TERM_CRITERIA_EPS = 2
OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_LK_GET_MIN_EIGENVALS = 8
OPTFLOW_FARNEBACK_GAUSSIAN = 256
ALGO_HINT_DEFAULT, ALGO_HINT_ACCURATE, ALGO_HINT_APPROX = 0, 1, 2


def release():
    """Free what the library caches between calls (per-device streams and staging buffers, Farneback plans)."""
    _lib.check(_lib.lib().b2of_release())


def _assert(cond, text, fn):
    if not cond:
        raise error(-215, f"(-215:Assertion failed) {text} in function '{fn}'")


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _image_u8(a, name, fn):
    _assert(isinstance(a, np.ndarray), f"{name} is not a numpy array", fn)
    _assert(a.dtype == np.uint8 and a.ndim == 2, f"{name}.type() == CV_8UC1", fn)
    if a.strides[1] != 1 or a.strides[0] < a.shape[1]:
        a = np.ascontiguousarray(a)
    return a


def _new_host(shape, dtype):
    """Result array the caller will own.  Page-locked when torch can provide it, so the D2H lands
    straight in the returned array (no staging copy)."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return t.numpy()
    except Exception:
        pass
    return np.empty(shape, dtype)


def _out_buffer(buf, shape, dtype, fn, holds_input=False):
    """A caller-supplied result array is used only if the D2H copy can land in it as is (exact shape, dtype,
    C-contiguous); anything else gets a fresh array -- with the caller's values copied in when the buffer is also an
    input (OPTFLOW_USE_INITIAL_FLOW).  A wrong-sized buffer that must hold an input is the caller's error."""
    ok = (isinstance(buf, np.ndarray) and buf.shape == tuple(shape) and buf.dtype == dtype and buf.flags.c_contiguous
          and buf.flags.writeable)
    if ok:
        return buf
    if holds_input:
        _assert(isinstance(buf, np.ndarray) and buf.shape == tuple(shape) and buf.dtype == dtype,
                "flow0.size() == prev0.size() && flow0.type() == CV_32FC2", fn)
    out = _new_host(shape, dtype)
    if holds_input:
        out[...] = buf
    return out


def cvtColor(src, code, dst=None, dstCn=0, hint=0):
    """``cv2.cvtColor(src, code[, dst[, dstCn[, hint]]])``; ``hint`` (cv2.AlgorithmHint) selects between cv2's exact and
    approximate colour paths and does not change BGR2GRAY, so it is accepted and ignored."""
    fn = "cvtColor"
    if code not in (COLOR_BGR2GRAY, COLOR_RGB2GRAY, COLOR_BGRA2GRAY, COLOR_RGBA2GRAY):
        raise error(-213, f"only the *2GRAY codes 6 / 7 / 10 / 11 are supported (the reference uses COLOR_BGR2GRAY); got code {code}")
    _assert(isinstance(src, np.ndarray) and src.size > 0, "!_src.empty()", fn)
    # cv2 takes three or four channels for every one of these codes and ignores the fourth
    _assert(src.dtype == np.uint8 and src.ndim == 3 and src.shape[2] in (3, 4), "(scn == 3 || scn == 4) && depth == CV_8U", fn)
    if code in (COLOR_RGB2GRAY, COLOR_RGBA2GRAY):
        src = np.ascontiguousarray(src[..., 2::-1])      # same coefficients with the colour order swapped
    elif src.shape[2] == 4:
        src = np.ascontiguousarray(src[..., :3])
    if src.strides[2] != 1 or src.strides[1] != 3:
        src = np.ascontiguousarray(src)
    h, w = src.shape[:2]
    if dst is None or not (isinstance(dst, np.ndarray) and dst.shape == (h, w) and dst.dtype == np.uint8
                           and dst.flags.c_contiguous):
        dst = np.empty((h, w), np.uint8)
    _lib.check(_lib.lib().b2of_bgr2gray_u8_host(_ptr(src), h, w, src.strides[0], _ptr(dst), dst.strides[0]))
    return dst


def pyrDown(src, dst=None, dstsize=None, borderType=BORDER_DEFAULT):
    """``cv2.pyrDown(src[, dst[, dstsize[, borderType]]])`` for the default output size and border (what the reference
    and cv2's own LK pyramid use); any other ``dstsize`` / ``borderType`` is refused, not approximated."""
    src = _image_u8(src, "src", "pyrDown")
    h, w = src.shape
    shape = ((h + 1) // 2, (w + 1) // 2)
    if dstsize is not None and tuple(dstsize) not in ((0, 0), (shape[1], shape[0])):
        raise error(-213, f"pyrDown: only the default dstsize {(shape[1], shape[0])} is supported; got {tuple(dstsize)}")
    if borderType != BORDER_DEFAULT:
        raise error(-213, f"pyrDown: only BORDER_DEFAULT (BORDER_REFLECT_101) is supported; got {borderType}")
    if dst is None or not (isinstance(dst, np.ndarray) and dst.shape == shape and dst.dtype == np.uint8
                           and dst.flags.c_contiguous):
        dst = np.empty(shape, np.uint8)
    _lib.check(_lib.lib().b2of_pyrdown_u8_host(_ptr(src), h, w, src.strides[0], _ptr(dst), dst.strides[0]))
    return dst


def calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
    fn = "calcOpticalFlowFarneback"
    prev = _image_u8(prev, "prev0", fn)
    next = _image_u8(next, "next0", fn)
    _assert(prev.shape == next.shape, "prev0.size() == next0.size() && prev0.channels() == next0.channels()", fn)
    _assert(pyr_scale < 1, "pyrScale_ < 1", fn)
    h, w = prev.shape
    if prev.strides[0] != next.strides[0]:
        prev, next = np.ascontiguousarray(prev), np.ascontiguousarray(next)
    if flags & OPTFLOW_USE_INITIAL_FLOW:
        _assert(isinstance(flow, np.ndarray) and flow.shape == (h, w, 2) and flow.dtype == np.float32,
                "flow0.size() == prev0.size() && flow0.type() == CV_32FC2", fn)
    out = flow
    if not (isinstance(out, np.ndarray) and out.shape == (h, w, 2) and out.dtype == np.float32
            and out.flags.c_contiguous):
        out = _new_host((h, w, 2), np.float32)
        if flags & OPTFLOW_USE_INITIAL_FLOW:
            out[...] = flow          # a strided initial estimate (cv2 accepts any Mat step): upload its values
    p = FarnebackParams(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma),
                        int(flags))
    _lib.check(_lib.lib().b2of_farneback_host(_ptr(prev), _ptr(next), prev.strides[0], h, w, C.byref(p), _ptr(out)))
    return out


def calcOpticalFlowFarnebackBatch(prev, next, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                                  poly_sigma=1.2, flags=0, flow=None):
    """Extension: uint8 (B,H,W) prev/next stacks -> float32 (B,H,W,2); copies and compute are pipelined."""
    fn = "calcOpticalFlowFarneback"
    _assert(prev.dtype == np.uint8 and prev.ndim == 3 and prev.shape == next.shape and next.dtype == np.uint8,
            "prev0.size() == next0.size()", fn)
    prev, next = np.ascontiguousarray(prev), np.ascontiguousarray(next)
    b, h, w = prev.shape
    flow = _out_buffer(flow, (b, h, w, 2), np.float32, fn, bool(flags & OPTFLOW_USE_INITIAL_FLOW))
    p = FarnebackParams(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma),
                        int(flags))
    _lib.check(_lib.lib().b2of_farneback_pairs_host(_ptr(prev), _ptr(next), w, h * w, b, h, w, C.byref(p),
                                                    _ptr(flow)))
    return flow


def calcOpticalFlowFarnebackSequence(frames, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                                     poly_sigma=1.2, flags=0, flow=None):
    """Extension for video: uint8 (F,H,W) consecutive gray frames -> float32 (F-1,H,W,2), flow[i] = frame i -> i+1
    (what the reference's per-frame loop computes, DenseOF.py:510-525).  Per-frame work is done once per frame;
    host<->device copies are pipelined with the kernels."""
    fn = "calcOpticalFlowFarneback"
    _assert(isinstance(frames, np.ndarray) and frames.dtype == np.uint8 and frames.ndim == 3,
            "prev0.type() == CV_8UC1", fn)
    frames = np.ascontiguousarray(frames)
    f, h, w = frames.shape
    flow = _out_buffer(flow, (max(f - 1, 0), h, w, 2), np.float32, fn, bool(flags & OPTFLOW_USE_INITIAL_FLOW))
    p = FarnebackParams(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma),
                        int(flags))
    _lib.check(_lib.lib().b2of_farneback_sequence_host(_ptr(frames), w, h * w, f, h, w, C.byref(p), _ptr(flow)))
    return flow


def calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status=None, err=None, winSize=(21, 21), maxLevel=3,
                         criteria=(TERM_CRITERIA_COUNT + TERM_CRITERIA_EPS, 30, 0.01), flags=0,
                         minEigThreshold=1e-4):
    fn = "calcOpticalFlowPyrLK"
    prevImg = _image_u8(prevImg, "prevImg", fn)
    nextImg = _image_u8(nextImg, "nextImg", fn)
    _assert(prevImg.shape == nextImg.shape, "prevPyr[level].size() == nextPyr[level].size()", fn)
    _assert(maxLevel >= 0 and winSize[0] > 2 and winSize[1] > 2, "maxLevel >= 0 && winSize.width > 2 && winSize.height > 2", fn)
    _assert(isinstance(prevPts, np.ndarray) and prevPts.dtype == np.float32 and prevPts.size % 2 == 0
            and (prevPts.ndim >= 1 and prevPts.shape[-1] == 2),
            "(npoints = prevPtsMat.checkVector(2, CV_32F, true)) >= 0", fn)
    shape = prevPts.shape
    pts = np.ascontiguousarray(prevPts.reshape(-1, 2))
    n = len(pts)
    # cv2 writes into caller-supplied nextPts / status / err of the right size and type and returns those objects
    def _reuse(buf, want_shape, dtype):
        return (isinstance(buf, np.ndarray) and buf.dtype == dtype and buf.shape == tuple(want_shape)
                and buf.flags.c_contiguous and buf.flags.writeable)
    if flags & OPTFLOW_USE_INITIAL_FLOW:
        _assert(isinstance(nextPts, np.ndarray) and nextPts.dtype == np.float32 and nextPts.size == pts.size,
                "nextPtsMat.checkVector(2, CV_32F, true) == npoints", fn)
    if _reuse(nextPts, shape, np.float32):
        nxt_ret, nxt = nextPts, nextPts.reshape(-1, 2)
    else:
        nxt = np.empty((n, 2), np.float32)
        if flags & OPTFLOW_USE_INITIAL_FLOW:
            nxt[...] = nextPts.reshape(-1, 2)
        nxt_ret = nxt.reshape(shape)
    st = status if _reuse(status, (n, 1), np.uint8) else np.empty((n, 1), np.uint8)
    er = err if _reuse(err, (n, 1), np.float32) else np.empty((n, 1), np.float32)
    if n == 0:
        return None, None, None          # what cv2's binding hands back for an empty point set
    if prevImg.strides[0] != nextImg.strides[0]:
        prevImg, nextImg = np.ascontiguousarray(prevImg), np.ascontiguousarray(nextImg)
    h, w = prevImg.shape
    p = LKParams(int(winSize[0]), int(winSize[1]), int(maxLevel), int(criteria[0]), int(criteria[1]),
                 float(criteria[2]), int(flags), float(minEigThreshold))
    _lib.check(_lib.lib().b2of_pyrlk_host(_ptr(prevImg), _ptr(nextImg), prevImg.strides[0], h, w, _ptr(pts), n,
                                          _ptr(nxt), _ptr(st), _ptr(er), C.byref(p)))
    return nxt_ret, st, er


def goodFeaturesToTrack(image, maxCorners, qualityLevel, minDistance, corners=None, mask=None, blockSize=3,
                        gradientSize=None, useHarrisDetector=False, k=0.04):
    fn = "goodFeaturesToTrack"
    # cv2 has two overloads; the second inserts gradientSize before useHarrisDetector
    if isinstance(gradientSize, bool):
        useHarrisDetector, gradientSize = gradientSize, None
    if gradientSize is None:
        gradientSize = 3
    _assert(qualityLevel > 0 and minDistance >= 0 and maxCorners >= 0,
            "qualityLevel > 0 && minDistance >= 0 && maxCorners >= 0", fn)
    image = _image_u8(image, "image", fn)
    h, w = image.shape
    mask_step = 0
    if mask is not None:
        _assert(isinstance(mask, np.ndarray) and mask.dtype == np.uint8 and mask.shape == (h, w),
                "_mask.empty() || (_mask.type() == CV_8UC1 && _mask.sameSize(_image))", fn)
        mask = _image_u8(mask, "mask", fn)
        mask_step = mask.strides[0]
    _assert(gradientSize in (3, 5, 7), "ksize == 3 || ksize == 5 || ksize == 7", "Sobel")
    p = GFTTParams(int(maxCorners), float(qualityLevel), float(minDistance), int(blockSize), int(gradientSize),
                   int(bool(useHarrisDetector)), float(k))
    cap = int(maxCorners) if maxCorners > 0 else h * w // 4 + 1
    n_out = C.c_int(0)
    while True:
        buf = np.empty((max(cap, 1), 1, 2), np.float32)
        _lib.check(_lib.lib().b2of_gftt_host(_ptr(image), _ptr(mask) if mask is not None else None, image.strides[0],
                                             mask_step, h, w, C.byref(p), _ptr(buf), cap, C.byref(n_out)))
        if n_out.value <= cap:
            break
        cap = n_out.value
    if n_out.value == 0:
        return None
    res = buf[:n_out.value]
    if (isinstance(corners, np.ndarray) and corners.dtype == np.float32 and corners.shape == res.shape
            and corners.flags.c_contiguous and corners.flags.writeable):
        corners[...] = res           # cv2 reuses a caller-supplied array of exactly the result's size
        return corners
    return res.copy()
