"""ctypes binding of libb2of.so (C-ABI declared in include/b2of.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a only).  There is no CPU
fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# B2OF_LIB: developer override used by scripts/ab_variants.py to A/B alternative builds of the same sources
LIB_PATH = os.environ.get("B2OF_LIB") or os.path.join(CSRC, "libb2of.so")
SOURCES = ["api.cu", "gray_pyr.cu", "farneback.cu", "pyrlk.cu", "gftt.cu", "pathfinder.cu", "overlay.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class error(Exception):
    """Raised where the reference's user would have seen ``cv2.error``."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code
        self.msg = msg


class FarnebackParams(C.Structure):
    _fields_ = [("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iterations", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("flags", C.c_int)]


class LKParams(C.Structure):
    _fields_ = [("win_w", C.c_int), ("win_h", C.c_int), ("max_level", C.c_int), ("crit_type", C.c_int),
                ("crit_max_count", C.c_int), ("crit_eps", C.c_double), ("flags", C.c_int),
                ("min_eig_threshold", C.c_double)]


class GFTTParams(C.Structure):
    _fields_ = [("max_corners", C.c_int), ("quality_level", C.c_double), ("min_distance", C.c_double),
                ("block_size", C.c_int), ("gradient_size", C.c_int), ("use_harris", C.c_int), ("k", C.c_double)]


_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int
_PF, _PL, _PG = C.POINTER(FarnebackParams), C.POINTER(LKParams), C.POINTER(GFTTParams)

# name -> (restype, argtypes); every symbol include/b2of.h declares
SIGNATURES = {
    "b2of_version": (_i, []),
    "b2of_last_error": (C.c_char_p, []),
    "b2of_release": (_i, []),
    "b2of_launch_count": (C.c_ulonglong, []),
    "b2of_profile_enable": (None, [_i]),
    "b2of_profile_reset": (None, []),
    "b2of_profile_tag_count": (_i, []),
    "b2of_profile_tag_name": (C.c_char_p, [_i]),
    "b2of_profile_read": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), C.POINTER(C.c_double)]),
    "b2of_bgr2gray_u8_dev": (_i, [_vp, _i, _i, _sz, _sz, _vp, _sz, _sz, _i, _vp]),
    "b2of_bgr2gray_u8_host": (_i, [_vp, _i, _i, _sz, _vp, _sz]),
    "b2of_pyrdown_u8_dev": (_i, [_vp, _i, _i, _sz, _sz, _vp, _sz, _sz, _i, _vp]),
    "b2of_pyrdown_u8_host": (_i, [_vp, _i, _i, _sz, _vp, _sz]),
    "b2of_farneback_workspace_bytes": (_sz, [_i, _i, _PF, _i, _i]),
    "b2of_farneback_pairs_dev": (_i, [_vp, _vp, _sz, _sz, _i, _i, _i, _PF, _vp, _vp, _sz, _vp]),
    "b2of_farneback_sequence_dev": (_i, [_vp, _sz, _sz, _i, _i, _i, _PF, _vp, _vp, _sz, _vp]),
    "b2of_farneback_sequence_stats_dev": (_i, [_vp, _sz, _sz, _i, _i, _i, _PF, _vp, _vp, _vp, _sz, _vp]),
    "b2of_farneback_host": (_i, [_vp, _vp, _sz, _i, _i, _PF, _vp]),
    "b2of_farneback_pairs_host": (_i, [_vp, _vp, _sz, _sz, _i, _i, _i, _PF, _vp]),
    "b2of_farneback_sequence_host": (_i, [_vp, _sz, _sz, _i, _i, _i, _PF, _vp]),
    "b2of_pyrlk_workspace_bytes": (_sz, [_i, _i, _PL, _i]),
    "b2of_pyrlk_dev": (_i, [_vp, _vp, _sz, _sz, _i, _i, _i, _vp, _sz, _i, _vp, _vp, _vp, _PL, _vp, _sz, _vp]),
    "b2of_pyrlk_host": (_i, [_vp, _vp, _sz, _i, _i, _vp, _i, _vp, _vp, _vp, _PL]),
    "b2of_gftt_workspace_bytes": (_sz, [_i, _i, _PG, _i]),
    "b2of_gftt_dev": (_i, [_vp, _vp, _sz, _sz, _i, _i, _i, _PG, _vp, _i, _vp, _vp, _sz, _vp]),
    "b2of_gftt_host": (_i, [_vp, _vp, _sz, _sz, _i, _i, _PG, _vp, _i, _vp]),
    "b2of_pathfinder_filter_dev": (_i, [_vp, _sz, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp]),
    "b2of_overlay_vectors_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "b2of_overlay_lamps_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "b2of_flow_sample_dev": (_i, [_vp, _i, _i, _i, _vp, _sz, _i, _vp, _vp]),
    "b2of_flow_hsv_dev": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "b2of_flow_stats_dev": (_i, [_vp, _i, _i, _i, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into csrc/libb2of.so for sm_100a (cross-compiles without a GPU).
    `defines` / `out` build a variant (extra -D macros) to another path without touching the default library."""
    if out is not None:
        cmd = ["nvcc"] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return out
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("common.cuh", "fb_ws.cuh")] + [os.path.join(_HERE, "..", "include", "b2of.h")]
    if not force and os.path.exists(LIB_PATH):
        mt = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= mt for d in deps):
            return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU fallback)")
            l = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(l, name)
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().b2of_last_error().decode("utf-8", "replace")
        raise error(rc, msg)


def profile(enable=None, reset=False):
    """Per-kernel device timing.  ``profile(True)`` / ``profile(False)`` switch it; ``profile()`` returns
    {kernel family: {"ms": total device ms, "launches": n, "bytes": algorithmic bytes}} since the last reset."""
    l = lib()
    if enable is not None:
        l.b2of_profile_enable(1 if enable else 0)
    out = {}
    if enable is None:
        for tag in range(l.b2of_profile_tag_count()):
            ms, n, by = C.c_double(0), C.c_ulonglong(0), C.c_double(0)
            check(l.b2of_profile_read(tag, C.byref(ms), C.byref(n), C.byref(by)))
            if n.value:
                out[l.b2of_profile_tag_name(tag).decode()] = {"ms": ms.value, "launches": n.value, "bytes": by.value}
    if reset:
        l.b2of_profile_reset()
    return out
