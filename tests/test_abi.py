"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b2of.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "b2of.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2of_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from hackathonopticalflow_b200 import _lib
    path = _lib.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b2of.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_error_channel_without_gpu():
    from hackathonopticalflow_b200 import _lib
    l = _lib.lib()
    assert l.b2of_version() == 102
    # argument checks fire before any CUDA call and mirror cv2's assertion text
    rc = l.b2of_bgr2gray_u8_dev(None, 4, 4, 2, 0, None, 4, 0, 1, None)
    assert rc == -215
    assert b"Assertion failed" in l.b2of_last_error()


def test_no_oracle_or_cv2_in_product_path():
    """The shipped package must not import oracle/ or cv2 (no CPU fallback)."""
    pkg = os.path.join(ROOT, "hackathonopticalflow_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), fn
            assert not re.search(r"^\s*(import|from)\s+cv2", src, flags=re.M), fn
