"""The C-ABI library loads on a CPU-only box and exports every function include/k2b_b200.h declares
(no compute calls here)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "k2b_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(k2b_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from keypoints2body_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        pytest.skip("library not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) == set(_native.EXPORTS)
    lib.k2b_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.k2b_version()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under keypoints2body_b200/ may import it."""
    pkg = os.path.join(ROOT, "keypoints2body_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(base, f)


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np

    import keypoints2body_b200 as k2b
    from keypoints2body_b200 import synthetic as syn

    with pytest.raises(RuntimeError, match="CUDA"):
        k2b.optimize_params_frame(np.zeros((22, 3), np.float32), model=syn.make_body_model("smpl"))
