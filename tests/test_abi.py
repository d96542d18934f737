"""The C-ABI library loads on a CPU-only box and exports every symbol include/hole_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from graphembeddings_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hole_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hole_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared_symbols()
    assert "hole_train_steps_host" in names and "hole_rank" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in hole_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_row_stride(lib):
    assert lib.hole_abi_version() == _lib.ABI_VERSION
    assert lib.hole_row_stride(150) == 152      # halves of 75 padded to 76
    assert lib.hole_row_stride(256) == 256
    assert lib.hole_row_stride(128) == 128
    assert lib.hole_row_stride(20) == 24
    assert lib.hole_row_stride(7) < 0           # holE.py:164-165 needs an even dim
    assert lib.hole_row_stride(0) < 0


def test_no_cpu_fallback(lib):
    """Without a GPU the context cannot be created and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.hole_ctx_create(ctypes.byref(h), 0, 100, 16)
    assert rc == -2
    assert b"no CPU fallback" in lib.hole_last_error()
    from graphembeddings_b200.engine import HoleEngine, HoleError
    with pytest.raises(HoleError):
        HoleEngine(100, 16)
