"""The C-ABI shared library: loads, exports every symbol include/sppark_b200.h declares, and
fails loudly (never silently falls back) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "sppark_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(\w+)\s*\([^;{]*\)\s*;", h)))


def test_library_exports_every_declared_symbol(lib):
    names = _declared_symbols()
    assert {"mult_pippenger", "mult_pippenger_inf", "mult_pippenger_fp2_inf", "compute_ntt", "cuda_available",
            "drop_error_message"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sppark_b200.h but not exported"
    from sppark_b200 import _lib
    assert set(_lib.EXPORTS) <= set(names)


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    assert lib.cuda_available() == 0
    buf = np.arange(8, dtype=np.uint64)
    before = buf.copy()
    err = lib.compute_ntt(0, buf.ctypes.data, 3, 0, 0, 0)
    assert err.code != 0 and np.array_equal(buf, before)
    if err.message:
        msg = C.cast(err.message, C.c_char_p).value
        assert msg
        lib.drop_error_message(err.message)
    out = np.ones(18, dtype=np.uint64)
    pts = np.zeros((4, 12), dtype=np.uint64)
    sc = np.zeros((4, 4), dtype=np.uint64)
    err = lib.mult_pippenger(out.ctypes.data, pts.ctypes.data, 4, sc.ctypes.data)
    assert err.code != 0
    if err.message:
        lib.drop_error_message(err.message)


def test_python_wrappers_raise(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only host")
    from sppark_b200 import _lib, ntt
    with pytest.raises(_lib.SpparkError):
        ntt.NTT(0, np.arange(8, dtype=np.uint64))
    with pytest.raises(ValueError):
        ntt.NTT(0, np.arange(6, dtype=np.uint64))       # "inout.len() is not power of 2"


def test_rusterror_layout():
    from sppark_b200 import _lib
    assert C.sizeof(_lib.RustError) == 16 and _lib.RustError.message.offset == 8
