"""The C-ABI library: builds in-tree for sm_100a, loads, exports every symbol include/sfx.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from sfx_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name="sfx.h"):
    txt = open(os.path.join(ROOT, "include", name)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sfx_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert getattr(lib, s) is not None
    # the measurement helpers live in their own library, not in the product
    bench = ctypes.CDLL(build.BENCH_LIB_PATH)
    bsyms = header_symbols("sfx_bench.h")
    assert set(bsyms) == set(_lib.BENCH_EXPORTS)
    for s in bsyms:
        assert getattr(bench, s) is not None
        assert not hasattr(lib, s)


def test_abi_version_and_argument_validation():
    lib = _lib.load()
    assert lib.sfx_abi_version() == 3
    assert lib.sfx_launches_per_extract() == 1
    # argument validation happens before any CUDA work
    rc = lib.sfx_extract(99, 22050, None, 0, None, 0, 0, 1, 40, None, 56, None, 0, None)
    assert rc == -1 and b"device" in lib.sfx_last_error()
    rc = lib.sfx_extract(0, 22050, None, 0, None, 0, 0, 1, 400, None, 56, None, 0, None)
    assert rc == -1
    assert lib.sfx_workspace_bytes(0, 66150) == 0 or torch.cuda.is_available()
    assert lib.sfx_workspace_bytes_batch(0, 66150, 1) == 0 or torch.cuda.is_available()
    assert lib.sfx_set_pipeline(5) == -1 and lib.sfx_set_pipeline(-1) == -1
    for mode in (4, 3, 2, 1, 0):
        assert lib.sfx_set_pipeline(mode) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from sfx_b200 import NoCudaDeviceError, get_extractor
    lib = _lib.load()
    assert lib.sfx_device_count() <= 0
    with pytest.raises(NoCudaDeviceError):
        get_extractor()
    x = np.zeros((1, 66150), dtype=np.float32)
    out = np.zeros((1, 56), dtype=np.float32)
    rc = lib.sfx_extract_host(0, 22050, x.ctypes.data, 66150, None, 66150, 1, 40, out.ctypes.data, 56, 0)
    assert rc == -3                                   # SFX_ERR_NOT_INIT: nothing ran, nothing was computed
    pcm = np.zeros((1, 66150), dtype=np.int16)
    rc = lib.sfx_extract_host_pcm16(0, 22050, pcm.ctypes.data, 66150, None, 66150, 1, 40, out.ctypes.data, 56, 0)
    assert rc == -3 and not out.any()
    tf = ctypes.c_double(-1.0)
    assert _lib.load_bench().sfx_measure_fp32_peak(0, ctypes.byref(tf)) == -2 and tf.value == 0.0   # SFX_ERR_CUDA
    from preprocessing.audio_preprocessing import extract_mfcc
    with pytest.raises(NoCudaDeviceError):
        extract_mfcc(np.zeros(66150, dtype=np.float32), 22050)
