"""ctypes binding of libsfx_b200.so -- one Python function per entry point of include/sfx.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class SfxError(RuntimeError):
    """A libsfx_b200 call returned a negative status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libsfx_b200 error {code}: {msg}")
        self.code = code


class TablesHost(C.Structure):
    _fields_ = [("sr", C.c_int32), ("pip_kmin", C.c_int32), ("pip_kmax", C.c_int32), ("mel_ps", C.c_int32),
                ("mel_flush32", C.c_int32),
                ("hann", C.c_void_p), ("tw1", C.c_void_p), ("tw2", C.c_void_p), ("mel_ab", C.c_void_p),
                ("mel_mask", C.c_void_p), ("mel_src", C.c_void_p),
                ("chroma16", C.c_void_p), ("chroma_ny", C.c_void_p), ("dct", C.c_void_p), ("edges", C.c_void_p),
                ("chroma_frag", C.c_void_p), ("chroma_umma", C.c_void_p)]


class DebugOut(C.Structure):
    _fields_ = [("P", C.c_void_p), ("logmel", C.c_void_p), ("frame_feat", C.c_void_p),
                ("clip_info", C.c_void_p), ("T_dbg", C.c_int32)]


class ResamplerHost(C.Structure):
    _fields_ = [("up", C.c_int32), ("down", C.c_int32), ("n_taps", C.c_int32), ("n_pre_remove", C.c_int32),
                ("taps", C.c_void_p)]


class DnnHost(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_void_p), ("kernel", C.c_void_p), ("bias", C.c_void_p),
                ("bn_gamma", C.c_void_p), ("bn_beta", C.c_void_p), ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p),
                ("bn_eps", C.c_float), ("scaler_mean", C.c_void_p), ("scaler_scale", C.c_void_p)]


EXPORTS = ["sfx_dnn_create", "sfx_dnn_destroy", "sfx_dnn_workspace_bytes", "sfx_dnn_launches_per_forward",
           "sfx_dnn_last_error", "sfx_dnn_forward", "sfx_abi_version", "sfx_last_error", "sfx_device_count", "sfx_init_tables", "sfx_workspace_bytes",
           "sfx_launches_per_extract", "sfx_set_pipeline", "sfx_extract", "sfx_extract_debug", "sfx_extract_host", "sfx_extract_host_pcm16",
           "sfx_preprocess_host_pcm16", "sfx_frontend_last_error", "sfx_frontend_release", "sfx_release",
           "sfx_workspace_bytes_batch"]
BENCH_EXPORTS = ["sfx_measure_fp32_peak"]          # include/sfx_bench.h -> libsfx_bench.so (bench.py / tests only)
PIPELINES = {"auto": 0, "fused": 1, "split": 2, "stream": 3, "fused_umma": 4}
ERR_BAD_CLIP = -5


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is missing and nvcc exists).  Raises if it cannot: no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("SFX_B200_LIB") or _build.LIB_PATH      # SFX_B200_LIB: an A/B build (tools/ab_modes.py, profiling)
    if path == _build.LIB_PATH and not os.path.exists(path):
        _build.build()
    lib = C.CDLL(path)
    lib.sfx_abi_version.restype = C.c_int
    lib.sfx_last_error.restype = C.c_char_p
    lib.sfx_device_count.restype = C.c_int
    lib.sfx_launches_per_extract.restype = C.c_int
    lib.sfx_init_tables.restype = C.c_int
    lib.sfx_init_tables.argtypes = [C.c_int, C.POINTER(TablesHost)]
    lib.sfx_workspace_bytes.restype = C.c_size_t
    lib.sfx_workspace_bytes.argtypes = [C.c_int, C.c_int64]
    common = [C.c_int, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
              C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sfx_extract.restype = C.c_int
    lib.sfx_extract.argtypes = common
    lib.sfx_extract_debug.restype = C.c_int
    lib.sfx_extract_debug.argtypes = common + [C.POINTER(DebugOut)]
    lib.sfx_extract_host.restype = C.c_int
    lib.sfx_extract_host.argtypes = [C.c_int, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_int64, C.c_int32]
    lib.sfx_extract_host_pcm16.restype = C.c_int
    lib.sfx_extract_host_pcm16.argtypes = [C.c_int, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_int64, C.c_int32]
    lib.sfx_preprocess_host_pcm16.restype = C.c_int
    lib.sfx_preprocess_host_pcm16.argtypes = [C.c_int, C.c_int32, C.POINTER(ResamplerHost), C.c_void_p, C.c_int64, C.c_int32,
                                              C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                              C.c_int32]
    lib.sfx_frontend_last_error.restype = C.c_char_p
    lib.sfx_frontend_release.restype = C.c_int
    lib.sfx_frontend_release.argtypes = [C.c_int]
    lib.sfx_release.restype = C.c_int
    lib.sfx_release.argtypes = [C.c_int]
    lib.sfx_workspace_bytes_batch.restype = C.c_size_t
    lib.sfx_workspace_bytes_batch.argtypes = [C.c_int, C.c_int64, C.c_int64]
    lib.sfx_set_pipeline.restype = C.c_int
    lib.sfx_set_pipeline.argtypes = [C.c_int]
    lib.sfx_dnn_create.restype = C.c_int
    lib.sfx_dnn_create.argtypes = [C.c_int, C.POINTER(DnnHost), C.POINTER(C.c_void_p)]
    lib.sfx_dnn_destroy.restype = C.c_int
    lib.sfx_dnn_destroy.argtypes = [C.c_void_p]
    lib.sfx_dnn_workspace_bytes.restype = C.c_size_t
    lib.sfx_dnn_workspace_bytes.argtypes = [C.c_void_p, C.c_int32]
    lib.sfx_dnn_launches_per_forward.restype = C.c_int
    lib.sfx_dnn_launches_per_forward.argtypes = [C.c_void_p]
    lib.sfx_dnn_last_error.restype = C.c_char_p
    lib.sfx_dnn_forward.restype = C.c_int
    lib.sfx_dnn_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
    if lib.sfx_abi_version() != 3:
        raise RuntimeError("libsfx_b200.so ABI version mismatch; rebuild with sfx_b200.build.build(force=True)")
    _LIB = lib
    return lib


_BENCH_LIB = None


def load_bench():
    """libsfx_bench.so: measurement helpers of include/sfx_bench.h.  Only bench.py and the tests load it."""
    global _BENCH_LIB
    if _BENCH_LIB is None:
        if not os.path.exists(_build.BENCH_LIB_PATH):
            _build.build()
        lib = C.CDLL(_build.BENCH_LIB_PATH)
        lib.sfx_measure_fp32_peak.restype = C.c_int
        lib.sfx_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        _BENCH_LIB = lib
    return _BENCH_LIB


def check(rc: int):
    if rc < 0:
        raise SfxError(rc, load().sfx_last_error().decode("utf-8", "replace"))
    return rc


def make_tables_struct(tb: dict):
    """TablesHost pointing into the numpy arrays of tables.build_tables (keeps them alive via .keep)."""
    keep = {k: np.ascontiguousarray(tb[k]) for k in
            ("hann", "tw1", "tw2", "mel_ab", "mel_mask", "mel_src", "chroma16", "chroma_ny", "dct", "edges", "chroma_frag", "chroma_umma")}
    assert keep["chroma_frag"].dtype == np.uint32
    assert keep["hann"].dtype == np.float32 and keep["chroma16"].dtype == np.float16 and keep["chroma_ny"].dtype == np.float32
    assert keep["dct"].dtype == np.float64 and keep["edges"].dtype == np.float64
    assert keep["mel_ab"].dtype == np.float32 and keep["mel_mask"].dtype == np.uint32 and keep["mel_src"].dtype == np.int32
    t = TablesHost(sr=int(tb["sr"]), pip_kmin=int(tb["pip_kmin"]), pip_kmax=int(tb["pip_kmax"]),
                   mel_ps=int(tb["mel_ps"]), mel_flush32=int(tb["mel_flush32"]),
                   **{k: v.ctypes.data for k, v in keep.items()})
    t.keep = keep
    return t
