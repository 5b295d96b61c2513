"""In-tree build of libsfx_b200.so (nvcc, sm_100a only).  The .so sits next to this file so that it
travels with the repo snapshot; there is no JIT cache and no other architecture."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB_PATH = os.path.join(HERE, "libsfx_b200.so")
BENCH_LIB_PATH = os.path.join(HERE, "libsfx_bench.so")      # measurement helpers (include/sfx_bench.h), not the product
SOURCES = ["sfx_kernels.cu", "sfx_stream.cu", "sfx_split.cu", "sfx_abi.cu", "sfx_dnn.cu", "sfx_frontend.cu"]
BENCH_SOURCES = ["sfx_peak.cu"]
HEADERS = ["sfx_internal.h", "sfx_device.cuh", "sfx_phases.cuh", os.path.join("..", "..", "include", "sfx.h"),
           os.path.join("..", "..", "include", "sfx_bench.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _stale(lib: str, sources) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in list(sources) + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(lib: str, sources, verbose: bool, extra=()) -> None:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError(f"nvcc not found: cannot build {os.path.basename(lib)} (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", lib] + [os.path.join(CSRC, s) for s in sources]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)


def build(force: bool = False, verbose: bool = False, extra=(), out: str | None = None) -> str:
    """Compile csrc/*.cu into libsfx_b200.so (and the bench helpers into libsfx_bench.so) if missing or older than their
    sources.  `extra` = additional nvcc flags (e.g. -DSFX_STREAM_SLOTS=3 for an A/B build written to `out`)."""
    lib = out or LIB_PATH
    if force or out or _stale(lib, SOURCES):
        _compile(lib, SOURCES, verbose, extra)
    if not out and (force or _stale(BENCH_LIB_PATH, BENCH_SOURCES)):
        _compile(BENCH_LIB_PATH, BENCH_SOURCES, False)
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose=True))
