"""Cycle accounting of the fused kernel (library built with -DSFX_FUSED_DIAG): cycles per clip per CTA in each phase, per
signal kind.  usage: SFX_B200_LIB=ab/libfdiag.so python tools/fused_prof.py [clips]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from sfx_b200 import get_extractor  # noqa: E402

dev = torch.device("cuda", 0)
ex = get_extractor(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
pool = bench.synth_pool(B, 66150, seed=7, device=dev)
ex.set_pipeline(os.environ.get("SFX_PROF_MODE", "fused"))
buf = np.zeros(16, dtype=np.uint64)
names = ["frames", "per-peak", "select+hist", "mfcc", "bank wait", "chroma", "tail total", "clips"]
for kind, w in [("mix", pool)] + [(bench.KINDS[k], pool[k::4].contiguous()) for k in range(4)]:
    out = torch.empty((w.shape[0], 56), device=dev)
    ex.extract(w, out=out)
    torch.cuda.synchronize()
    ex.lib.sfx_fused_prof(buf.ctypes.data_as(ctypes.c_void_p), 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ex.extract(w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ex.lib.sfx_fused_prof(buf.ctypes.data_as(ctypes.c_void_p), 1)
    n = float(buf[7])
    print(f"{kind}: {w.shape[0] / e0.elapsed_time(e1) / 1e3:.3f} M clips/s; cycles per clip per CTA: " +
          ", ".join(f"{nm} {float(buf[i]) / n:.0f}" for i, nm in enumerate(names[:7])) +
          f"; inside select+hist: redo {float(buf[8]) / n:.0f}, radix select {float(buf[9]) / n:.0f}, upper median "
          f"{float(buf[10]) / n:.0f}, (select+hist above = histogram + argmax only); inside mfcc: fence+barrier+bank copy issue "
          f"{float(buf[11]) / n:.0f}, minimum + barrier {float(buf[12]) / n:.0f}, pooling + barrier {float(buf[13]) / n:.0f} (mfcc above = the DCT only)", flush=True)
ex.set_pipeline("auto")
