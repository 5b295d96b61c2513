"""Cycle accounting of the stream kernel (library built with -DSFX_STREAM_DIAG): where the warps' time goes, per signal kind.
usage: python tools/stream_prof.py [clips]"""
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
names = ["frames", "tails", "waiting", "sched", "n_frames", "n_tails", "t:peaks", "t:median+hist", "t:mfcc", "t:chroma",
         "t:epilogue", "n_peaks"]
ex.set_pipeline("stream")
buf = np.zeros(148 * 16 * 16, dtype=np.int64)


def prof(reset):
    rc = ex.lib.sfx_stream_prof(buf.ctypes.data_as(ctypes.c_void_p), buf.size, reset)
    assert rc == 0
    return buf.reshape(148, 16, 16).copy()


for kind, w in [("mix", pool)] + [(bench.KINDS[k], pool[k::4].contiguous()) for k in range(4)]:
    out = torch.empty((w.shape[0], 56), device=dev)
    ex.extract(w, out=out)
    torch.cuda.synchronize()
    prof(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ex.extract(w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    p = prof(1).sum(axis=(0, 1)).astype(np.float64)
    total = p[0] + p[1] + p[2] + p[3]
    print(f"{kind}: {w.shape[0] / ms / 1e3:.3f} M clips/s; warp time: frames {p[0] / total:.1%} tails {p[1] / total:.1%} "
          f"waiting {p[2] / total:.1%} sched {p[3] / total:.1%}; cycles/frame {p[0] / max(p[4], 1):.0f}; "
          f"cycles/tail {p[1] / max(p[5], 1):.0f} = peaks {p[6] / max(p[5], 1):.0f} + median/hist {p[7] / max(p[5], 1):.0f} + "
          f"mfcc {p[8] / max(p[5], 1):.0f} + chroma {p[9] / max(p[5], 1):.0f} + epilogue {p[10] / max(p[5], 1):.0f}; "
          f"peaks/clip {p[11] / max(p[5], 1):.0f}", flush=True)
ex.set_pipeline("auto")
