"""Device-resident time per batch for the fused and split pipelines over batch sizes (where should auto mode switch?).
usage: python tools/batch_sweep.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
import bench
from sfx_b200 import get_extractor

dev = torch.device("cuda", 0)
ex = get_extractor(dev)
pool = bench.synth_pool(2048, 66150, seed=11, device=dev)
for B in (128, 200, 256, 296, 300, 400, 512, 592, 600, 768, 1024, 1440, 2048):
    w = pool[:B]
    out = torch.empty((B, 56), device=dev)
    res = []
    for mode in (1, 2):
        ex.lib.sfx_set_pipeline(mode)
        for _ in range(3):
            ex.extract(w, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            ex.extract(w, out=out)
        torch.cuda.synchronize()
        res.append((time.perf_counter() - t0) / 10 * 1e3)
    print(f"B={B:5d}  fused {res[0]:.3f} ms  split {res[1]:.3f} ms  -> {'split' if res[1] < res[0] else 'fused'}")
ex.lib.sfx_set_pipeline(0)
