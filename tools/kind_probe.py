"""Device-resident throughput per signal kind of the bench mix (noise, harmonic, noise_tail, harmonic_tail).
usage: python tools/kind_probe.py [B]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
import bench
from sfx_b200 import get_extractor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
ex = get_extractor(dev)
pool = bench.synth_pool(B, 66150, seed=7, device=dev)
out = torch.empty((B // 4, 56), device=dev)
for k, name in enumerate(bench.KINDS):
    w = pool[k::4].contiguous()
    for _ in range(2):
        ex.extract(w, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ex.extract(w, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    dbg = {}
    ex.extract(w[:64], debug=dbg)
    ci = dbg["clip_info"].cpu().numpy()
    print(f"{name:14s} {w.shape[0] / dt / 1e6:.3f} M clips/s   peaks/clip {ci[:, 2].mean():.0f}")
