"""Short single-GPU run for ncu on the bench mix: python tools/prof_mix.py <pipeline> [clips] [iters]."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from sfx_b200 import get_extractor  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
ex = get_extractor(dev)
ex.set_pipeline(mode)
w = bench.synth_pool(B, 66150, seed=7, device=dev)
out = torch.empty((B, 56), device=dev)
for _ in range(iters):
    ex.extract(w, out=out)
torch.cuda.synchronize()
print("ok", float(out[:, 52:].sum()))
