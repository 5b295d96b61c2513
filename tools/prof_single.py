"""ncu target: the split pipeline on one harmonic 3 s clip (what one request of the reference's Flask app costs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
import bench
from sfx_b200 import get_extractor

dev = torch.device("cuda", 0)
ex = get_extractor(dev)
w = bench.synth_pool(4, 66150, seed=3, device=dev)[1:2].contiguous()      # kind 1 = harmonic
out = torch.empty((1, 56), device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    ex.extract(w, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
