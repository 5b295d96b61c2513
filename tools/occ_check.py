import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
from sfx_b200 import get_extractor
ex = get_extractor(torch.device("cuda", 0))
ex.set_pipeline("fused")
a = ex.lib.sfx_workspace_bytes_batch(0, 66150, 1)
b = ex.lib.sfx_workspace_bytes_batch(0, 66150, 0)
hdr = 256 + 4 * 65536
print("ws(1 clip)", a, "ws(any)", b, "=> resident CTAs", (b - hdr) / (a - hdr))
