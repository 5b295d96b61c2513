"""Front-end probe: 48 kHz (or argv[2]) 16-bit PCM rows through sfx_preprocess_host_pcm16; prints end-to-end clips/s.
usage: python tools/frontend_probe.py [B] [native_sr] [channels]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
from sfx_b200 import get_extractor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
native = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
ch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ex = get_extractor(torch.device("cuda", 0))
n = native * 3 * ch
pcm = torch.empty((B, n), dtype=torch.int16).pin_memory()
pcm.copy_((torch.randn((B, n), generator=torch.Generator().manual_seed(1)) * 3000).round().clamp(-32768, 32767).to(torch.int16))
out = torch.empty((B, 56), dtype=torch.float32).pin_memory()
for _ in range(2):
    ex.preprocess_pcm16(pcm.numpy(), None, native, channels=ch, out=out.numpy())
t0 = time.perf_counter()
for _ in range(5):
    ex.preprocess_pcm16(pcm.numpy(), None, native, channels=ch, out=out.numpy())
dt = (time.perf_counter() - t0) / 5
print(f"B={B} native={native} ch={ch}: {B / dt / 1e3:.1f} k clips/s end to end, {B * n * 2 / dt / 1e9:.1f} GB/s over PCIe")
