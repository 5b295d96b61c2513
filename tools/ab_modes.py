"""A/B of library builds x pipeline modes on one box: python tools/ab_modes.py [--clips N] mode:lib.so ... (lib '-' = installed).
Each entry runs in a fresh process; prints per-kind and mix throughput (M clips/s) and a digest of the rows."""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-emotion-classification_b200")
LIB = os.path.join(PKG, "sfx_b200", "libsfx_b200.so")
CHILD = r"""
import sys, hashlib
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
import torch, bench
from sfx_b200 import get_extractor
B = {clips}
dev = torch.device("cuda", 0)
ex = get_extractor(dev)
ex.set_pipeline({mode!r})
pool = bench.synth_pool(B, 66150, seed=7, device=dev)
def rate(w, out, reps):
    for _ in range(2): ex.extract(w, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): ex.extract(w, out=out)
    b.record(); torch.cuda.synchronize()
    return w.shape[0] * reps / (a.elapsed_time(b) * 1e-3)
kinds = []
for k in range(4):
    w = pool[k::4].contiguous()
    kinds.append(rate(w, torch.empty((w.shape[0], 56), device=dev), 4) / 1e6)
out = torch.empty((B, 56), device=dev)
total = rate(pool, out, 6)
digest = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12]
print(" ".join(f"{{v:.3f}}" for v in kinds), f"| mix {{total/1e6:.3f}}", digest)
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=1)
    ap.add_argument("--clips", type=int, default=32768)
    ap.add_argument("entries", nargs="+")
    a = ap.parse_args()
    backup = LIB + ".ab_backup"
    shutil.copy2(LIB, backup)
    try:
        for _ in range(a.rounds):
            for e in a.entries:
                mode, lib = e.split(":", 1)
                shutil.copy2(backup if lib == "-" else lib, LIB)
                res = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, pkg=PKG, clips=a.clips, mode=mode)],
                                     capture_output=True, text=True, timeout=300)
                tag = f"{mode}:{os.path.basename(lib)}"
                if res.returncode != 0:
                    print(tag, "FAILED", res.stderr[-400:], flush=True)
                    continue
                print(f"{tag:28s}", res.stdout.strip().splitlines()[-1], flush=True)
    finally:
        shutil.move(backup, LIB)


if __name__ == "__main__":
    main()
