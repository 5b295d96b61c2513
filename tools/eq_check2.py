import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench
from sfx_b200 import get_extractor
dev = torch.device("cuda", 0)
ex = get_extractor(dev)
B, Be = 65536, 4096
pool = bench.synth_pool(B, 66150, seed=1234, device=dev)
full = ex.extract(pool).cpu().numpy()
sub = ex.extract(pool[:Be].contiguous()).cpu().numpy()
host = ex.extract_host(pool[:Be].cpu().numpy())
for name, v in (("fused 4096 alone", sub), ("host path", host)):
    d = v != full[:Be]
    print(name, "vs rows 0..4095 of the 65536 launch: rows differing", int(d.any(axis=1).sum()), "cols",
          np.nonzero(d.any(axis=0))[0].tolist()[:20], "max abs", float(np.abs(v - full[:Be]).max()))
    if d.any():
        r = np.nonzero(d.any(axis=1))[0]
        print("  first rows", r[:8].tolist(), "kinds", (r[:8] % 4).tolist(), "count by kind", np.bincount(r % 4, minlength=4).tolist())
full2 = ex.extract(pool).cpu().numpy()
print("65536 launch repeated: identical", bool((full2 == full).all()))
