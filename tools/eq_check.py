import os, sys
import numpy as np, torch
ROOT="/root/repo"
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench
from sfx_b200 import get_extractor
dev=torch.device("cuda",0)
ex=get_extractor(dev)
B=4096
pool=bench.synth_pool(B, 66150, seed=7, device=dev)
outs={}
for mode in ("fused","split","fused_umma","stream"):
    ex.set_pipeline(mode)
    outs[mode]=ex.extract(pool).cpu().numpy()
ex.set_pipeline("auto")
h=pool.cpu().numpy()
outs["host"]=ex.extract_host(h)
ref=outs["fused"]
for k,v in outs.items():
    d=(v!=ref)
    print(k, "rows differing", int(d.any(axis=1).sum()), "cols", np.nonzero(d.any(axis=0))[0].tolist()[:60], "max abs", float(np.abs(v-ref).max()))
    if d.any():
        r=np.nonzero(d.any(axis=1))[0][:5]
        print("  rows", r.tolist(), "kinds", (r%4).tolist())
