"""Bucket an ncu source-page CSV export into kernel regions (line ranges read from the '// ----' markers).
usage: python tools/ncu_regions.py src.csv kernel.cu [frames]"""
import csv
import re
import sys
from collections import OrderedDict

path, cu = sys.argv[1], sys.argv[2]
frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
src = open(cu).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"// (?:----|=====+) ?(.*?)(?: =+)?$", l)
    if m and i > 200:
        marks.append((i, m.group(1).strip()[:40]))
def region(ln):
    if ln < 141: return "fft32/helpers (inlined)"
    if ln < 199: return "radix_select"
    if ln < 230: return "reduce_scatter"
    name = "prologue"
    for i, n in marks:
        if ln >= i: name = n
    return name
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
col = {h: i for i, h in enumerate(rows[hi])}
def num(v):
    try: return int(float(v.split("(")[0]))
    except ValueError: return 0
agg = OrderedDict()
ti = ts = 0
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]) or not r[0].isdigit(): continue
    if "sfx_kernels.cu" not in path and False: pass
    g = region(int(r[0]))
    a = agg.setdefault(g, [0, 0, 0, 0])
    a[0] += num(r[col["Instructions Executed"]]); a[1] += num(r[col["# Samples"]])
    a[2] += num(r[col["L1 Wavefronts Shared"]]); a[3] += num(r[col["L1 Wavefronts Shared Excessive"]])
    ti += num(r[col["Instructions Executed"]]); ts += num(r[col["# Samples"]])
print(f"{'region':42s} {'inst%':>6s} {'samp%':>6s} {'inst/frame':>10s} {'smem wf/frame':>13s} {'excess':>8s}")
for g, a in agg.items():
    per = f"{a[0]/frames:10.0f} {a[2]/frames:13.0f} {a[3]/frames:8.0f}" if frames else ""
    print(f"{g:42s} {100*a[0]/ti:6.1f} {100*a[1]/ts:6.1f} {per}")
print("total inst", ti, "samples", ts, (f"inst/frame {ti/frames:.0f}" if frames else ""))
