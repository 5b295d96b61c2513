"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: python tools/ncu_lines.py src.csv [topN]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
# the export has the CUDA view first (Line No, Source...) -- rows with a numeric line number
agg = defaultdict(lambda: [0, 0, 0, ""])
tot_inst = tot_samp = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    ln = int(r[0])
    def num(v):
        try:
            return int(float(v.split("(")[0]))
        except ValueError:
            return 0
    inst = num(r[col["Instructions Executed"]])
    samp = num(r[col["# Samples"]])
    exc = num(r[col["L1 Wavefronts Shared Excessive"]])
    a = agg[ln]
    a[0] += inst; a[1] += samp; a[2] += exc; a[3] = r[1].strip()[:90]
    tot_inst += inst; tot_samp += samp
print(f"total warp-inst {tot_inst:,}  samples {tot_samp:,}")
print("by samples:")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:5d} inst {100*a[0]/max(tot_inst,1):5.1f}%  samp {100*a[1]/max(tot_samp,1):5.1f}%  smem-excess {a[2]:>10,}  {a[3]}")
