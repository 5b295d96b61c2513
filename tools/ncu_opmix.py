"""Dynamic SASS opcode mix (executed warp-instructions per frame) from an ncu source-page CSV export.
usage: python tools/ncu_opmix.py src.csv frames [line_lo line_hi]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2])
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi_ = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = hdr.index("Instructions Executed")
cur = 0
c = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= ci: continue
    if r[0].isdigit():
        cur = int(r[0]); continue
    if not r[2].startswith("0x"): continue
    if not (lo <= cur <= hi_): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    if not m: continue
    try: n = int(float(r[ci]))
    except ValueError: continue
    c[m.group(2)] += n
tot = sum(c.values())
print(f"lines [{lo},{hi_}] total {tot/frames:.0f} inst/frame")
print("  ".join(f"{k}:{v/frames:.0f}" for k, v in c.most_common(32)))
