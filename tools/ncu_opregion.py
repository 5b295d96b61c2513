"""opcode x region table (executed warp-inst per frame) from an ncu source CSV. usage: src.csv kernel.cu frames"""
import csv, re, sys, collections
path, cu, frames = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = open(cu).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"// (?:----|=====+) ?(.*?)(?: =+)?$", l)
    if m and i > 230: marks.append((i, m.group(1).strip()[:22]))
kern_start = next(i for i, l in enumerate(src, 1) if "__global__ void __launch_bounds__" in l)
def region(ln):
    if ln < kern_start: return "helpers<%d" % kern_start
    name = "prologue"
    for i, n in marks:
        if ln >= i: name = n
    return name
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
ci = rows[hi].index("Instructions Executed")
cur = 0
tab = collections.defaultdict(collections.Counter)
for r in rows[hi + 1:]:
    if len(r) <= ci: continue
    if r[0].isdigit(): cur = int(r[0]); continue
    if not r[2].startswith("0x"): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    if not m: continue
    try: n = int(float(r[ci]))
    except ValueError: continue
    tab[region(cur)][m.group(2)] += n
ops = ["FADD","FFMA","FMUL","LDS","STS","SHFL","IADD3","MOV","ISETP","LOP3","FSEL","BRA","BSYNC","LDG","STG","MUFU","IMAD","FSETP","SEL"]
print(f"{'region':24s}" + "".join(f"{o:>6s}" for o in ops) + "   other  total")
for g, c in tab.items():
    tot = sum(c.values())
    if tot / frames < 5: continue
    oth = tot - sum(c[o] for o in ops)
    print(f"{g:24s}" + "".join(f"{c[o]/frames:6.0f}" for o in ops) + f"  {oth/frames:6.0f} {tot/frames:6.0f}")
