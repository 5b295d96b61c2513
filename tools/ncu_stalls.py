"""stall-reason samples by kernel region from an ncu source CSV. usage: src.csv kernel.cu"""
import csv, re, sys, collections
path, cu = sys.argv[1], sys.argv[2]
src = open(cu).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"// (?:----|=====+) ?(.*?)(?: =+)?$", l)
    if m and i > 230: marks.append((i, m.group(1).strip()[:26]))
kern_start = next(i for i, l in enumerate(src, 1) if "__global__ void __launch_bounds__" in l)
def region(ln):
    if ln < kern_start: return "helpers"
    name = "prologue"
    for i, n in marks:
        if ln >= i: name = n
    return name
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
cols = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tab = collections.defaultdict(collections.Counter)
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit(): continue
    g = region(int(r[0]))
    for h, i in cols:
        try: tab[g][h] += int(float(r[i]))
        except ValueError: pass
names = ["selected","long_sb","short_sb","wait","not_selected","barrier","branch_resolving","dispatch","no_inst","math","mio","lg"]
tot = sum(sum(c.values()) for c in tab.values())
print(f"{'region':28s}" + "".join(f"{n[:8]:>9s}" for n in names) + "    total%")
for g, c in tab.items():
    t = sum(c.values())
    if t < 0.004 * tot: continue
    print(f"{g:28s}" + "".join(f"{100*c[n]/tot:9.1f}" for n in names) + f"  {100*t/tot:8.1f}")
allc = collections.Counter()
for c in tab.values(): allc.update(c)
print(f"{'ALL':28s}" + "".join(f"{100*allc[n]/tot:9.1f}" for n in names))
