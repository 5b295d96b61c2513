"""Per-region instruction / stall-sample table from an `ncu --page source --csv --print-source cuda,sass` export of a kernel
whose code lives in several files (sfx_phases.cuh, sfx_device.cuh, sfx_kernels.cu).  Regions are the `// ----` / `// ====`
markers of sfx_phases.cuh; lines of other files are reported per file.
usage: python tools/ncu_phase_regions.py src.csv frames [repo_root]"""
import collections
import csv
import os
import re
import sys

path, frames = sys.argv[1], float(sys.argv[2])
root = sys.argv[3] if len(sys.argv) > 3 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
phases = open(os.path.join(root, "multimodal-emotion-classification_b200", "csrc", "sfx_phases.cuh")).read().split("\n")
marks = []
for i, l in enumerate(phases, 1):
    m = re.search(r"^\s*// (?:----|=====+) ?(.*?)(?: =+)?$", l)
    if m and m.group(1).strip() and not set(m.group(1).strip()) <= set("-="):
        marks.append((i, m.group(1).strip()[:44]))


def region(fname, ln):
    if not fname.endswith("sfx_phases.cuh"):
        return os.path.basename(fname)
    name = "phases: prologue"
    for i, n in marks:
        if ln >= i:
            name = n
    return name


rows = list(csv.reader(open(path)))
# an instruction is listed once per source line of its inline call stack: group the rows by SASS address first
insts = collections.OrderedDict()      # address -> [opcode, executed, samples, [(file, line), ...]]
fname, col, cur = "", None, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1]; continue
    if r[0] == "Line No":
        col = {h: i for i, h in enumerate(r)}; continue
    if col is None or len(r) <= col["Instructions Executed"]:
        continue
    if r[0].isdigit():
        cur = int(r[0]); continue
    if not r[2].startswith("0x"):
        continue
    try:
        n = int(float(r[col["Instructions Executed"]])); smp = int(float(r[col["# Samples"]]))
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    e = insts.setdefault(r[2], [m.group(2) if m else "?", n, smp, [], None])
    e[3].append((fname, cur))
    if e[4] is None:
        e[4] = {h[6:]: int(float(r[i] or 0)) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h}
agg = collections.OrderedDict()
ops = collections.defaultdict(collections.Counter)
stalls = collections.defaultdict(collections.Counter)
for op, n, smp, locs, st in insts.values():
    ph = [l for l in locs if l[0].endswith("sfx_phases.cuh")]
    ker = [l for l in locs if l[0].endswith(".cu")]
    f, ln = (max(ph, key=lambda l: l[1]) if ph else ker[0] if ker else locs[0])
    g = region(f, ln)
    a = agg.setdefault(g, [0, 0])
    a[0] += n; a[1] += smp
    ops[g][op] += n
    stalls[g].update(st or {})
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"{'region':46s} {'inst%':>6s} {'samp%':>6s} {'inst/frame':>10s}  top opcodes (per frame)")
for g, a in agg.items():
    top = " ".join(f"{k}:{v / frames:.0f}" for k, v in ops[g].most_common(7))
    print(f"{g:46s} {100 * a[0] / ti:6.1f} {100 * a[1] / ts:6.1f} {a[0] / frames:10.0f}  {top}")
print(f"total inst/frame {ti / frames:.0f}  samples {ts}")
tot = collections.Counter()
for c in ops.values():
    tot.update(c)
print("all: " + " ".join(f"{k}:{v / frames:.0f}" for k, v in tot.most_common(30)))

names = ["selected", "long_sb", "short_sb", "wait", "not_selected", "barrier", "branch_resolving", "dispatch", "no_inst", "math",
         "mio", "lg"]
print()
print(f"{'stall samples by region (% of all)':46s}" + "".join(f"{n[:8]:>9s}" for n in names))
tot_s = sum(sum(c.values()) for c in stalls.values()) or 1
for g, c in stalls.items():
    if sum(c.values()) / tot_s < 0.004:
        continue
    print(f"{g:46s}" + "".join(f"{100 * c[n] / tot_s:9.1f}" for n in names))
allc = collections.Counter()
for c in stalls.values():
    allc.update(c)
print(f"{'all':46s}" + "".join(f"{100 * allc[n] / tot_s:9.1f}" for n in names))
