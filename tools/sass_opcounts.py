"""Static SASS opcode histogram of the shipped library, per kernel -> profiles/r02_sass_opcounts.txt
usage: python tools/sass_opcounts.py [lib.so] > profiles/r02_sass_opcounts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodal-emotion-classification_b200", "sfx_b200", "libsfx_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "HMMA", "LDSM", "UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UBLKPF",
       "SYNCS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOMS", "ATOMG", "SHFL", "MUFU", "DFMA", "DADD", "DMUL", "F2F", "BAR", "MOV"]
print("""# SASS opcode histogram of the shipped libsfx_b200.so (cuobjdump -sass, sm_100a), per kernel (static instruction counts;
# tools/sass_opcounts.py).  Blackwell-native forms: FFMA2 / FADD2 / FMUL2 (packed FP32), UBLKCP (cp.async.bulk: TMA bulk copy),
# UBLKPF (bulk L2 prefetch), SYNCS (mbarrier), and in sfx_extract_kernel<., true> (pipeline mode 4, fused_umma): UTCHMMA
# (tcgen05.mma), LDTM (tcgen05.ld from tensor memory), UTCBAR (tcgen05.commit), UTCATOMSWS (tensor-memory allocation).  HMMA =
# warp-level mma.sync (the chroma projection of the default pipelines, the fused DNN forward, with LDSM = ldmatrix).  UTMALDG
# (tensor-map TMA) does not appear: operand blocks are stored as contiguous shared-memory images, so plain bulk copies move them.
# ptxas reports 0 bytes of register spills for sfx_extract_kernel<false, false> (the L1 left beside 2 x 104 KB of shared memory is
# too small to absorb any: builds that spilled a few hundred bytes lost a fifth of their throughput); its few LDL / STL are the
# 48-byte stack frame of the rare reference-form calls in the clip tail.
""")
for part in re.split(r"\n\s*Function : ", sass)[1:]:
    name = part.split("\n")[0].strip()
    ops = collections.Counter()
    for line in part.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    n = sum(ops.values())
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    print(f"{dem}  ({n} instructions)")
    print("    " + "  ".join(f"{k}:{ops[k]}" for k in KEY if ops[k]))
    rest = [(k, v) for k, v in ops.most_common() if k not in KEY][:10]
    print("    other: " + "  ".join(f"{k}:{v}" for k, v in rest))
