#!/usr/bin/env python
"""Per-kernel instruction census of libgvc.so from `cuobjdump -sass` (no GPU needed): what the hot kernels
are made of -- 128-bit gathers, FP32 pipes, shared-memory traffic, spills, tensor-core / bulk-copy
instructions if any.  usage: python tools/sass_summary.py [lib] > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gnn-mwvc_b200" / "libgvc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
BUCKETS = OrderedDict([
    ("LDG.128 (row gathers)", r"^LDG\.E\.128"), ("LDG other", r"^LDG(?!\.E\.128)"), ("LD generic", r"^LD\."),
    ("STG/ST", r"^(STG|ST\.)"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDL (spill load)", r"^LDL"), ("STL (spill store)", r"^STL"),
    ("FFMA", r"^FFMA(?!2)"), ("FFMA2", r"^FFMA2"), ("FMUL", r"^FMUL"), ("FADD", r"^FADD"), ("FADD2/FMUL2", r"^(FADD2|FMUL2)"),
    ("DFMA/DADD/DMUL", r"^D(FMA|ADD|MUL)"), ("HMMA/mma.sync", r"^HMMA"), ("UTC*MMA (tcgen05)", r"^UTC.*MMA"),
    ("UTMALDG/UBLKCP (TMA)", r"^(UTMALDG|UTMASTG|UBLKCP)"), ("LDGSTS (cp.async)", r"^LDGSTS"), ("SHFL", r"^SHFL"),
    ("VOTE", r"^VOTE"), ("BAR", r"^BAR"), ("ATOM/RED", r"^(ATOM|RED)"), ("MUFU", r"^MUFU"), ("NANOSLEEP", r"^NANOSLEEP"),
])
cur, counts, total = None, {}, Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for name, pat in BUCKETS.items():
            if re.match(pat, op):
                counts[cur][name] += 1
                break


def short(name):
    m = re.search(r"stage_kernelILi(\d)ELb(\d)", name)
    if m:
        return f"stage_kernel<{m.group(1)},{'exact' if m.group(2) == '1' else 'fast'}>"
    m = re.search(r"\d+([a-z_0-9]+_kernel)", name)
    return m.group(1) if m else name[:40]


print(f"# SASS census of {Path(lib).name} (cuobjdump -sass; static instruction counts per kernel, device functions included)")
print("# rows: instruction class; columns: kernels.  0 in 'UTC*MMA' / 'UTMALDG' = no tcgen05 / TMA instruction in that kernel.")
names = [k for k in counts if "stage_kernel" in k] + [k for k in counts if "stage_kernel" not in k and total[k] > 150]
print(f"{'':28s}" + "".join(f"{short(k)[:22]:>24s}" for k in names))
print(f"{'instructions':28s}" + "".join(f"{total[k]:24d}" for k in names))
for b in BUCKETS:
    if any(counts[k][b] for k in names):
        print(f"{b:28s}" + "".join(f"{counts[k][b]:24d}" for k in names))
