#!/usr/bin/env python
"""Region-level (150-instruction windows) sample/instruction breakdown of one kernel from `ncu --page source --csv`."""
import csv, sys
path, ki = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kern.append(cur)
    elif r and r[0] == "Address": cur["hdr"] = r
    elif cur is not None and r: cur["rows"].append(r)
k = kern[ki]; h = k["hdr"]; si = h.index("# Samples"); ex = h.index("Instructions Executed")
names = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_not_selected", "stall_selected",
         "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_mio", "stall_dispatch", "stall_lg", "stall_membar", "stall_sleep"]
cols = [h.index(c) for c in names if c in h]; names = [c for c in names if c in h]
R = k["rows"]; tot = sum(int(r[si]) for r in R); totex = sum(int(r[ex]) for r in R)
print(k["name"][:50], "samples", tot, "exec %.1fM" % (totex / 1e6))
print("  totals", {n[6:]: sum(int(r[c]) for r in R) for n, c in zip(names, cols)})
for a in range(0, len(R), 150):
    seg = R[a:a + 150]; s = sum(int(r[si]) for r in seg); exs = sum(int(r[ex]) for r in seg)
    if s < tot * 0.012 and exs < totex * 0.012: continue
    st = {n[6:]: sum(int(r[c]) for r in seg) for n, c in zip(names, cols)}
    st = {a2: v for a2, v in st.items() if v > 0.08 * s}
    ops = {}
    for r in seg:
        t = r[1].strip().split(); op = t[0] if not t[0].startswith("@") else t[1]
        ops[op] = ops.get(op, 0) + int(r[ex])
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print(f"  [{a:5d}) samp={s:6d} ({100*s/tot:4.1f}%) exec={exs/1e6:7.1f}M ({100*exs/totex:4.1f}%) {st} top={top}")
