#!/usr/bin/env python
"""Print the hottest SASS instructions of each kernel from `ncu --page source --csv` output."""
import csv
import sys

path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.006
rows = list(csv.reader(open(path)))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for ki, k in enumerate(kern):
    h = k["hdr"]
    si, ex = h.index("# Samples"), h.index("Instructions Executed")
    tot = sum(int(r[si]) for r in k["rows"])
    print(f"[{ki}] {k['name'][:60]} instrs={len(k['rows'])} samples={tot}")
    if which != ki and which != -1:
        continue
    names = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_barrier",
             "stall_not_selected", "stall_math", "stall_branch_resolving", "stall_selected", "stall_no_inst"]
    cols = [h.index(c) for c in names]
    agg = [sum(int(r[c]) for r in k["rows"]) for c in cols]
    print("   stall totals:", {n.replace('stall_', ''): a for n, a in zip(names, agg)})
    for i, r in enumerate(k["rows"]):
        if int(r[si]) > tot * frac:
            st = {n.replace("stall_", ""): int(r[c]) for n, c in zip(names, cols) if int(r[c]) > 0.2 * int(r[si])}
            print(f"   {i:5d} {r[1].strip()[:64]:64s} samp={r[si]:>7s} exec={r[ex]:>9s} {st}")
