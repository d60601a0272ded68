#!/usr/bin/env python
"""BASELINE config 5 / SURVEY 8(d): a full GNN_VC run on an ER graph (N vertices, 5N edges), where
predict() is called 10-14 times on a shrinking graph.  Runs the CPU reference (oracle/_ref/GNN_VC_ref,
OpenBLAS pinned to Prescott, one thread: the order the exact mode reproduces) and the drop-in binary
(reference src/GNN_VC.cpp + gnn-mwvc_b200/host + libgvc) with time = 0 (deterministic), compares the
result files byte for byte and reports the predict() time of both.
usage: python tools/replay_config5.py [n_vertices=1000000] [only=<arm>,<arm>...]"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200  # noqa: E402,F401
from gnn_mwvc_b200 import graphs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
ref_bin = ROOT / "oracle" / "_ref" / "GNN_VC_ref"
our_bin = ROOT / "gnn-mwvc_b200" / "host" / "_build" / "GNN_VC"
with tempfile.TemporaryDirectory() as td:
    t = time.time()
    g = graphs.er_graph(n, 5 * n, seed=1)
    gp = Path(td) / "er.graph"
    graphs.write_metis(g, gp)
    print(f"graph n={g.n} E={g.n_edges} written in {time.time() - t:.1f} s", flush=True)
    out = {}
    extra = tuple((kv.split("=")[0], ROOT / kv.split("=")[1], {"GVC_PROFILE": "1", "GVC_MODE": "exact"})
                  for kv in os.environ.get("REPLAY_EXTRA", "").split(",") if kv)     # name=path of another drop-in build
    for name, exe, env in extra + (("reference_cpu", ref_bin, {"OPENBLAS_CORETYPE": "Prescott", "OPENBLAS_NUM_THREADS": "1"}),
                           ("reference_cpu_all_threads", ref_bin, {"OPENBLAS_CORETYPE": "Prescott"}),
                           ("b200_exact", our_bin, {"GVC_PROFILE": "1", "GVC_MODE": "exact"}),
                           ("b200_exact_packed", our_bin, {"GVC_PROFILE": "1", "GVC_MODE": "exact", "GVC_UPLOAD": "packed"}),
                           ("b200_fast", our_bin, {"GVC_PROFILE": "1", "GVC_MODE": "fast"})):
        if only and name not in only:
            continue
        res = Path(td) / f"{name}.out"
        t = time.time()
        r = subprocess.run([str(exe), str(gp), str(res), "0", "-1", "0"], capture_output=True, text=True,
                           env=dict(os.environ, **env))
        wall = time.time() - t
        prof = [l for l in r.stderr.splitlines() if l.startswith("gvc profile:")]
        total = prof[-1] if prof else ""
        m = re.search(r"(\d+) predict calls.*total ([\d.]+) s", total)
        out[name] = {"wall_s": round(wall, 2), "stdout": r.stdout.strip(), "md5": hashlib.md5(res.read_bytes()).hexdigest(),
                     "predict_calls": int(m.group(1)) if m else None, "predict_total_s": float(m.group(2)) if m else None,
                     "profile": total, "per_call": prof[:-1]}
        print(name, json.dumps(out[name]), flush=True)
        if prof:
            print("\n".join(prof[:-1][:20]), flush=True)
    if only:
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"config5_n{n}_{'_'.join(sorted(only))}.json").write_text(json.dumps(out, indent=1))
        sys.exit(0)
    same = out["b200_exact"]["md5"] == out["reference_cpu"]["md5"]
    print(f"CONFIG5 cover identical to the single-thread CPU reference: {same}; "
          f"fast == exact: {out['b200_fast']['md5'] == out['b200_exact']['md5']}; "
          f"reference all threads == one thread: {out['reference_cpu_all_threads']['md5'] == out['reference_cpu']['md5']}")
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"config5_n{n}.json").write_text(json.dumps(out, indent=1))
