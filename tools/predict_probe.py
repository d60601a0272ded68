#!/usr/bin/env python
"""Where does one gnn::model::predict go?  Builds a reduction_graph through the drop-in's C binding
(gnn-mwvc_b200/dropin.py), calls predict a few times with GVC_TRACE=1 (libgvc prints the steps of the
upload and the forward on stderr, synchronising at each) and without (wall time of the call).
usage: python tools/predict_probe.py [er|rmat] [size: vertices for er, scale for rmat] [packed]"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
kind = sys.argv[1] if len(sys.argv) > 1 else "er"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
if len(sys.argv) > 3 and sys.argv[3] == "packed":
    os.environ["GVC_UPLOAD"] = "packed"
trace = os.environ.get("GVC_TRACE")
import torch  # noqa: E402
import gnn_mwvc_b200  # noqa: E402,F401
from gnn_mwvc_b200 import capi, dropin, graphs  # noqa: E402

dev = "cuda" if torch.cuda.is_available() else "cpu"
g = graphs.er_graph(size, 5 * size, seed=1, device=dev) if kind == "er" else graphs.rmat_graph(size, 16, seed=42, device=dev)
eu, ev = g.edges_numpy()
W = g.weights.cpu().numpy().view(np.uint32)
x = W.astype(np.float32) / np.float32(200.0)
d = dropin.Dropin()
m = d.model(dropin.model_text(capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz")))
gh = d.graph(g.n, eu, ev, W)
secs = []
for i in range(8):
    if trace:
        print(f"--- predict call {i}", file=sys.stderr, flush=True)
    d.predict(m, gh, x, 200.0)
    secs.append(d.last_seconds)
print(json.dumps({"graph": g.name, "n": g.n, "edges": g.n_edges, "upload": os.environ.get("GVC_UPLOAD", "stream"),
                  "traced": bool(trace), "predict_ms": [round(1e3 * s, 3) for s in secs]}), flush=True)
