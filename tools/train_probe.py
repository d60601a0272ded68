#!/usr/bin/env python
"""One training step of the GNN_VC architecture on a synthetic graph, timed: gvc_trainer_predict +
gvc_trainer_mse_backprop + gvc_trainer_sgd_step in fast and exact mode, next to the
reference's own old_files/src/lib/gnn_training.cpp (oracle/_ref/libgnntrainref.so, all host threads) where that
library travelled.  usage: python tools/train_probe.py [rmat scale=18] [steps=5]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200 as pkg  # noqa: E402
from gnn_mwvc_b200 import capi, graphs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 18
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = graphs.rmat_graph(scale, 16, seed=42)
row_ptr, col, W, NW = g.numpy()
s = float(W.max())
x = (W.astype(np.float32) / np.float32(s)).reshape(-1, 1)
rng = np.random.default_rng(1)
y = (rng.random((g.n, 1)) < 0.4).astype(np.float32)
layers = capi.random_model(7)
ctx = pkg.Context(0)
ctx.graph_upload(row_ptr, col, W, NW)
out = {"graph": g.name, "n": g.n, "edges": g.n_edges, "steps": steps}
for mode, name in ((pkg.MODE_FAST, "fast"), (pkg.MODE_EXACT, "exact")):
    tr = capi.Trainer(ctx, layers)
    ms, loss = [], None
    for it in range(steps + 1):
        t0 = time.perf_counter()
        tr.predict(x, s, mode, want_out=False)
        t1 = time.perf_counter()
        loss = tr.mse_backprop(y, mode)
        t2 = time.perf_counter()
        tr.sgd_step(g.n, lr=0.01, momentum=0.9, weight_decay=0.0)
        tr.zero_grad()
        t3 = time.perf_counter()
        if it:
            ms.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)])
    m = np.median(np.array(ms), axis=0)
    out[name] = {"predict_ms": round(float(m[0]), 3), "mse_backprop_ms": round(float(m[1]), 3), "sgd_zero_ms": round(float(m[2]), 3),
                 "step_ms": round(float(m.sum()), 3), "edges_per_s": round(g.n_edges / (m.sum() * 1e-3)), "last_loss": loss}
    tr.close()
if po.TRAIN_REF_SO.exists() and scale <= 20:
    import os
    os.environ.pop("OPENBLAS_NUM_THREADS", None)
    ref = po.TrainHarness(threads=None)
    eu, ev = g.edges_numpy()
    h = ref.create(layers, scales=np.full(len(layers), s, np.float32))
    ref.set_graph(h, g.n, eu, ev, W)
    ms = []
    for it in range(2):
        t0 = time.perf_counter()
        ref.predict(h, x)
        t1 = time.perf_counter()
        ref.mse_step(h, y)
        t2 = time.perf_counter()
        ref.sgd_step(h, g.n, lr=0.01, momentum=0.9, wd=0.0)
        ref.zero_grad(h)
        t3 = time.perf_counter()
        ms.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)])
    m = np.array(ms)[-1]
    out["reference_cpu"] = {"predict_ms": round(float(m[0]), 1), "mse_backprop_ms": round(float(m[1]), 1), "sgd_zero_ms": round(float(m[2]), 3),
                            "step_ms": round(float(m.sum()), 1), "edges_per_s": round(g.n_edges / (m.sum() * 1e-3)),
                            "threads": os.cpu_count()}
print(json.dumps(out), flush=True)
