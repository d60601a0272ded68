#!/usr/bin/env python
"""How fast is the sequential neighbour sum of ONE huge vertex?  Star graphs (hub of degree D, D
leaves) leave nothing else to do, so the stage time is the hub's chain: ns per neighbour, exact and
fast mode, for every stage.  With `busy` > 0 the graph also gets `busy` vertices of degree 32 that
keep the other warps of the GPU occupied while the chain runs.
usage: python tools/chain_probe.py [D=262144 | rmatSCALE] [busy=0]"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200 as pkg  # noqa: E402
from gnn_mwvc_b200 import capi, graphs  # noqa: E402

RMAT = len(sys.argv) > 1 and sys.argv[1].startswith("rmat")      # "rmat20": the bench graph instead of a star
D = 262144 if RMAT or len(sys.argv) < 2 else int(sys.argv[1])
busy = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
n = 1 + D + busy
eu = [torch.zeros(D, dtype=torch.int64), ]
ev = [torch.arange(1, D + 1, dtype=torch.int64)]
if busy:
    b0 = 1 + D
    src = torch.arange(busy, dtype=torch.int64).repeat_interleave(16)
    dst = (src + torch.randint(1, busy, (src.numel(),))) % busy
    eu.append(b0 + src); ev.append(b0 + dst)
a, b = graphs._canonical_edges(torch.cat(eu).to(dev), torch.cat(ev).to(dev), n)
g = graphs.graph_from_edges(n, a, b, graphs.random_weights(n, 5, dev), name="star")
if RMAT:
    g = graphs.rmat_graph(int(sys.argv[1][4:]), 16, seed=42, device=dev)
    n = g.n
    D = int((g.row_ptr[1:] - g.row_ptr[:-1]).max())
ctx = pkg.Context(0)
ctx.model_upload(capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz"))
ctx.graph_adopt(g.row_ptr.to(torch.int32).contiguous(), g.col, g.weights, g.nw)
stream = ctx.torch_stream()
x = (g.weights.to(torch.float32) / 200.0).contiguous()
h1 = torch.rand(n, 16, device=dev)
h2 = torch.rand(n, 16, device=dev)
sc = torch.zeros(n, device=dev)
out = {"hub_degree": D, "busy_vertices": busy}
with torch.cuda.stream(stream):
    for mode, name in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
        ms = [[], [], []]
        for it in range(8):
            for st, (p, q) in enumerate(((x, h1), (h1, h2), (h2, sc))):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.stage_device(st, p, q, 200.0, mode)
                e1.record(stream)
                e1.synchronize()
                if it >= 3:
                    ms[st].append(e0.elapsed_time(e1))
        if mode == pkg.MODE_EXACT:
            px = []
            for st, (p, q) in enumerate(((x, h1), (h1, h2), (h2, sc))):
                ctx.stage_device(st, p, q, 200.0, mode)
                px.append(ctx.px_stats())
            out["px"] = px
        out[name + "_ms"] = [round(float(np.mean(m)), 4) for m in ms]
        out[name + "_ns_per_neighbour"] = [round(float(np.mean(m)) * 1e6 / D, 2) for m in ms]
print(json.dumps(out), flush=True)
os._exit(0)
