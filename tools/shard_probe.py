#!/usr/bin/env python
"""Compute time of ONE rank of a P-GPU run, measured on one GPU: build the R-MAT graph of the given
scale, relabel it over P equal-work shards (as bench.py does for N > 1), adopt shard 0 and time the
three stage kernels against full-size h1/h2 buffers.  Shows what the slowest part of a rank's step
is without needing P GPUs (the exchange is not part of it).
usage: python tools/shard_probe.py <scale> <parts> [steps]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200 as pkg  # noqa: E402
from gnn_mwvc_b200 import capi, graphs, dist as gdist  # noqa: E402

scale, parts = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
g = graphs.rmat_graph(scale, 16, seed=42, device=dev)
n = g.n
if parts > 1:
    g, perm = graphs.balanced_relabel(g, parts)
per = g.n // parts
bounds = [r * per for r in range(parts + 1)]
shard = gdist.make_shard(g, bounds, 0)
deg = (shard.row_ptr[1:] - shard.row_ptr[:-1])
ctx = pkg.Context(0)
ctx.model_upload(capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz"))
ctx.graph_adopt(shard.row_ptr.to(torch.int32).contiguous(), shard.col, shard.weights, shard.nw, n_global=g.n,
                v_begin=0, v_end=per)
stream = ctx.torch_stream()
x = (g.weights.to(torch.float32) / 200.0).contiguous()
h1 = torch.rand(g.n, 16, device=dev)
h2 = torch.rand(g.n, 16, device=dev)
sc = torch.zeros(per, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"scale": scale, "parts": parts, "n_local": per, "nnz_local": shard.nnz, "max_degree": int(deg.max().item())}
with torch.cuda.stream(stream):
    for mode, name in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
        ms = [[], [], []]
        for it in range(steps + 3):
            for st, (a, b) in enumerate(((x, h1), (h1, h2), (h2, sc))):
                flush.fill_(it & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.stage_device(st, a, b, 200.0, mode)
                e1.record(stream)
                e1.synchronize()
                if it >= 3:
                    ms[st].append(e0.elapsed_time(e1))
        out[name] = [round(float(np.mean(m)), 4) for m in ms]
print(json.dumps(out), flush=True)
import os
os._exit(0)
