#!/usr/bin/env python
"""How much of the large-graph stage time is L2 capacity?  Times the three stage kernels on an R-MAT graph as
generated and on the same graph with its vertices relabelled in order of descending degree (the rows of the
hubs then lie next to each other: a hot 64-byte row no longer shares its 128-byte line with a cold one).  Results
are bit-identical up to the permutation; this only measures what an in-library relabelling of the row storage
could win.  usage: python tools/relabel_probe.py [scale=23] [steps=10]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200 as pkg  # noqa: E402
from gnn_mwvc_b200 import capi, graphs  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 23
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
g0 = graphs.rmat_graph(scale, 16, seed=42, device=dev)
out = {"scale": scale, "n": g0.n, "nnz": g0.nnz}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in ("as_generated", "by_degree"):
    g = g0 if name == "as_generated" else graphs.balanced_relabel(g0, 1)[0]
    ctx = pkg.Context(0)
    ctx.model_upload(capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz"))
    ctx.graph_adopt(g.row_ptr.to(torch.int32).contiguous(), g.col, g.weights, g.nw)
    stream = ctx.torch_stream()
    x = (g.weights.to(torch.float32) / 200.0).contiguous()
    h1 = torch.empty(g.n, 16, device=dev)
    h2 = torch.empty(g.n, 16, device=dev)
    sc = torch.zeros(g.n, device=dev)
    with torch.cuda.stream(stream):
        for mode, mname in ((pkg.MODE_EXACT, "exact"), (pkg.MODE_FAST, "fast")):
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
            out[f"{name}_{mname}"] = [round(float(np.mean(m)), 4) for m in ms]
    checksum = float(sc.double().sum().item())
    out[f"{name}_score_sum"] = checksum
    ctx.close()
    del ctx, h1, h2, sc, x
print(json.dumps(out), flush=True)
