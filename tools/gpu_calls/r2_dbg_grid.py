import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gnn_mwvc_b200 as pkg
from gnn_mwvc_b200 import capi, graphs
dev = torch.device("cuda:0")
ctx = pkg.Context(0)
ctx.model_upload(capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz"))
for side in (2001, 3000, 4472):
    g = graphs.grid_graph(side, side, device=dev)
    s = 200.0
    rp = g.row_ptr.to(torch.int32).contiguous()
    for what, w, nw in (("random", g.weights, g.nw), ("uniform", torch.full_like(g.weights, 7), (g.row_ptr[1:] - g.row_ptr[:-1]).to(torch.int32) * 7)):
        ctx.graph_adopt(rp, g.col, w, nw)
        dx = (w.to(torch.float32) / s).contiguous()
        a = torch.empty(g.n, device=dev)
        torch.cuda.synchronize()
        for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
            try:
                ctx.forward_device(dx, s, a, mode); ctx.sync()
                print(side, what, mode, "ok", float(a[:3].sum()), flush=True)
            except Exception as e:
                print(side, what, mode, "FAILED", str(e)[:200], flush=True)
                sys.exit(1)
