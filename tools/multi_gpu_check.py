#!/usr/bin/env python
"""Multi-GPU parity: the sharded forward on N GPUs (one process per GPU, NCCL) must equal the
single-GPU forward bit for bit.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/multi_gpu_check.py [rmat_scale]
Rank 0 prints one line `MULTI_GPU_PARITY ok ...` or raises."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200 as pkg  # noqa: E402
from gnn_mwvc_b200 import capi, graphs  # noqa: E402
from gnn_mwvc_b200 import dist as gdist  # noqa: E402


def main():
    scale_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    layers = capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz")
    g = graphs.rmat_graph(scale_log2, 16, seed=7, device=dev, n_limit=(1 << scale_log2) - 3)   # odd vertex count
    s = 200.0
    x = (g.weights.to(torch.float32) / s).contiguous()
    for layout in ("ranges", "balanced"):
        run_layout(layout, g, x, s, layers, rank, world, local, dev)
    if rank == 0:
        print(f"MULTI_GPU_PARITY ok world={world} n={g.n} E={g.n_edges}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_layout(layout, g0, x0, s, layers, rank, world, local, dev):
    """ranges: work-balanced contiguous ranges of the original ids (unequal, per-owner broadcasts);
    balanced: equal ranges of the degree-dealt relabelled graph (one all-gather per exchange)."""
    if layout == "balanced":
        g, perm = graphs.balanced_relabel(g0, world)
        x = (g.weights.to(torch.float32) / s).contiguous()
        per = g.n // world
        bounds = [r * per for r in range(world + 1)]
        tail = int(perm[g0.n - 1].item()) if g0.n % 2 else None
    else:
        g, perm, x = g0, None, x0
        bounds = graphs.nnz_balanced_ranges(g.row_ptr, world)
        tail = "default"
    shard = gdist.make_shard(g, bounds, rank, skip_isolated=(layout == "balanced"))
    for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
        ctx = pkg.Context(local)
        ctx.model_upload(layers)
        ctx.graph_adopt(shard.row_ptr.to(torch.int32).contiguous(), shard.col, shard.weights, shard.nw,
                        n_global=g.n, v_begin=shard.v_begin, v_end=shard.v_end)
        if tail != "default":
            ctx.graph_set_tail(tail)
        h1 = torch.zeros(g.n, 16, device=dev)
        h2 = torch.zeros(g.n, 16, device=dev)
        sc = torch.zeros(shard.n_local, device=dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(ctx.torch_stream()):
            gdist.sharded_forward(ctx.stage_device, shard, x, h1, h2, sc, s, mode)
            full = gdist.gather_scores(sc, bounds)
        torch.cuda.synchronize()
        # the same with the rows delivered by the stage kernels themselves (peer memory) -- twice,
        # the second pass overwrites buffers the first one filled
        pr = gdist.PeerRows(ctx, g.n, bounds=bounds if layout == "balanced" else None)
        sc2 = torch.zeros(shard.n_local, device=dev)
        with torch.cuda.stream(ctx.torch_stream()):
            for _ in range(2):
                gdist.sharded_forward(ctx.stage_device, shard, x, pr.h1, pr.h2, sc2, s, mode, peer_rows=pr)
            pr.barrier()
        torch.cuda.synchronize()
        live = torch.zeros(g.n, dtype=torch.bool, device=dev)
        live[:] = (g.row_ptr[1:] - g.row_ptr[:-1]).to(dev) > 0
        if layout == "balanced":      # rows only travel to the ranks that read them: compare what this rank reads
            live[:] = False
            live[shard.col.to(torch.int64) & 0xFFFFFFFF] = True
            live[shard.v_begin:shard.v_end] = True
        peer_ok = torch.equal(sc2, sc) and torch.equal(pr.h1[live], h1[live]) and torch.equal(pr.h2[live], h2[live])
        pr.close()
        # single-GPU forward of the whole graph on every rank
        one = pkg.Context(local)
        one.model_upload(layers)
        one.graph_adopt(g0.row_ptr.to(torch.int32).contiguous(), g0.col, g0.weights, g0.nw)
        ref = torch.zeros(g0.n, device=dev)
        torch.cuda.synchronize()
        one.forward_device(x0, s, ref, mode)
        one.sync()
        same = torch.equal(full[perm] if perm is not None else full, ref) and peer_ok
        t = torch.tensor([int(same)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            assert int(t.item()) == 1, f"{layout}, mode {mode}: {world}-GPU scores differ from 1-GPU scores"
        del h1, h2, sc, full, ref
        torch.cuda.synchronize()
        ctx.close()
        one.close()


if __name__ == "__main__":
    main()
    sys.stdout.flush()
    os._exit(0)
