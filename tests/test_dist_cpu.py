"""CPU, world_size 2 and 3 over gloo: the host logic of the vertex-range sharding
(gnn-mwvc_b200/dist.py) -- shard cutting, the in-place exchange of unequal row slices, the
three-stage driver -- with the oracle standing in for the per-shard stage kernels.  The
sharded result must equal the single-process forward bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret, layout):
    import sys
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gnn_mwvc_b200  # noqa: F401
    from gnn_mwvc_b200 import capi, graphs
    from gnn_mwvc_b200 import dist as gdist
    from helpers import inputs_of
    from oracle import pyoracle as po

    layers = capi.load_model_npz(GOLDEN / "mwvc_model.npz")
    orc = po.Oracle()
    g = graphs.rmat_graph(11, 12, seed=5, n_limit=1777)
    if layout == "balanced":
        # the benchmark's layout: equal shards dealt by descending degree, rows of isolated
        # vertices stay home
        g0, (rp0, col0, W0, NW0, x0, s0) = g, inputs_of(g)
        g, perm = graphs.balanced_relabel(g0, world)
        per = g.n // world
        bounds = [r * per for r in range(world + 1)]
    else:
        s0 = None
        bounds = graphs.nnz_balanced_ranges(g.row_ptr, world)
    rp, col, W, NW, x, s = inputs_of(g, s0)
    shard = gdist.make_shard(g, bounds, rank, skip_isolated=(layout == "balanced"))
    a, b = shard.v_begin, shard.v_end
    assert shard.row_ptr[0] == 0 and shard.nnz == int(rp[b] - rp[a])
    assert np.array_equal(shard.col.numpy().view(np.uint32), col[int(rp[a]):int(rp[b])])

    # the stage "kernel": whole-graph oracle layers applied to the (exchanged) full input,
    # of which only this rank's rows are kept -- a wrong or missing exchange changes them
    groups, cur = [], []
    for L in layers:
        if L[0] == po.GRAPH and cur:
            groups.append(cur)
            cur = []
        cur.append(L)
    groups.append(cur)

    def stage_fn(stage, d_in, d_out, scale, mode):
        act = d_in.numpy().reshape(g.n, -1).astype(np.float32)
        for kind, Wm, bias in groups[stage]:
            if kind == po.GRAPH:
                act = orc.graph_forward(rp, col, W, NW, scale, act)
            elif kind == po.LINEAR:
                act = orc.linear_forward(act, Wm, bias)
            elif kind == po.RELU:
                act = orc.relu(act)
            else:
                act = orc.sigmoid(act)
        if stage < 2:
            d_out[a:b] = torch.from_numpy(act[a:b])
        else:
            d_out.copy_(torch.from_numpy(act[a:b, 0]))

    x_full = torch.from_numpy(x.copy())
    h1 = torch.full((g.n, 16), float("nan"))
    h2 = torch.full((g.n, 16), float("nan"))
    scores = torch.empty(b - a)
    gdist.sharded_forward(stage_fn, shard, x_full, h1, h2, scores, s, 0)
    full = gdist.gather_scores(scores, bounds)
    if layout == "balanced":
        deg = np.diff(rp.astype(np.int64))
        missing = torch.isnan(h1).any(1).numpy()
        assert not missing[deg > 0].any()                  # every row somebody can read arrived
        assert not missing[a:b].any()
        assert missing.sum() > 0                           # and rows of remote isolated vertices did not travel
        assert len(set(shard.live)) >= 1 and max(shard.live) < per
    else:
        assert not torch.isnan(h1).any() and not torch.isnan(h2).any()     # every row arrived everywhere
    if rank == 0:
        if layout == "balanced":
            want = orc.predict(orc.parse(po.layers_to_text(layers)), rp0, col0, W0, NW0, x0, s)[:, 0]
            got = full.numpy()[perm.numpy()]
        else:
            want = orc.predict(orc.parse(po.layers_to_text(layers)), rp, col, W, NW, x, s)[:, 0]
            got = full.numpy()
        ret["equal"] = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
        ret["bounds"] = bounds
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_forward_over_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, "ranges"), nprocs=world, join=True)
    assert ret["equal"]
    b = ret["bounds"]
    assert b[0] == 0 and len(b) == world + 1 and len(set(np.diff(b))) > 1   # unequal slices were exchanged


def test_balanced_shards_skip_isolated_rows_over_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret, "balanced"), nprocs=2, join=True)
    assert ret["equal"]


def test_shard_cut_is_a_partition():
    import gnn_mwvc_b200  # noqa: F401
    from gnn_mwvc_b200 import graphs
    from gnn_mwvc_b200 import dist as gdist
    g = graphs.er_graph(1000, 4000, seed=3)
    bounds = graphs.nnz_balanced_ranges(g.row_ptr, 4)
    shards = [gdist.make_shard(g, bounds, r) for r in range(4)]
    assert sum(s.n_local for s in shards) == g.n and sum(s.nnz for s in shards) == g.nnz
    assert torch.equal(torch.cat([s.col for s in shards]), g.col)
    assert torch.equal(torch.cat([s.weights for s in shards]), g.weights)
    for s in shards:
        assert int(s.row_ptr[-1]) == s.nnz
