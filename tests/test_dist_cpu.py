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


class _HostPeerCtx:
    """CPU stand-in for the peer-memory entry points of capi.Context: buffers are named shared-memory
    blocks (the 64-byte handle is the name), and `mirror` does what the store epilogue of the stage
    kernels does -- write each finished row into the buffers of the ranks that own a neighbour."""

    def __init__(self, g):
        from multiprocessing import shared_memory
        self.shm_mod, self.g = shared_memory, g
        self.blocks, self.peers, self.owners = {}, {0: [], 1: []}, None
        self.rp = g.row_ptr.numpy().astype(np.int64)
        self.col = g.col.numpy().view(np.uint32).astype(np.int64)

    def torch_device(self):
        return torch.device("cpu")

    def peer_alloc(self, nbytes):
        b = self.shm_mod.SharedMemory(create=True, size=nbytes)
        np.frombuffer(b.buf, np.float32)[:] = np.nan           # rows that never arrive stay visible
        self.blocks[id(b)] = b
        return id(b), b.name.encode().ljust(64, b"\0")

    def peer_open(self, handle):
        b = self.shm_mod.SharedMemory(name=handle.rstrip(b"\0").decode())
        self.blocks[id(b)] = b
        return id(b)

    def peer_tensor(self, ptr, shape):
        return torch.frombuffer(self.blocks[ptr].buf, dtype=torch.float32).view(shape)

    def peer_close(self, ptr):
        self.blocks.pop(ptr).close()

    def peer_free(self, ptr):
        b = self.blocks.pop(ptr)
        b.close()
        b.unlink()

    def stage_peers(self, stage, ptrs):
        self.peers[stage] = list(ptrs)

    def peer_owners(self, bounds, peer_of_part):
        self.owners = (list(bounds), list(peer_of_part)) if len(peer_of_part) else None

    def mirror(self, stage, rows, a, b):
        """rows: this rank's freshly computed [b - a, 16] block of the stage's output."""
        bounds, peer_of_part = self.owners
        part_of = np.searchsorted(np.asarray(bounds[1:]), np.arange(self.g.n), side="right")
        for u in range(a, b):
            readers = {peer_of_part[k] for k in part_of[self.col[self.rp[u]:self.rp[u + 1]]]} - {-1}
            for q in readers:
                self.peer_tensor(self.peers[stage][q], (self.g.n, 16))[u] = torch.from_numpy(rows[u - a].copy())


def _peer_worker(rank, world, port, ret):
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
    g0 = graphs.rmat_graph(10, 8, seed=9, n_limit=1021)
    rp0, col0, W0, NW0, x0, s = inputs_of(g0)
    g, perm = graphs.balanced_relabel(g0, world)
    per = g.n // world
    bounds = [r * per for r in range(world + 1)]
    rp, col, W, NW, x, _ = inputs_of(g, s)
    shard = gdist.make_shard(g, bounds, rank)
    a, b = shard.v_begin, shard.v_end
    ctx = _HostPeerCtx(g)
    pr = gdist.PeerRows(ctx, g.n, bounds=bounds)
    assert len(ctx.peers[0]) == world - 1 and ctx.owners[1][rank] == -1

    groups, cur = [], []
    for L in layers:
        if L[0] == po.GRAPH and cur:
            groups.append(cur)
            cur = []
        cur.append(L)
    groups.append(cur)

    def stage_fn(stage, d_in, d_out, scale, mode):
        act = np.nan_to_num(d_in.numpy().reshape(g.n, -1).astype(np.float32), nan=123.0)   # unread rows hold NaN
        for kind, Wm, bias in groups[stage]:
            act = (orc.graph_forward(rp, col, W, NW, scale, act) if kind == po.GRAPH else
                   orc.linear_forward(act, Wm, bias) if kind == po.LINEAR else
                   orc.relu(act) if kind == po.RELU else orc.sigmoid(act))
        if stage < 2:
            d_out[a:b] = torch.from_numpy(act[a:b])
            ctx.mirror(stage, act[a:b], a, b)
        else:
            d_out.copy_(torch.from_numpy(act[a:b, 0]))

    scores = torch.empty(b - a)
    for _ in range(2):                                          # the second pass reuses the buffers
        gdist.sharded_forward(stage_fn, shard, torch.from_numpy(x.copy()), pr.h1, pr.h2, scores, s, 0, peer_rows=pr)
    pr.barrier()
    full = gdist.gather_scores(scores, bounds)
    # exactly the rows this rank reads (its vertices' neighbours) and its own arrived
    read = np.zeros(g.n, bool)
    read[col[int(rp[a]):int(rp[b])]] = True
    read[a:b] = True
    missing = torch.isnan(pr.h1).any(1).numpy()
    ok_rows = bool((~missing[read]).all() and missing[~read].all())
    if rank == 0:
        want = orc.predict(orc.parse(po.layers_to_text(layers)), rp0, col0, W0, NW0, x0, s)[:, 0]
        ret["equal"] = bool(np.array_equal(full.numpy()[perm.numpy()].view(np.uint32), want.view(np.uint32)))
    ret[f"rows{rank}"] = ok_rows
    pr.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_rows_protocol_over_gloo(world):
    """dist.PeerRows -- what bench.py runs at N > 1 -- with a CPU stand-in for the CUDA side: handles
    travel with an all-gather, every rank maps the others' buffers, the 'kernels' store their rows
    into the buffers of the ranks that read them, the exchange is a barrier."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["equal"]
    assert all(ret[f"rows{r}"] for r in range(world))


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
