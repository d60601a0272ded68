"""GPU: the CUDA path, called through the C ABI (include/gvc.h), against the oracle and the
committed golden vectors.  Exact mode is bit-for-bit; fast mode within the 1e-4 relative
tolerance BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
import gnn_mwvc_b200 as pkg
from gnn_mwvc_b200 import capi, graphs
from helpers import (assert_bit_equal, assert_rel_close, golden_graph, golden_names, inputs_of,
                     oracle_stages)
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
FAST_RTOL = 1e-4     # north_star: per-vertex scores within 1e-4 relative in fp32


@pytest.fixture(scope="module")
def ctx(model_layers):
    c = pkg.Context(0)
    c.model_upload(model_layers)
    assert c.fused, "the GNN_VC architecture must take the fused path"
    yield c
    c.close()


@pytest.fixture(scope="module")
def vec():
    return np.load(GOLDEN / "predict_vectors.npz")


@pytest.fixture(scope="module")
def lay():
    return np.load(GOLDEN / "layer_vectors.npz")


def run(ctx, g, scale=None, mode=pkg.MODE_EXACT):
    rp, col, W, NW, x, s = inputs_of(g, scale)
    ctx.graph_upload(rp, col, W, NW)
    return ctx.forward(x, s, mode)


def test_golden_vectors_exact(ctx, vec):
    for name in golden_names(vec):
        g, s, want = golden_graph(vec, name)
        assert_bit_equal(run(ctx, g, s), want, name)


def test_golden_vectors_fast(ctx, vec):
    for name in golden_names(vec):
        g, s, want = golden_graph(vec, name)
        assert_rel_close(run(ctx, g, s, pkg.MODE_FAST), want, FAST_RTOL, name)


def test_known_answers(ctx, vec):
    g, s, _ = golden_graph(vec, "readme")
    np.testing.assert_allclose(run(ctx, g, s), [0.129362747, 0.129362747, 0.934816003], rtol=0, atol=5e-9)


def test_empty_graph_is_noop(ctx):
    z32, z64 = np.zeros(0, np.uint32), np.zeros(1, np.uint64)
    ctx.graph_upload(z64, z32, z32, z32)
    out = ctx.forward(np.zeros(0, np.float32), 20.0)
    assert out.shape == (0,)


@pytest.mark.parametrize("maker", [
    lambda: graphs.er_graph(20001, 100000, seed=31),      # odd n: OpenBLAS 1-row tail
    lambda: graphs.er_graph(9951, 49701, seed=32),        # first predict of the ER-10k run has n=9951
    lambda: graphs.rmat_graph(15, 16, seed=33),           # hubs of degree ~10^3..10^4
    lambda: graphs.grid_graph(150, 151, seed=34),
    lambda: graphs.er_graph(33, 40, seed=35),             # a single partial tile
    lambda: graphs.er_graph(257, 300, seed=36),           # one CTA + one vertex
    lambda: graphs.graph_from_edges(64, torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64),
                                    graphs.random_weights(64, 37)),   # all isolated
])
def test_forward_vs_oracle(ctx, oracle, oracle_model, maker):
    g = maker()
    rp, col, W, NW, x, s = inputs_of(g)
    want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
    ctx.graph_upload(rp, col, W, NW)
    assert_bit_equal(ctx.forward(x, s, pkg.MODE_EXACT), want, g.name)
    assert_rel_close(ctx.forward(x, s, pkg.MODE_FAST), want, FAST_RTOL, g.name)
    # identical downstream decision wherever the score is not within tolerance of 0.5
    fast = ctx.forward(x, s, pkg.MODE_FAST)
    clear = np.abs(want - 0.5) > 1e-4
    assert np.array_equal((fast > 0.5)[clear], (want > 0.5)[clear])


def test_degree_ladder_hits_every_gather_path(ctx, oracle, oracle_model):
    """Degrees around the class thresholds of the gather schedule (64: 4-lanes-per-vertex mid tasks,
    2048: the CTA-wide ring with its warp-to-warp hand-over and the single-warp stage-0 chains,
    4096: the fast-mode hub chunks, 16384: the CTA-wide stage-0 sums with their 256-neighbour blocks
    and 7 loader warps) and a giant of ~70 000 neighbours whose ring runs many rounds on all 8 warps;
    shuffled adjacency so the order matters.  Bit-exact."""
    rng = np.random.default_rng(11)
    n = 90_000
    want_deg = {0: 70_001, 1: 2047, 2: 2048, 3: 2049, 4: 4097, 5: 63, 6: 64, 7: 65, 8: 511, 9: 129, 10: 1, 11: 8191,
                12: 4096, 13: 8192, 14: 16_383, 15: 16_384, 16: 16_385, 17: 16_384 + 7 * 256, 18: 16_384 + 255}
    eu, ev = [], []
    for u, d in want_deg.items():
        nb = rng.choice(np.arange(100, n), size=d, replace=False)
        eu.append(np.full(d, u)); ev.append(nb)
    extra = rng.integers(100, n, size=(200_000, 2))
    eu.append(extra[:, 0]); ev.append(extra[:, 1])
    eu, ev = torch.from_numpy(np.concatenate(eu)), torch.from_numpy(np.concatenate(ev))
    a, b = graphs._canonical_edges(eu, ev, n)
    g = graphs.graph_from_edges(n, a, b, graphs.random_weights(n, 12), name="ladder")
    rp, col, W, NW, x, s = inputs_of(g)
    col = col.copy()
    for u in want_deg:                                   # the big lists in a non-ascending order
        rng.shuffle(col[int(rp[u]):int(rp[u + 1])])
    want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
    ctx.graph_upload(rp, col, W, NW)
    assert_bit_equal(ctx.forward(x, s, pkg.MODE_EXACT), want, "ladder")
    assert_rel_close(ctx.forward(x, s, pkg.MODE_FAST), want, FAST_RTOL, "ladder fast")


def test_stage_outputs_vs_oracle(ctx, oracle, model_layers):
    g = graphs.rmat_graph(12, 16, seed=41, n_limit=4001)
    rp, col, W, NW, x, s = inputs_of(g)
    h1, h2, scores = oracle_stages(oracle, model_layers, rp, col, W, NW, x, s)
    ctx.graph_upload(rp, col, W, NW)
    dev = torch.device("cuda:0")
    dx = torch.from_numpy(x).to(dev)
    d1 = torch.empty(g.n, 16, device=dev)
    d2 = torch.empty(g.n, 16, device=dev)
    ds = torch.empty(g.n, device=dev)
    torch.cuda.synchronize()
    ctx.stage_device(0, dx, d1, s)
    ctx.stage_device(1, d1, d2, s)
    ctx.stage_device(2, d2, ds, s)
    ctx.sync()
    rows = ctx.row_order()                       # the identity unless the context keeps its rows in schedule order
    got1, got2 = np.empty_like(h1), np.empty_like(h2)
    got1[rows], got2[rows] = d1.cpu().numpy(), d2.cpu().numpy()
    assert_bit_equal(got1, h1, "h1")
    assert_bit_equal(got2, h2, "h2")
    assert_bit_equal(ds.cpu().numpy(), scores, "scores")


def test_weight_scale_and_x_are_honoured(ctx, oracle, oracle_model):
    # predict must use the `in` it is given and the model's scale, not W/max(W)
    g = graphs.er_graph(3000, 12000, seed=42)
    rp, col, W, NW, _, _ = inputs_of(g)
    x = np.random.default_rng(1).random(g.n).astype(np.float32) * 3
    want = oracle.predict(oracle_model, rp, col, W, NW, x, 120.0)[:, 0]
    ctx.graph_upload(rp, col, W, NW)
    assert_bit_equal(ctx.forward(x, 120.0), want, "scale 120")


def test_adjacency_order_is_respected(ctx, oracle, oracle_model):
    # after folds the reference's adjacency is not ascending; sums follow the given order
    g = graphs.er_graph(2000, 16000, seed=43)
    rp, col, W, NW, x, s = inputs_of(g)
    rng = np.random.default_rng(2)
    col = col.copy()
    for u in range(g.n):
        rng.shuffle(col[int(rp[u]):int(rp[u + 1])])
    want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
    ctx.graph_upload(rp, col, W, NW)
    assert_bit_equal(ctx.forward(x, s), want, "shuffled adjacency")


def test_graph_is_reuploaded_between_calls(ctx, oracle, oracle_model):
    # GNN_VC calls predict on a shrinking graph 4-14 times with one model
    for n, m, seed in ((5000, 20000, 51), (1500, 4000, 52), (301, 500, 53), (0, 0, 54), (77, 100, 55)):
        g = graphs.er_graph(n, m, seed=seed) if n else None
        if g is None:
            z32, z64 = np.zeros(0, np.uint32), np.zeros(1, np.uint64)
            ctx.graph_upload(z64, z32, z32, z32)
            assert ctx.forward(np.zeros(0, np.float32), 200.0).size == 0
            continue
        rp, col, W, NW, x, s = inputs_of(g)
        want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
        ctx.graph_upload(rp, col, W, NW)
        assert_bit_equal(ctx.forward(x, s), want, g.name)


def test_shards_are_bit_identical_to_one_gpu(ctx, model_layers):
    """Vertex-range shards (the multi-GPU layout) emulated with several contexts on one GPU:
    per-vertex arithmetic is unchanged, so P shards == 1 shard bit for bit."""
    g = graphs.rmat_graph(13, 16, seed=61, n_limit=8191)
    rp, col, W, NW, x, s = inputs_of(g)
    ctx.graph_upload(rp, col, W, NW)
    want = ctx.forward(x, s)
    dev = torch.device("cuda:0")
    for parts in (2, 3, 8):
        bounds = graphs.nnz_balanced_ranges(g.row_ptr, parts)
        shards = []
        for p in range(parts):
            a, b = bounds[p], bounds[p + 1]
            c = pkg.Context(0)
            c.model_upload(model_layers)
            lo, hi = int(rp[a]), int(rp[b])
            c.graph_upload(rp[a:b + 1] - rp[a], col[lo:hi], W[a:b], NW[a:b], n_global=g.n, v_begin=a, v_end=b)
            shards.append(c)
        dx = torch.from_numpy(x).to(dev)
        d1 = torch.zeros(g.n, 16, device=dev)
        d2 = torch.zeros(g.n, 16, device=dev)
        outs = [torch.empty(bounds[p + 1] - bounds[p], device=dev) for p in range(parts)]
        torch.cuda.synchronize()
        for stage, (src, dst) in enumerate(((dx, d1), (d1, d2), (d2, None))):
            for p, c in enumerate(shards):
                c.stage_device(stage, src, dst if dst is not None else outs[p], s)
            for c in shards:
                c.sync()                      # the "exchange": every shard wrote its rows of the full buffer
        got = torch.cat(outs).cpu().numpy()
        assert_bit_equal(got, want, f"{parts} shards")
        for c in shards:
            c.close()


def test_relabelled_shards_are_bit_identical(ctx, model_layers):
    """The benchmark's multi-GPU layout: vertices dealt to equal shards by descending degree
    (graphs.balanced_relabel), odd vertex count so the exact-mode tail vertex moves too."""
    g0 = graphs.rmat_graph(13, 16, seed=62, n_limit=8189)
    rp0, col0, W0, NW0, x0, s = inputs_of(g0)
    ctx.graph_upload(rp0, col0, W0, NW0)
    want = ctx.forward(x0, s)
    dev = torch.device("cuda:0")
    for parts in (2, 4):
        g, perm = graphs.balanced_relabel(g0, parts)
        rp, col, W, NW, x, _ = inputs_of(g, s)
        per = g.n // parts
        tail = int(perm[g0.n - 1])
        shards = []
        for p in range(parts):
            a, b = p * per, (p + 1) * per
            c = pkg.Context(0)
            c.model_upload(model_layers)
            lo, hi = int(rp[a]), int(rp[b])
            c.graph_upload(rp[a:b + 1] - rp[a], col[lo:hi], W[a:b], NW[a:b], n_global=g.n, v_begin=a, v_end=b)
            c.graph_set_tail(tail)
            shards.append(c)
        dx = torch.from_numpy(x).to(dev)
        d1 = torch.zeros(g.n, 16, device=dev)
        d2 = torch.zeros(g.n, 16, device=dev)
        out = torch.empty(g.n, device=dev)
        torch.cuda.synchronize()
        for stage, (src, dst) in enumerate(((dx, d1), (d1, d2), (d2, None))):
            for p, c in enumerate(shards):
                c.stage_device(stage, src, dst if dst is not None else out[p * per:(p + 1) * per], s)
            for c in shards:
                c.sync()
        got = out.cpu().numpy()[perm.numpy()]
        assert_bit_equal(got, want, f"{parts} relabelled shards")
        for c in shards:
            c.close()


def test_peer_mirrored_rows_replace_the_exchange(ctx, model_layers):
    """Multi-GPU without the collective: every shard's stage kernels also store their rows into the
    OTHER shards' private h1/h2 copies (gvc_stage_peers; here plain device pointers on one GPU, over
    NVLink via CUDA IPC in a real run).  Afterwards every copy holds all rows anybody can read --
    a row only goes to the shards that own a neighbour of its vertex (gvc_peer_owners) -- and the
    scores equal the single-shard ones bit for bit."""
    g0 = graphs.rmat_graph(13, 16, seed=63, n_limit=8189)
    rp0, col0, W0, NW0, x0, s = inputs_of(g0)
    ctx.graph_upload(rp0, col0, W0, NW0)
    want = ctx.forward(x0, s)
    dev = torch.device("cuda:0")
    parts = 3
    g, perm = graphs.balanced_relabel(g0, parts)
    rp, col, W, NW, x, _ = inputs_of(g, s)
    per = g.n // parts
    tail = int(perm[g0.n - 1])
    dx = torch.from_numpy(x).to(dev)
    h1 = [torch.full((g.n, 16), float("nan"), device=dev) for _ in range(parts)]
    h2 = [torch.full((g.n, 16), float("nan"), device=dev) for _ in range(parts)]
    out = torch.empty(g.n, device=dev)
    shards = []
    for p in range(parts):
        a, b = p * per, (p + 1) * per
        c = pkg.Context(0)
        c.model_upload(model_layers)
        c.graph_upload(rp[a:b + 1] - rp[a], col[int(rp[a]):int(rp[b])], W[a:b], NW[a:b], n_global=g.n, v_begin=a, v_end=b)
        c.graph_set_tail(tail)
        c.stage_peers(0, [h1[q].data_ptr() for q in range(parts) if q != p])
        c.stage_peers(1, [h2[q].data_ptr() for q in range(parts) if q != p])
        c.peer_owners([q * per for q in range(parts + 1)], [-1 if q == p else (q if q < p else q - 1) for q in range(parts)])
        shards.append(c)
    torch.cuda.synchronize()
    for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
        for stage in range(3):
            for p, c in enumerate(shards):
                src = (dx, h1[p], h2[p])[stage]
                dst = (h1[p], h2[p], out[p * per:(p + 1) * per])[stage]
                c.stage_device(stage, src, dst, s, mode)
            for c in shards:
                c.sync()                                  # the barrier between two stages
        got = out.cpu().numpy()[perm.numpy()]
        if mode == pkg.MODE_EXACT:
            assert_bit_equal(got, want, "peer-mirrored rows")
        else:
            assert_rel_close(got, want, FAST_RTOL, "peer-mirrored rows, fast")
    src = np.repeat(np.arange(g.n), np.diff(rp.astype(np.int64)))
    for p in range(parts):
        missing = torch.isnan(h1[p]).any(1).cpu().numpy()
        own = np.zeros(g.n, bool); own[p * per:(p + 1) * per] = True
        read = np.zeros(g.n, bool); read[col[own[src]]] = True      # what shard p gathers: its vertices' neighbours
        assert not missing[read | own].any()                  # every row shard p reads arrived in its copy
        assert missing[~read & ~own].all()                    # and nothing else travelled
    for c in shards:
        c.close()


def test_generic_path_any_layer_sequence(oracle):
    """operator>> accepts any sequence of the four layer kinds; non-GNN_VC models run on
    the per-layer kernels (SURVEY.md 8(b) genericity)."""
    rng = np.random.default_rng(7)

    def lin(K, N):
        return (po.LINEAR, (rng.standard_normal((K, N)) * 0.4).astype(np.float32), rng.standard_normal(N).astype(np.float32) * 0.1)
    layers = [(po.GRAPH, None, None), lin(5, 8), (po.RELU, None, None), (po.GRAPH, None, None), lin(19, 4),
              (po.SIGMOID, None, None), lin(4, 1), (po.SIGMOID, None, None)]
    c = pkg.Context(0)
    c.model_upload(layers)
    assert not c.fused
    h = oracle.parse(po.layers_to_text(layers))
    for g in (graphs.er_graph(1001, 4000, seed=71), graphs.er_graph(64, 100, seed=72)):
        rp, col, W, NW, x, s = inputs_of(g)
        want = oracle.predict(h, rp, col, W, NW, x, s)[:, 0]
        c.graph_upload(rp, col, W, NW)
        assert_bit_equal(c.forward(x, s), want, g.name)
        assert_rel_close(c.forward(x, s, pkg.MODE_FAST), want, FAST_RTOL, g.name)
    c.close()


def test_single_layers_vs_golden(ctx, vec, lay):
    """The layer structs' own forward() entry points (host-buffer ABI) against the reference's output."""
    g, _, _ = golden_graph(vec, "readme")
    rp, col, W, NW = g.numpy()
    ctx.graph_upload(rp, col, W, NW)
    assert_bit_equal(ctx.graph_layer_host(lay["graph16.in"], 20.0), lay["graph16.out"], "graph16")
    g, _, _ = golden_graph(vec, "er607")
    rp, col, W, NW = g.numpy()
    ctx.graph_upload(rp, col, W, NW)
    for w in (1, 16, 3):
        assert_bit_equal(ctx.graph_layer_host(lay[f"graph_er607_w{w}.in"], 200.0), lay[f"graph_er607_w{w}.out"], f"w={w}")
    for k in sorted(k[:-3] for k in lay.files if k.startswith("linear_") and k.endswith(".in")):
        assert_bit_equal(ctx.linear_host(lay[k + ".in"], lay[k + ".W"], lay[k + ".b"]), lay[k + ".out"], k)
    assert_bit_equal(ctx.relu_host(lay["relu.in"]), lay["relu.out"], "relu")
    assert_bit_equal(ctx.sigmoid_host(lay["sigmoid.in"]), lay["sigmoid.out"], "sigmoid")
    A = np.random.default_rng(3).standard_normal((7, 5)).astype(np.float32)
    B = np.random.default_rng(4).standard_normal((5, 3)).astype(np.float32)
    np.testing.assert_allclose(ctx.sgemm_host(A, B), A @ B, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ctx.sgemm_host(A.T.copy(), B, trans_a=True), A @ B, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ctx.sgemm_host(A, B.T.copy(), trans_b=True), A @ B, rtol=1e-5, atol=1e-6)


def test_upload_from_pinned_staging(ctx, oracle, oracle_model):
    """The drop-in's way in: CSR written into the context's pinned buffers, uploaded by DMA."""
    for g in (graphs.rmat_graph(12, 8, seed=9), graphs.er_graph(1001, 4000, seed=2)):
        rp, col, W, NW, x, s = inputs_of(g)
        want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
        for _ in range(2):                                   # second round reuses the buffers
            srp, scol, sW, sNW = ctx.graph_staging(g.n, len(col))
            srp[:], scol[:], sW[:], sNW[:] = rp, col, W, NW
            ctx.graph_upload(srp, scol, sW, sNW)
            assert_bit_equal(ctx.forward(x, s), want, g.name)


def test_upload_rejects_broken_csr(ctx, vec):
    g, s, want = golden_graph(vec, "er607")
    rp, col, W, NW, x, _ = inputs_of(g, s)
    bad = col.copy(); bad[len(bad) // 2] = g.n                # an id one past the last vertex
    with pytest.raises(capi.GvcError, match="neighbour id"):
        ctx.graph_upload(rp, bad, W, NW)
    with pytest.raises(capi.GvcError, match="no graph"):
        ctx.forward_device(0, s, 0)                           # the rejected graph is not left behind
    bad = rp.copy(); bad[5], bad[6] = bad[6] + 1, bad[5]       # offsets going backwards
    with pytest.raises(capi.GvcError, match="row_ptr"):
        ctx.graph_upload(bad, col, W, NW)
    ctx.graph_upload(rp, col, W, NW)                          # the context is still good
    assert_bit_equal(ctx.forward(x, s), want, "after rejected uploads")


def test_errors_are_reported_not_fatal(ctx):
    with pytest.raises(capi.GvcError):
        ctx.forward(np.zeros(3, np.float32), 1.0)     # wrong length is caught on the Python side
    c = pkg.Context(0)
    with pytest.raises(capi.GvcError, match="no graph|no model"):
        c.forward_device(0, 1.0, 0)
    c.close()
    with pytest.raises(capi.GvcError, match="out of range"):
        pkg.Context(99)


def _checker_scores(g, scale, layers):
    """Scores of the CPU checker for a graph generated on the GPU: the compiled reference (oracle/_ref,
    one OpenBLAS thread, Prescott kernel -- the order exact mode reproduces) where it is built, else the
    C restatement.  Both finish a 16 M-edge graph in seconds."""
    eu, ev = g.edges_numpy()
    W = g.weights.cpu().numpy().view(np.uint32)
    x = W.astype(np.float32) / np.float32(scale)
    if po.REF_SO.exists():
        ref = po.Reference(threads=1)
        h = ref.model(po.layers_to_text(layers))
        gh = ref.graph_create(g.n, eu, ev, W)
        want = ref.predict_on(h, gh, x, scale)
        ref.graph_destroy(gh)
        ref.destroy(h)
        return want, "oracle/_ref"
    orc = po.Oracle()
    h = orc.parse(po.layers_to_text(layers))
    rp, col, W, NW = g.numpy()
    return orc.predict(h, rp, col, W, NW, x, scale)[:, 0], "oracle"


@pytest.mark.parametrize("maker", [
    lambda dev: graphs.rmat_graph(20, 16, seed=42, device=dev),     # BASELINE config 2: the graph bench.py times
    lambda dev: graphs.grid_graph(2001, 2003, device=dev),          # config 3 in small: 4 M vertices, odd count
], ids=["rmat_scale20", "grid_2001x2003"])
def test_bench_size_graphs_are_bit_exact(ctx, model_layers, maker):
    """The configurations the bench times, at size, against the CPU checker: exact mode bit for bit
    (every vertex, hubs of 64 452 neighbours included), fast mode within 1e-4 relative, identical
    `> 0.5` decisions away from ties, and determinism of repeated launches."""
    dev = torch.device("cuda:0")
    g = maker(dev)
    s = 200.0
    want, who = _checker_scores(g, s, model_layers)
    # x exactly as the checker got it: W / s in IEEE division (torch divides by a scalar on the GPU by
    # multiplying with its reciprocal, which differs in the last bit for some W)
    dx = torch.from_numpy(g.weights.cpu().numpy().view(np.uint32).astype(np.float32) / np.float32(s)).to(dev)
    ctx.graph_adopt(g.row_ptr.to(torch.int32).contiguous(), g.col, g.weights, g.nw)
    a = torch.empty(g.n, device=dev)
    b = torch.empty(g.n, device=dev)
    f = torch.empty(g.n, device=dev)
    torch.cuda.synchronize()
    ctx.forward_device(dx, s, a, pkg.MODE_EXACT)
    ctx.forward_device(dx, s, b, pkg.MODE_EXACT)
    ctx.forward_device(dx, s, f, pkg.MODE_FAST)
    ctx.sync()
    assert torch.equal(a, b)                                   # deterministic
    assert_bit_equal(a.cpu().numpy(), want, f"{g.name} vs {who}")
    fast = f.cpu().numpy()
    assert_rel_close(fast, want, FAST_RTOL, f"{g.name} fast vs {who}")
    clear = np.abs(want - 0.5) > 1e-4
    assert np.array_equal((fast > 0.5)[clear], (want > 0.5)[clear])


def test_full_size_grid_properties(ctx):
    """BASELINE config 3 at full size (4472 x 4472, 20 M vertices): the CPU checker needs minutes for it,
    so size-independent properties -- determinism, range, agreement of the two arithmetic modes -- and
    bit equality with the SAME vertices computed inside the 2001 x 2003 grid is covered above; here a
    translation property the domain offers: interior vertices of a grid with uniform weights all see the
    same neighbourhood, so their scores must be identical bit for bit."""
    dev = torch.device("cuda:0")
    side = 4472
    g = graphs.grid_graph(side, side, device=dev)
    s = 200.0
    ctx.graph_adopt(g.row_ptr.to(torch.int32).contiguous(), g.col, g.weights, g.nw)
    dx = (g.weights.to(torch.float32) / s).contiguous()
    a = torch.empty(g.n, device=dev)
    b = torch.empty(g.n, device=dev)
    f = torch.empty(g.n, device=dev)
    torch.cuda.synchronize()
    ctx.forward_device(dx, s, a, pkg.MODE_EXACT)
    ctx.forward_device(dx, s, b, pkg.MODE_EXACT)
    ctx.forward_device(dx, s, f, pkg.MODE_FAST)
    ctx.sync()
    assert torch.equal(a, b)
    assert bool(((a >= 0) & (a <= 1)).all()) and bool(torch.isfinite(a).all())
    rel = ((f - a).abs() / a.abs().clamp_min(1e-30)).max().item()
    assert rel < FAST_RTOL, rel
    # uniform weights: every vertex at distance >= 3 from the border has the same 3-hop neighbourhood
    wu = torch.full_like(g.weights, 7)
    nwu = (g.row_ptr[1:] - g.row_ptr[:-1]).to(torch.int32) * 7
    ctx.graph_adopt(g.row_ptr.to(torch.int32).contiguous(), g.col, wu, nwu)
    xu = (wu.to(torch.float32) / s).contiguous()
    torch.cuda.synchronize()                                   # xu is written on torch's stream, read on libgvc's
    ctx.forward_device(xu, s, a, pkg.MODE_EXACT)
    ctx.sync()
    inner = a.view(side, side)[3:-3, 3:-3]
    assert bool((inner.view(torch.int32) == inner.view(torch.int32)[0, 0]).all())


def test_multi_gpu_parity_when_several_gpus_are_visible():
    """On a box with >= 2 GPUs: N processes over NCCL == 1 GPU, bit for bit (tools/multi_gpu_check.py)."""
    import subprocess
    import sys
    from conftest import ROOT
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible; the sharding is covered by test_shards_are_bit_identical_to_one_gpu "
                    "and by tests/test_dist_cpu.py (gloo)")
    n = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        str(ROOT / "tools" / "multi_gpu_check.py"), "15"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_PARITY ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---- round 2 -----------------------------------------------------------------------------------------
def _scatter_into_span(rp, col, rng, slack=3):
    """The lists of a CSR placed the way a reduction_graph holds them after reductions: one edge array with
    holes (garbage between the lists), lists in arbitrary order, empty lists anywhere."""
    n = len(rp) - 1
    deg = np.diff(rp.astype(np.int64))
    order = rng.permutation(n)
    gaps = rng.integers(0, slack + 1, size=n)
    start = np.zeros(n, np.int64)
    pos = int(rng.integers(0, 5))
    for u in order:
        start[u] = pos
        pos += int(deg[u]) + int(gaps[u])
    span = rng.integers(0, 2 ** 32 - 1, size=max(pos, 1), dtype=np.uint64).astype(np.uint32)   # garbage ids in the holes
    for u in range(n):
        span[start[u]:start[u] + deg[u]] = col[int(rp[u]):int(rp[u + 1])]
    begin = start.astype(np.uint32)
    end = (start + deg).astype(np.uint32)
    empty = deg == 0
    begin[empty] = end[empty] = 0
    return span, begin, end


def test_streamed_upload_builds_the_same_graph(ctx, oracle, oracle_model):
    """the graph as 'edge span with holes + a range per vertex' (what begin(u)/end(u)
    expose), streamed through the pinned ring by several worker threads, compacted on the device ==
    the packed CSR uploaded the old way, bit for bit; several span chunks and vertex chunks."""
    rng = np.random.default_rng(8)
    z = np.load(GOLDEN / "reduced_graphs.npz")
    for key in ("er3000.r0", "er800_dense.r1", "grid40.r2"):
        rp, col, w, nw = z[key + ".row_ptr"], z[key + ".col"], z[key + ".w"], z[key + ".nw"]
        span, b, e = _scatter_into_span(rp, col, rng)
        ctx.graph_upload_ranges(span, b, e, w, nw, n_threads=2)
        x = w.astype(np.float32) / np.float32(200.0)
        assert_bit_equal(ctx.forward(x, 200.0), z[key + ".scores"], key)
    g = graphs.rmat_graph(16, 16, seed=9)                       # 65 536 vertices, ~2 M entries: many chunks
    rp, col, W, NW, x, s = inputs_of(g)
    ctx.graph_upload(rp, col, W, NW)
    want = ctx.forward(x, s)
    span, b, e = _scatter_into_span(rp, col, rng, slack=1)
    for threads in (1, 3, 0):
        ctx.graph_upload_ranges(span, b, e, W, NW, n_threads=threads)
        assert_bit_equal(ctx.forward(x, s), want, f"rmat16 streamed, {threads} threads")
    # the forward's input travelling with the graph (gvc_graph_upload_stream_x), forward(x = NULL) -- twice
    ctx.graph_upload_ranges(span, b, e, W, NW, n_threads=3, x=x)
    assert_bit_equal(ctx.forward(None, s), want, "rmat16 streamed with x")
    assert_bit_equal(ctx.forward(None, s), want, "rmat16 streamed with x, second forward")
    x2 = (x * np.float32(0.5)).astype(np.float32)
    want2 = ctx.forward(x2, s)                                    # an explicit x replaces it ...
    with pytest.raises(capi.GvcError, match="x is null"):         # ... and the resident one is gone
        ctx.forward(None, s)
    ctx.graph_upload_ranges(span, b, e, W, NW, x=x2)
    assert_bit_equal(ctx.forward(None, s), want2, "rmat16 streamed with another x")
    ctx.graph_upload_ranges(span, b, e, W, NW)                    # no x came with this graph
    with pytest.raises(capi.GvcError, match="x is null"):
        ctx.forward(None, s)
    g2 = graphs.rmat_graph(20, 8, seed=10)                        # > 32 MB: the 1 MB slots, overlapped schedule
    rp, col, W, NW, x, s = inputs_of(g2)
    ctx.graph_upload(rp, col, W, NW)
    want = ctx.forward(x, s)
    ctx.graph_upload_ranges(col, rp[:-1].astype(np.uint32), rp[1:].astype(np.uint32), W, NW, x=x)
    assert_bit_equal(ctx.forward(None, s), want, "rmat20 streamed with x, large slots")
    ctx.graph_upload_ranges(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32),
                            np.zeros(0, np.uint32), np.zeros(0, np.uint32))          # predict on an empty graph
    assert ctx.forward(np.zeros(0, np.float32), 200.0).size == 0


def test_streamed_upload_rejects_broken_ranges(ctx, vec):
    g, s, want = golden_graph(vec, "er607")
    rp, col, W, NW, x, _ = inputs_of(g, s)
    b, e = rp[:-1].astype(np.uint32), rp[1:].astype(np.uint32)
    bad = e.copy(); bad[7] = len(col) + 1                        # a range that ends past the span
    with pytest.raises(capi.GvcError, match="range"):
        ctx.graph_upload_ranges(col, b, bad, W, NW)
    bad = b.copy(); bad[9] = e[9] + 1                            # reversed
    with pytest.raises(capi.GvcError, match="range"):
        ctx.graph_upload_ranges(col, bad, e, W, NW)
    badc = col.copy(); badc[3] = g.n
    with pytest.raises(capi.GvcError, match="neighbour id"):
        ctx.graph_upload_ranges(badc, b, e, W, NW)
    with pytest.raises(capi.GvcError, match="no graph"):
        ctx.forward_device(0, s, 0)
    ctx.graph_upload_ranges(col, b, e, W, NW)
    assert_bit_equal(ctx.forward(x, s), want, "after rejected streams")


def test_dot_vs_golden_every_block_class(ctx):
    """dot() (src/matrix.cpp:106-122) bit for bit against the reference's outputs: OpenBLAS' row classes
    4/2/1 x column classes 8/4/2/1, both transposes, beta, k cut into blocks of 128."""
    d = np.load(GOLDEN / "dot_vectors.npz")
    names = sorted({k.split(".")[0] for k in d.files}, key=lambda s: int(s[1:]))
    for c in names:
        at, bt = (bool(v) for v in d[c + ".flags"])
        got = ctx.sgemm_host(d[c + ".A"], d[c + ".B"], d[c + ".C0"], trans_a=at, trans_b=bt, beta=float(d[c + ".beta"]))
        assert_bit_equal(got, d[c + ".out"], c)
    if po.REF_SO.exists():                                        # and live against the compiled reference
        ref = po.Reference(threads=1)
        rng = np.random.default_rng(12)
        for m, n, k in ((5, 7, 33), (66, 35, 32), (3, 1, 17), (1, 2, 48), (35, 16, 700)):
            A = (np.exp(rng.uniform(-5, 5, (m, k))) * rng.choice([-1, 1], (m, k))).astype(np.float32)
            B = (np.exp(rng.uniform(-5, 5, (k, n))) * rng.choice([-1, 1], (k, n))).astype(np.float32)
            assert_bit_equal(ctx.sgemm_host(A, B), ref.dot(A, B), f"live {m}x{n}x{k}")
            assert_bit_equal(ctx.sgemm_host(A.T.copy(), B.T.copy(), trans_a=True, trans_b=True),
                             ref.dot(A.T.copy(), B.T.copy(), at=True, bt=True), f"live T {m}x{n}x{k}")


def test_generic_linear_every_row_class(oracle):
    """The per-layer path's linear kernel uses the same block classes: a 4-column layer on 6 and 7 rows
    hits the 2-row and 1-row kernels of the 4-column class."""
    rng = np.random.default_rng(13)
    c = pkg.Context(0)
    for n in (4, 6, 7, 9, 10, 11):
        for K, N in ((19, 4), (7, 2), (33, 6), (5, 8), (40, 3), (16, 1)):
            x = rng.standard_normal((n, K)).astype(np.float32)
            Wm = rng.standard_normal((K, N)).astype(np.float32)
            b = rng.standard_normal(N).astype(np.float32)
            assert_bit_equal(c.linear_host(x, Wm, b), oracle.linear_forward(x, Wm, b), f"{n}x{K}x{N}")
    c.close()


def test_weight_scale_is_per_graph_layer(ctx, model_layers, oracle):
    """graph_layer::WEIGHT_SCALE belongs to each layer (src/gnn_inference.cpp:38-40): golden vector of a
    model whose three graph layers carry 20 / 200 / 57, on the fused path and on the per-layer path."""
    z = np.load(GOLDEN / "mixed_scales.npz")
    g = graphs.graph_from_edges(len(z["w"]), torch.from_numpy(z["eu"].astype(np.int64)),
                                torch.from_numpy(z["ev"].astype(np.int64)), torch.from_numpy(z["w"].astype(np.int64)))
    rp, col, W, NW = g.numpy()
    ctx.graph_upload(rp, col, W, NW)
    ctx.weight_scales(z["scales"])
    try:
        assert_bit_equal(ctx.forward(z["x"], 123.0), z["scores"], "mixed scales, fused")     # the scalar is ignored
    finally:
        ctx.weight_scales(None)
    assert not np.array_equal(ctx.forward(z["x"], 200.0), z["scores"])
    with pytest.raises(capi.GvcError, match="scales"):
        ctx.weight_scales([1.0, 2.0])
    # per-layer path: same model with an extra ReLU in front of the sigmoid (a no-op on the scores' bits
    # only if the pre-activation is positive -- so compare against the oracle instead)
    layers = list(model_layers[:-1]) + [(po.RELU, None, None), model_layers[-1]]
    c = pkg.Context(0)
    c.model_upload(layers)
    assert not c.fused
    c.graph_upload(rp, col, W, NW)
    c.weight_scales(z["scales"])
    h = oracle.parse(po.layers_to_text(layers))
    for i, s in enumerate(z["scales"]):
        oracle.set_graph_layer_scale(h, i, float(s))
    assert_bit_equal(c.forward(z["x"], 1.0), oracle.predict(h, rp, col, W, NW, z["x"])[:, 0], "mixed scales, per-layer path")
    c.close()


def test_failed_model_upload_leaves_no_model(model_layers):
    c = pkg.Context(0)
    c.model_upload(model_layers)
    assert c.fused
    bad = list(model_layers)
    bad[3] = (7, None, None)                                   # unknown layer kind: rejected before anything changes
    with pytest.raises(capi.GvcError, match="unknown kind"):
        c.model_upload(bad)
    assert c.fused                                             # the old model is still there and usable
    g = graphs.er_graph(100, 200, seed=3)
    rp, col, W, NW, x, s = inputs_of(g)
    c.graph_upload(rp, col, W, NW)
    assert np.isfinite(c.forward(x, s)).all()
    c.close()


def test_peer_lists_built_after_the_adjacency_landed(ctx, model_layers):
    """with gvc_peer_owners set BEFORE an upload from the pinned staging buffers, the
    per-vertex peer lists were read off an adjacency that was still in flight.  Owners first, then
    upload a DIFFERENT shard through the staging buffers, then check which rows were mirrored."""
    g0 = graphs.rmat_graph(13, 16, seed=64, n_limit=8190)
    parts = 2
    g, perm = graphs.balanced_relabel(g0, parts)
    rp, col, W, NW, x, s = inputs_of(g)
    per = g.n // parts
    dev = torch.device("cuda:0")
    dx = torch.from_numpy(x).to(dev)
    c = pkg.Context(0)
    c.model_upload(model_layers)
    # a first, different graph so that the device buffers hold stale ids of another adjacency
    g1 = graphs.er_graph(per, 3 * per, seed=5)
    r1 = inputs_of(g1)
    c.graph_upload(r1[0], r1[1], r1[2], r1[3], n_global=g.n, v_begin=0, v_end=per)
    mine = torch.full((g.n, 16), float("nan"), device=dev)
    other = torch.full((g.n, 16), float("nan"), device=dev)
    c.stage_peers(0, [other.data_ptr()])
    c.peer_owners([0, per, g.n], [-1, 0])                      # BEFORE the upload of the real shard
    a, b = 0, per
    srp, scol, sW, sNW = c.graph_staging(per, int(rp[b] - rp[a]))
    srp[:], scol[:], sW[:], sNW[:] = rp[a:b + 1] - rp[a], col[int(rp[a]):int(rp[b])], W[a:b], NW[a:b]
    c.graph_upload(srp, scol, sW, sNW, n_global=g.n, v_begin=a, v_end=b)
    torch.cuda.synchronize()
    c.stage_device(0, dx, mine, s)
    c.sync()
    src = np.repeat(np.arange(a, b), np.diff(rp[a:b + 1].astype(np.int64)))
    nb = col[int(rp[a]):int(rp[b])]
    must = np.zeros(g.n, bool)
    must[src[nb >= per]] = True                                # rows of shard 0 with a neighbour in shard 1
    arrived = ~torch.isnan(other).any(1).cpu().numpy()
    assert np.array_equal(arrived, must), (arrived.sum(), must.sum())
    c.close()


def _star(deg, extra=0, seed=3):
    """vertex 0 with `deg` neighbours (+ `extra` random edges among the leaves)"""
    n = deg + 1
    eu = [torch.zeros(deg, dtype=torch.int64)]
    ev = [torch.arange(1, deg + 1, dtype=torch.int64)]
    if extra:
        r = torch.from_numpy(np.random.default_rng(seed).integers(1, n, size=(extra, 2)))
        eu.append(r[:, 0]); ev.append(r[:, 1])
    a, b = graphs._canonical_edges(torch.cat(eu), torch.cat(ev), n)
    return graphs.graph_from_edges(n, a, b, graphs.random_weights(n, seed), name=f"star{deg}")


@pytest.mark.parametrize("deg", [16_384, 40_000, 262_144])
def test_parallel_exact_hub_sums(ctx, oracle, oracle_model, deg):
    """Vertices of degree >= 16 384 take the parallel emulation of the sequential sum (gvc_px.cuh):
    bit-equal to the oracle's element-wise chain, shuffled adjacency, all three stages."""
    g = _star(deg, extra=3 * deg)
    rp, col, W, NW, x, s = inputs_of(g)
    col = col.copy()
    np.random.default_rng(deg).shuffle(col[int(rp[0]):int(rp[1])])
    want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
    ctx.graph_upload(rp, col, W, NW)
    assert_bit_equal(ctx.forward(x, s, pkg.MODE_EXACT), want, g.name)
    assert_rel_close(ctx.forward(x, s, pkg.MODE_FAST), want, FAST_RTOL, g.name + " fast")


def test_parallel_exact_hub_sums_survive_hostile_inputs(ctx, oracle, oracle_model):
    """predict must use the `in` it is given: negative, zero, huge, denormal and NaN inputs under a hub
    (stage 0 sees them raw).  The fast way's preconditions fail and the batches fall back to the chain."""
    g = _star(50_000, extra=20_000)
    rp, col, W, NW, _, s = inputs_of(g)
    rng = np.random.default_rng(4)
    base = rng.random(g.n).astype(np.float32)
    variants = {
        "negatives": (rng.standard_normal(g.n)).astype(np.float32),
        "zeros and ties": (rng.integers(0, 4, g.n) * 0.25).astype(np.float32),
        "outliers": np.where(rng.random(g.n) < 0.001, 3e7, base).astype(np.float32),
        "denormals": (base * 1e-39).astype(np.float32),
        "wide": np.exp(rng.uniform(-30, 10, g.n)).astype(np.float32),
    }
    ctx.graph_upload(rp, col, W, NW)
    for name, x in variants.items():
        want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
        assert_bit_equal(ctx.forward(x, s, pkg.MODE_EXACT), want, name)
    x = base.copy(); x[col[int(rp[0]) + 777]] = np.nan          # one NaN among the hub's neighbours
    want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
    got = ctx.forward(x, s, pkg.MODE_EXACT)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert_bit_equal(got[ok], want[ok], "NaN under the hub")


def test_group_of_devices_in_one_process(ctx, model_layers):
    """gvc_group (SURVEY 8(e) behind one call): the forward sharded over several contexts of THIS process,
    rows exchanged by the stage kernels' peer stores, stages separated by CUDA events.  Two and three
    shards on one GPU (devices may repeat) and, where a second GPU is visible, two real devices:
    bit-identical to the one-context forward, odd vertex count (the exact-mode tail vertex), hubs."""
    cases = [graphs.rmat_graph(13, 16, seed=71, n_limit=8191), graphs.er_graph(5000, 20000, seed=72),
             _star(40_000, extra=50_000)]
    layouts = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        layouts.append([0, 1])
    if torch.cuda.device_count() >= 4:
        layouts.append([0, 1, 2, 3])
    for devices in layouts:
        grp = capi.Group(devices)
        grp.model_upload(model_layers)
        for g in cases:
            rp, col, W, NW, x, s = inputs_of(g)
            ctx.graph_upload(rp, col, W, NW)
            want = ctx.forward(x, s)
            want_fast = ctx.forward(x, s, pkg.MODE_FAST)
            grp.graph_upload(rp, col, W, NW)
            b = grp.bounds
            assert b[0] == 0 and b[-1] == g.n and all(b[i] <= b[i + 1] for i in range(len(b) - 1))
            for _ in range(2):                                   # twice: buffers of the previous forward are reused
                assert_bit_equal(grp.forward(x, s), want, f"{g.name} on devices {devices}")
            assert_rel_close(grp.forward(x, s, pkg.MODE_FAST), want_fast, FAST_RTOL, f"{g.name} fast on {devices}")
        grp.close()
