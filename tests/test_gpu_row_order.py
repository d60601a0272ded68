"""GPU: rows in schedule order (gvc_api.cu relabel_rows).  Whole-graph contexts of 500 000 vertices and more run the
fused path on an internal copy of the graph whose vertices are renumbered by the degree schedule (L2 then holds
the rows of the hubs without their cold neighbours).  Nothing a caller sees may change: x, scores and selection
keys stay in the caller's numbering and every bit stays the same.  GVC_ROW_ORDER_MIN_VERTICES=0 forces the
renumbering on the small graphs the oracle can check."""
import os

import numpy as np
import pytest
import torch

import gnn_mwvc_b200 as pkg
from gnn_mwvc_b200 import capi, graphs
from conftest import GOLDEN
from helpers import assert_bit_equal, assert_rel_close, golden_graph, golden_names, inputs_of, oracle_stages

pytestmark = pytest.mark.gpu


def _setenv(**kw):
    old = {k: os.environ.get(k) for k in kw}
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return old


@pytest.fixture()
def forced():
    """renumber every graph, at upload already"""
    old = _setenv(GVC_ROW_ORDER_MIN_VERTICES="0", GVC_ROW_ORDER_AFTER="0")
    yield
    _setenv(**old)


@pytest.fixture(scope="module")
def ctx(model_layers):
    c = pkg.Context(0)
    c.model_upload(model_layers)
    yield c
    c.close()


def test_golden_vectors_with_renumbered_rows(forced, ctx):
    vec = np.load(GOLDEN / "predict_vectors.npz")
    for name in golden_names(vec):
        g, s, want = golden_graph(vec, name)                 # includes odd vertex counts: the 1-row kernel's vertex moves
        rp, col, W, NW, x, _ = inputs_of(g, s)
        ctx.graph_upload(rp, col, W, NW)
        assert_bit_equal(ctx.forward(x, s), np.asarray(want).reshape(-1), name)
        assert_rel_close(ctx.forward(x, s, pkg.MODE_FAST), np.asarray(want).reshape(-1), 1e-4, name + " fast")


def test_every_gather_class_stage_outputs_and_keys(forced, ctx, oracle, oracle_model, model_layers):
    g = graphs.rmat_graph(13, 16, seed=77, n_limit=8191)      # odd n, hubs, isolated vertices
    rp, col, W, NW, x, s = inputs_of(g)
    h1, h2, scores = oracle_stages(oracle, model_layers, rp, col, W, NW, x, s)
    ctx.graph_upload(rp, col, W, NW)
    rows = ctx.row_order()
    assert sorted(rows.tolist()) == list(range(g.n))
    deg = np.diff(rp)
    assert deg[rows[0]] >= 0.8 * deg.max() and deg[rows[-1]] == deg.min()          # hubs first (by degree bin)
    got, keys, side = ctx.forward_keys(x, s)
    assert_bit_equal(got, scores, "scores")
    assert_bit_equal(keys, np.minimum(scores, np.float32(1.0) - scores), "keys")
    assert np.array_equal(side.astype(bool), scores > 0.5)
    dev = torch.device("cuda:0")
    dx = torch.from_numpy(x).to(dev)
    d1, d2, ds = torch.empty(g.n, 16, device=dev), torch.empty(g.n, 16, device=dev), torch.empty(g.n, device=dev)
    torch.cuda.synchronize()
    for st, (a, b) in enumerate(((dx, d1), (d1, d2), (d2, ds))):
        ctx.stage_device(st, a, b, s)
    ctx.sync()
    u1, u2 = np.empty_like(h1), np.empty_like(h2)
    u1[rows], u2[rows] = d1.cpu().numpy(), d2.cpu().numpy()
    assert_bit_equal(u1, h1, "h1 rows, un-permuted")
    assert_bit_equal(u2, h2, "h2 rows, un-permuted")
    assert_bit_equal(ds.cpu().numpy(), scores, "scores through the stage calls")


def test_huge_hub_and_streamed_upload_with_x(forced, ctx, oracle, oracle_model):
    # a 40 000-neighbour hub (parallel exact sums) next to low-degree vertices, uploaded as span + ranges with x
    n = 40_001 + 3000
    eu = [np.zeros(40_000, np.int64)]
    ev = [np.arange(1, 40_001, dtype=np.int64)]
    rng = np.random.default_rng(5)
    a = rng.integers(40_001, n, 6000)
    b = rng.integers(1, n, 6000)
    keep = a != b
    eu.append(np.minimum(a, b)[keep]); ev.append(np.maximum(a, b)[keep])
    w = rng.integers(1, 201, n)
    g = graphs.graph_from_edges(n, torch.from_numpy(np.concatenate(eu)), torch.from_numpy(np.concatenate(ev)), torch.from_numpy(w), name="hub")
    rp, col, W, NW, x, s = inputs_of(g)
    want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
    ctx.graph_upload_ranges(col, rp[:-1].astype(np.uint32), rp[1:].astype(np.uint32), W, NW, n_threads=3, x=x)
    assert ctx.row_order()[0] == 0                              # the hub is vertex 0 and row 0
    assert_bit_equal(ctx.forward(None, s), want, "hub graph, streamed, x with the graph")
    assert_bit_equal(ctx.forward(x, s), want, "hub graph, explicit x")


def test_rows_are_reordered_when_a_graph_is_forwarded_again(ctx, oracle, oracle_model):
    """the default policy: not at upload (a solver's predict runs one forward per graph), but at the start of the
    second forward on the same graph; a new upload starts over"""
    old = _setenv(GVC_ROW_ORDER_MIN_VERTICES="0", GVC_ROW_ORDER_AFTER=None)
    try:
        g = graphs.rmat_graph(12, 16, seed=3, n_limit=4095)
        rp, col, W, NW, x, s = inputs_of(g)
        want = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
        ident = np.arange(g.n)
        ctx.graph_upload(rp, col, W, NW)
        assert np.array_equal(ctx.row_order(), ident)
        assert_bit_equal(ctx.forward(x, s), want, "first forward")
        assert np.array_equal(ctx.row_order(), ident)
        assert_bit_equal(ctx.forward(x, s), want, "second forward")
        assert not np.array_equal(ctx.row_order(), ident)
        assert_bit_equal(ctx.forward(x, s, pkg.MODE_EXACT), want, "third forward")
        ctx.graph_upload(rp, col, W, NW)
        assert np.array_equal(ctx.row_order(), ident)
    finally:
        _setenv(**old)


def test_default_threshold_leaves_small_graphs_alone(ctx):
    old = _setenv(GVC_ROW_ORDER_MIN_VERTICES=None, GVC_ROW_ORDER_AFTER="0")
    try:
        g = graphs.er_graph(5000, 20000, seed=2)
        rp, col, W, NW, x, s = inputs_of(g)
        ctx.graph_upload(rp, col, W, NW)
        ctx.forward(x, s)
        ctx.forward(x, s)
        assert np.array_equal(ctx.row_order(), np.arange(g.n))
    finally:
        _setenv(**old)


def test_bench_graph_same_bits_before_and_after_the_reordering(ctx):
    """R-MAT scale 20 (what bench.py times; its first forward is checked against the compiled reference in
    test_gpu_parity.py): the forwards after the rows were reordered return the very same scores, both modes"""
    old = _setenv(GVC_ROW_ORDER_MIN_VERTICES=None, GVC_ROW_ORDER_AFTER=None)
    try:
        g = graphs.rmat_graph(20, 16, seed=42)
        rp, col, W, NW, x, s = inputs_of(g)
        ctx.graph_upload(rp, col, W, NW)
        first = ctx.forward(x, s)
        assert np.array_equal(ctx.row_order(), np.arange(g.n))
        second, keys, side = ctx.forward_keys(x, s)
        rows = ctx.row_order()
        assert not np.array_equal(rows, np.arange(g.n)) and np.array_equal(np.sort(rows), np.arange(g.n))
        assert_bit_equal(second, first, "second forward (rows reordered)")
        assert_bit_equal(keys, np.minimum(first, np.float32(1.0) - first), "keys")
        assert np.array_equal(side.astype(bool), first > 0.5)
        assert_bit_equal(ctx.forward(x, s), first, "third forward")
        fast = ctx.forward(x, s, pkg.MODE_FAST)
        assert_rel_close(fast, first, 1e-4, "fast mode on reordered rows")
    finally:
        _setenv(**old)
