"""CPU: the synthetic-graph generators produce what the reference's parser + graph ctor would."""
import hashlib

import numpy as np
import pytest
import torch

import gnn_mwvc_b200  # noqa: F401
from gnn_mwvc_b200 import graphs


def check_csr(g):
    rp, col, W, NW = g.numpy()
    assert rp[0] == 0 and rp[-1] == len(col) == 2 * g.n_edges
    assert np.all(np.diff(rp.astype(np.int64)) >= 0)
    src = np.repeat(np.arange(g.n), np.diff(rp.astype(np.int64)))
    assert np.all(col.astype(np.int64) != src)                                   # no self loops
    key = src.astype(np.int64) * g.n + col
    assert np.all(np.diff(key) > 0)                                               # ascending, unique
    assert set(zip(src.tolist(), col.tolist())) == set(zip(col.tolist(), src.tolist()))   # symmetric
    nw = np.zeros(g.n, np.int64)
    np.add.at(nw, src, W[col].astype(np.int64))
    assert np.array_equal(nw, NW.astype(np.int64))                                # NW = sum of neighbour weights
    assert W.min() >= 1 and W.max() <= 200


def test_er_exact_edge_count():
    g = graphs.er_graph(500, 2000, seed=1)
    assert g.n == 500 and g.n_edges == 2000
    check_csr(g)


def test_rmat_and_limit():
    g = graphs.rmat_graph(10, 8, seed=2)
    assert g.n == 1024
    check_csr(g)
    g2 = graphs.rmat_graph(10, 8, seed=2, n_limit=777)
    assert g2.n == 777 and int(g2.col.max()) < 777
    check_csr(g2)
    deg = np.diff(g.numpy()[0].astype(np.int64))
    assert deg.max() > 8 * deg.mean()          # power-law skew


def test_grid():
    g = graphs.grid_graph(7, 9)
    assert g.n == 63 and g.n_edges == 7 * 8 + 6 * 9
    check_csr(g)
    assert np.diff(g.numpy()[0].astype(np.int64)).max() == 4


def test_er10k_fixture_file(tmp_path):
    # the METIS file of config 1 is reproducible byte for byte
    g = graphs.er10k_fixture()
    p = tmp_path / "er10k.graph"
    graphs.write_metis(g, p)
    assert hashlib.md5(p.read_bytes()).hexdigest() == "f00c00870deedf9c9c15d20a8f92eba1"


def test_nnz_balanced_ranges():
    g = graphs.rmat_graph(12, 16, seed=3)
    for parts in (2, 4, 8):
        b = graphs.nnz_balanced_ranges(g.row_ptr, parts)
        assert b[0] == 0 and b[-1] == g.n and len(b) == parts + 1 and b == sorted(b)
        assert all(x % 32 == 0 for x in b[:-1])
        nnz = [int(g.row_ptr[b[i + 1]] - g.row_ptr[b[i]]) for i in range(parts)]
        cost = [nnz[i] + 16 * (b[i + 1] - b[i]) for i in range(parts)]
        assert max(cost) < 1.35 * (sum(cost) / parts)


@pytest.mark.parametrize("fn", ["cyclic_relabel", "balanced_relabel"])
def test_relabel_keeps_lists_and_balances(fn):
    g = graphs.rmat_graph(12, 16, seed=5, n_limit=4093)
    parts = 4
    h, perm = getattr(graphs, fn)(g, parts)
    per = -(-g.n // parts)
    assert h.n == per * parts and h.nnz == g.nnz
    assert torch.equal(torch.sort(perm).values.unique(), torch.sort(perm).values)        # one-to-one
    rp, col = g.row_ptr.numpy(), g.col.numpy().view(np.uint32)
    hrp, hcol = h.row_ptr.numpy(), h.col.numpy().view(np.uint32)
    pm = perm.numpy()
    for v in (0, 1, 17, g.n - 1, int(np.argmax(np.diff(rp)))):                          # lists: same order, new names
        want = pm[col[rp[v]:rp[v + 1]]]
        got = hcol[hrp[pm[v]]:hrp[pm[v] + 1]]
        assert np.array_equal(got, want)
        assert h.weights[pm[v]] == g.weights[v] and h.nw[pm[v]] == g.nw[v]
    pad = np.setdiff1d(np.arange(h.n), pm)
    assert all(hrp[p + 1] == hrp[p] for p in pad)                                       # padding is isolated
    shard_nnz = [int(hrp[(r + 1) * per] - hrp[r * per]) for r in range(parts)]
    if fn == "balanced_relabel":
        assert max(shard_nnz) - min(shard_nnz) <= int(np.diff(rp).max())                # equal work up to one hub
    else:
        assert max(shard_nnz) > 1.5 * min(shard_nnz)      # R-MAT: ids with zero low bits are the heavy ones


def test_live_rows_are_the_non_isolated_prefix():
    g = graphs.rmat_graph(11, 8, seed=3)
    h, _ = graphs.balanced_relabel(g, 4)
    per = h.n // 4
    bounds = [r * per for r in range(5)]
    live = graphs.live_rows(h.row_ptr, bounds)
    deg = (h.row_ptr[1:] - h.row_ptr[:-1]).numpy()
    for r in range(4):
        d = deg[bounds[r]:bounds[r + 1]]
        assert (d[:live[r]] > 0).all() and (d[live[r]:] == 0).all()     # sorted by degree: live rows are a prefix
    assert max(live) - min(live) <= 1 and 0 < max(live) < per
    assert graphs.live_rows(torch.zeros(9, dtype=torch.int64), [0, 4, 8]) == [0, 0]


def test_u32_bit_patterns():
    t = torch.tensor([0, 1, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1])
    assert graphs.to_u32(t).numpy().view(np.uint32).tolist() == t.tolist()
