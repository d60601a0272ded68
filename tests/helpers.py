"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

import gnn_mwvc_b200  # noqa: F401
from gnn_mwvc_b200 import graphs
from oracle import pyoracle as po


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(got, want, what=""):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    bad = np.nonzero(bits(got).ravel() != bits(want).ravel())[0]
    assert bad.size == 0, (f"{what}: {bad.size}/{got.size} values differ, first at {bad[0]}: "
                           f"{got.ravel()[bad[0]]!r} vs {want.ravel()[bad[0]]!r}")


def assert_rel_close(got, want, tol, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    assert rel.max(initial=0.0) <= tol, f"{what}: max relative error {rel.max():.3e} > {tol}"


def golden_graph(vec, name):
    eu = torch.from_numpy(vec[f"{name}.eu"].astype(np.int64))
    ev = torch.from_numpy(vec[f"{name}.ev"].astype(np.int64))
    w = torch.from_numpy(vec[f"{name}.w"].astype(np.int64))
    g = graphs.graph_from_edges(len(w), eu, ev, w, name=name)
    return g, float(vec[f"{name}.scale"]), vec[f"{name}.scores"]


def golden_names(vec):
    return sorted({k.split(".")[0] for k in vec.files})


def inputs_of(g, scale=None):
    row_ptr, col, W, NW = g.numpy()
    s = float(W.max()) if scale is None else scale
    x = W.astype(np.float32) / np.float32(s)
    return row_ptr, col, W, NW, x, s


def oracle_stages(orc, layers, row_ptr, col, W, NW, x, scale):
    """(h1, h2, scores) of the GNN_VC architecture computed layer by layer with the oracle."""
    a = np.ascontiguousarray(x, np.float32).reshape(-1, 1)
    outs = []
    n = a.shape[0]
    for kind, Wm, b in layers:
        if kind == po.GRAPH:
            if a.shape[1] == 16:
                outs.append(a.copy())
            a = orc.graph_forward(row_ptr, col, W, NW, scale, a)
        elif kind == po.LINEAR:
            a = orc.linear_forward(a, Wm, b)
        elif kind == po.RELU:
            a = orc.relu(a)
        else:
            a = orc.sigmoid(a)
    assert len(outs) == 2 and a.shape == (n, 1)
    return outs[0], outs[1], a[:, 0]
