"""CPU: the HOST half of the drop-in's predict -- reading the reduction_graph through its public
accessors and describing it to gvc_graph_upload_stream (edge span + ranges, or gathered lists) -- against
the reference's own view of the same graph, with libgvc replaced by a mock that only records what it
is given (tests/mock_gvc.cpp).  Graphs are mutated through the reference's real mutators between calls,
as the solver does."""
import ctypes as C
import random
import subprocess
from pathlib import Path

import numpy as np
import pytest

import gnn_mwvc_b200  # noqa: F401
from gnn_mwvc_b200 import graphs
from oracle import pyoracle as po

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
SO = ROOT / "tests" / "_build" / "libgnndropin_mock.so"


@pytest.fixture(scope="module")
def mock():
    if not (REF / "include" / "gnn_inference.hpp").exists():
        if not SO.exists():
            pytest.skip("needs the reference's headers (build container) to build the mock harness")
    else:
        SO.parent.mkdir(exist_ok=True)
        host = ROOT / "gnn-mwvc_b200" / "host"
        srcs = [ROOT / "oracle" / "ref_harness.cpp", host / "gvc_gnn_inference.cpp", host / "gvc_matrix.cpp",
                ROOT / "tests" / "mock_gvc.cpp"]
        if not SO.exists() or any(s.stat().st_mtime > SO.stat().st_mtime for s in srcs):
            subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-DNDEBUG", "-fPIC", "-shared",
                                   "-DGVC_HARNESS_DROPIN", "-I", str(REF / "include"), "-I", str(ROOT / "include"),
                                   "-I", str(host), "-o", str(SO), *map(str, srcs), "-pthread"])
    ours = po.Reference(so=SO)
    ours.L.mock_last_graph.restype = C.c_uint64
    ours.L.mock_last_graph.argtypes = [po._u64p, po._u32p, po._u32p, po._u32p, po._u64p, C.POINTER(C.c_int)]
    return ours


def last_graph(ours, n):
    span, streamed = C.c_uint64(), C.c_int()
    nnz = int(ours.L.mock_last_graph(None, None, None, None, C.byref(span), C.byref(streamed)))
    rp, col = np.empty(n + 1, np.uint64), np.empty(max(nnz, 1), np.uint32)
    w, nw = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint32)
    ours.L.mock_last_graph(po._p(rp, po._u64p), po._p(col, po._u32p), po._p(w, po._u32p), po._p(nw, po._u32p), None, None)
    return rp, col[:nnz], w[:n], nw[:n], int(span.value), bool(streamed.value)


@pytest.mark.parametrize("maker,steps", [
    (lambda: graphs.er_graph(4000, 12000, seed=41), 400),
    (lambda: graphs.rmat_graph(11, 8, seed=42), 150),
    (lambda: graphs.grid_graph(30, 33, seed=43), 200),
    (lambda: graphs.er_graph(150_000, 450_000, seed=44), 3000),      # large enough for the multi-threaded passes
])
def test_predict_hands_over_exactly_the_graph_the_reference_sees(mock, reference, model_layers, maker, steps):
    ours = mock
    g = maker()
    eu, ev = g.edges_numpy()
    W = g.numpy()[2]
    text = po.layers_to_text(model_layers)
    hr, ho = reference.model(text), ours.model(text)
    gr, go = reference.graph_create(g.n, eu, ev, W), ours.graph_create(g.n, eu, ev, W)
    rnd = random.Random(9)
    modes = set()
    for rounds in range(5):
        n = reference.graph_size(gr)
        assert ours.graph_size(go) == n
        x = (np.arange(n, dtype=np.float32) * np.float32(0.25) + np.float32(rounds)).astype(np.float32)
        out = ours.predict_on(ho, go, x, 200.0)
        assert out.shape == (n,) and (n == 0 or np.all(out == 0.5))         # the mock's "scores" made it into `out`
        if n:                                                                # predict's input travelled with the graph
            ours.L.mock_last_x.restype = C.c_uint64
            ours.L.mock_last_x.argtypes = [po._f32p]
            got = np.empty(n, np.float32)
            assert ours.L.mock_last_x(po._p(got, po._f32p)) == n and np.array_equal(got, x)
        rp, col, w, nw, act = reference.graph_csr(gr)
        rp2, col2, w2, nw2, span, streamed = last_graph(ours, n)
        assert streamed
        assert np.array_equal(rp, rp2) and np.array_equal(col, col2), f"round {rounds}: adjacency differs"
        assert np.array_equal(w, w2) and np.array_equal(nw, nw2)
        modes.add("gathered" if span == len(col) and rounds else "span")
        for _ in range(steps):                                               # reductions, then relabel, as gnn_solve does
            op = rnd.choice([0, 1, 2, 2, 4, 0])
            u = rnd.randrange(max(reference.graph_size(gr), 1))
            a, b = reference.graph_mutate(gr, op, u), ours.graph_mutate(go, op, u)
            assert a == b
        reference.graph_mutate(gr, reference.RELABEL)
        ours.graph_mutate(go, ours.RELABEL)
    assert modes == {"span", "gathered"}          # both presentations of the graph were exercised
    reference.graph_destroy(gr)
    ours.graph_destroy(go)


def test_mixed_weight_scales_reach_the_library(mock, model_layers):
    ours = mock
    ours.L.mock_last_scales.argtypes = [po._f32p]
    it = iter([20.0, 200.0, 57.0])
    hm = ours.model_build(model_layers, [next(it) if k == po.GRAPH else 0.0 for k, _, _ in model_layers])
    g = graphs.er_graph(50, 100, seed=1)
    eu, ev = g.edges_numpy()
    gh = ours.graph_create(g.n, eu, ev, g.numpy()[2])
    ours.predict_on_as_is(hm, gh, np.ones(g.n, np.float32))
    got = np.zeros(8, np.float32)
    assert ours.L.mock_last_scales(po._p(got, po._f32p)) == 3 and got[:3].tolist() == [20.0, 200.0, 57.0]
    ours.L.ref_model_set_weight_scale(hm, 99.0)                                # set_weight_scale makes them uniform again
    ours.predict_on_as_is(hm, gh, np.ones(g.n, np.float32))
    assert ours.L.mock_last_scales(po._p(got, po._f32p)) == 0
    ours.graph_destroy(gh)
