"""CPU: pin the C restatement against the UNMODIFIED reference compiled into oracle/_ref
(reference sources + OpenBLAS 0.3.15 Prescott, one thread).  Skipped where oracle/_ref
was never built; the committed golden vectors carry the same pin to the GPU box."""
import numpy as np

import gnn_mwvc_b200  # noqa: F401
from gnn_mwvc_b200 import graphs
from helpers import assert_bit_equal, inputs_of
from oracle import pyoracle as po


def test_blas_is_pinned(reference):
    cfg = reference.blas_config()
    assert "0.3.15" in cfg and "Prescott" in cfg


def test_forward_bit_exact(reference, oracle, oracle_model, model_layers):
    hr = reference.model(po.layers_to_text(model_layers))
    for g in [graphs.er_graph(4001, 20000, seed=21), graphs.er_graph(5000, 9000, seed=22),
              graphs.rmat_graph(12, 16, seed=23), graphs.grid_graph(37, 53, seed=24),
              graphs.er_graph(7, 9, seed=25), graphs.er_graph(1, 0, seed=26)]:
        rp, col, W, NW, x, s = inputs_of(g)
        eu, ev = g.edges_numpy()
        want, (rp2, col2, nw2) = reference.predict(hr, g.n, eu, ev, W, x, s, want_csr=True)
        # the generators build the very CSR the reduction_graph ctor builds
        assert np.array_equal(rp, rp2) and np.array_equal(col, col2) and np.array_equal(NW, nw2)
        got = oracle.predict(oracle_model, rp, col, W, NW, x, s)[:, 0]
        assert_bit_equal(got, want, g.name)


def test_linear_remainder_kernels(reference, oracle):
    # OpenBLAS takes rows 4/2/1 at a time; the 1-row and 1-column kernels sum differently
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 33, 1001):
        for K, N in ((5, 32), (32, 32), (35, 32), (32, 16), (16, 1)):
            x = rng.standard_normal((n, K)).astype(np.float32)
            Wm = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
            b = rng.standard_normal(N).astype(np.float32)
            assert_bit_equal(oracle.linear_forward(x, Wm, b), reference.linear_layer(x, Wm, b), f"{n}x{K}x{N}")


def test_linear_init_matches(reference):
    # linear_layer's random init (reference src/gnn_inference.cpp:7-18) is std::mt19937 +
    # uniform_real_distribution<float>: deterministic across builds of libstdc++
    W, b = reference.linear_init(5, 32, 7)
    lim = 1.0 / np.sqrt(6.0)
    assert W.shape == (5, 32) and np.all(np.abs(W) <= lim) and np.all(np.abs(b) <= lim)
