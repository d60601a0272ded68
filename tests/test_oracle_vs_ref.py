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


def test_dot_any_shape(reference, oracle):
    """dot() for every OpenBLAS block class (rows 4/2/1 x columns 8/4/2/1), transposes, beta, and k long
    enough to be cut into blocks of 128: the restatement equals the reference bit for bit."""
    rng = np.random.default_rng(5)

    def wide(shape):
        return (np.exp(rng.uniform(-6, 6, shape)) * rng.choice([-1, 1], shape)).astype(np.float32)
    for m in (1, 2, 3, 4, 5, 6, 7, 9, 64, 101):
        for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 35):
            for k in (1, 7, 16, 35, 47, 300):
                at, bt = bool(rng.integers(2)), bool(rng.integers(2))
                beta = float(rng.choice([0.0, 1.0, -0.75]))
                A, B, C0 = wide((m, k)), wide((k, n)), wide((m, n))
                As = A.T.copy() if at else A
                Bs = B.T.copy() if bt else B
                assert_bit_equal(oracle.dot(As, Bs, C0, at, bt, beta), reference.dot(As, Bs, C0, at, bt, beta),
                                 f"{m}x{n}x{k} at={at} bt={bt} beta={beta}")
    A, B = wide((35, 2500)), wide((2500, 32))                   # the weight-gradient shape of old_files' training
    assert_bit_equal(oracle.dot(A, B), reference.dot(A, B), "k = 2500")


def test_reduction_mutators_keep_the_oracle_in_step(reference, oracle, oracle_model, model_layers):
    """predict on graphs that real reduction_graph mutators produced (fresh random scripts, not the
    committed ones): restatement == reference."""
    import random
    hr = reference.model(po.layers_to_text(model_layers))
    g = graphs.er_graph(2500, 9000, seed=77)
    eu, ev = g.edges_numpy()
    gh = reference.graph_create(g.n, eu, ev, g.numpy()[2])
    rnd = random.Random(3)
    for _ in range(4):
        for _ in range(300):
            reference.graph_mutate(gh, rnd.choice([0, 1, 2, 2, 4]), rnd.randrange(reference.graph_size(gh)))
        reference.graph_mutate(gh, reference.RELABEL)
        rp, col, w, nw, act = reference.graph_csr(gh)
        assert act.all()
        x = w.astype(np.float32) / np.float32(200.0)
        want = reference.predict_on(hr, gh, x, 200.0)
        if reference.graph_size(gh):
            assert_bit_equal(oracle.predict(oracle_model, rp, col, w, nw, x, 200.0)[:, 0], want, "after reductions")
    reference.graph_destroy(gh)
