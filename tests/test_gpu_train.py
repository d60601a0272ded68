"""GPU: the TRAINING path against what the reference computed
(tests/golden/train_vectors.npz, written by the unmodified old_files/src/lib/gnn_training.cpp through
oracle/train_harness.cpp; tools/make_golden_train.py).

  * through the C ABI (gvc_trainer_*): exact mode must reproduce predict, the input gradient and every
    accumulated weight / bias gradient BIT FOR BIT (the three dot() calls of a linear layer run in the order
    of the OpenBLAS kernel the checker is pinned to), fast mode within 1e-4 of the gradient's scale;
  * through the reference's own C++ interface (the same harness built over the drop-in host units), in a
    child process because that interface reports CUDA problems by aborting;
  * the single-layer host entry points against the live reference where oracle/_ref travelled."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import gnn_mwvc_b200 as pkg
from gnn_mwvc_b200 import capi, graphs
from conftest import GOLDEN
from helpers import assert_bit_equal, inputs_of
from oracle import pyoracle as po
from test_train_cpu import CASES, csr_of, golden_layers, rel_to_scale

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def ctx():
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def z():
    return np.load(GOLDEN / "train_vectors.npz")


def run_sequence(tr, z, name, mode):
    """the sequence tools/make_golden_train.py recorded; returns what it read back"""
    x, y = z[f"{name}.x"], z[f"{name}.y"]
    scale = float(z[f"{name}.scale"])
    got = {"out": tr.predict(x, scale, mode)}
    got["loss"] = tr.mse_backprop(y, mode)
    tr.predict(x, scale, mode, want_out=False)
    tr.mse_backprop(y, mode)
    tr.predict(x, scale, mode)
    got["grad_x"] = tr.backprop(z[f"{name}.g_dir"], mode)
    for i, (k, _, _) in enumerate(tr.layers):
        if k == capi.LINEAR:
            got[f"gW{i}"], got[f"gb{i}"] = tr.read(1, i)
    tr.sgd_step(3 * len(x), lr=0.05, momentum=0.9, weight_decay=0.001)
    tr.zero_grad()
    for i, (k, _, _) in enumerate(tr.layers):
        if k == capi.LINEAR:
            got[f"W1_{i}"], got[f"b1_{i}"] = tr.read(0, i)
            got[f"vW{i}"], got[f"vb{i}"] = tr.read(2, i)
            zg, zb = tr.read(1, i)
            assert not zg.any() and not zb.any(), "zero_grad left something"
    got["out1"] = tr.predict(x, scale, mode)
    return got


@pytest.mark.parametrize("name", CASES)
def test_trainer_exact_mode_reproduces_the_reference_bit_for_bit(ctx, z, name):
    layers = golden_layers(z, name)
    row_ptr, col, W, NW = csr_of(z, name)
    ctx.graph_upload(row_ptr, col, W, NW)
    tr = capi.Trainer(ctx, layers)
    assert (tr.in_w, tr.out_w) == (1, 1)
    got = run_sequence(tr, z, name, pkg.MODE_EXACT)
    tr.close()
    assert_bit_equal(got["out"], z[f"{name}.out"], "predict")
    assert abs(got["loss"] - float(z[f"{name}.loss"])) < 1e-6          # the reference's row sum is one fp32 chain, ours is in double
    assert_bit_equal(got["grad_x"], z[f"{name}.grad_x"], "input gradient")
    for i, (k, _, _) in enumerate(layers):
        if k == capi.LINEAR:
            assert_bit_equal(got[f"gW{i}"], z[f"{name}.gW{i}"], f"grad_W of layer {i}")
            assert_bit_equal(got[f"gb{i}"], z[f"{name}.gb{i}"].reshape(-1), f"grad_bias of layer {i}")
            # SGD_step: the reference binary evaluates its a * b + c expressions fused (GCC, -O3, FMA hardware)
            assert_bit_equal(got[f"vW{i}"], z[f"{name}.vW{i}"], f"velocity of layer {i}")
            assert_bit_equal(got[f"W1_{i}"], z[f"{name}.W1_{i}"], f"weights of layer {i} after the step")
            assert_bit_equal(got[f"b1_{i}"], z[f"{name}.b1_{i}"].reshape(-1), f"bias of layer {i} after the step")
    assert_bit_equal(got["out1"], z[f"{name}.out1"], "predict after the step")


@pytest.mark.parametrize("name", CASES)
def test_trainer_fast_mode_within_tolerance(ctx, z, name):
    layers = golden_layers(z, name)
    row_ptr, col, W, NW = csr_of(z, name)
    ctx.graph_upload(row_ptr, col, W, NW)
    tr = capi.Trainer(ctx, layers)
    got = run_sequence(tr, z, name, pkg.MODE_FAST)
    tr.close()
    assert rel_to_scale(got["out"], z[f"{name}.out"]) < 1e-5
    assert abs(got["loss"] - float(z[f"{name}.loss"])) < 1e-5
    assert rel_to_scale(got["grad_x"], z[f"{name}.grad_x"]) < 1e-4
    for i, (k, _, _) in enumerate(layers):
        if k == capi.LINEAR:
            assert rel_to_scale(got[f"gW{i}"], z[f"{name}.gW{i}"]) < 1e-4, i
            assert rel_to_scale(got[f"gb{i}"], z[f"{name}.gb{i}"].reshape(-1)) < 1e-4, i
            assert rel_to_scale(got[f"W1_{i}"], z[f"{name}.W1_{i}"]) < 1e-5, i
    assert rel_to_scale(got["out1"], z[f"{name}.out1"]) < 1e-4


def test_trainer_on_a_larger_graph_fast_matches_exact_and_float64(ctx):
    """65 536 vertices / 1 M edges: fast against exact, and both against the float64 restatement"""
    g = graphs.rmat_graph(16, 16, seed=21)
    row_ptr, col, W, NW, x, s = inputs_of(g)
    layers = capi.random_model(5)
    ctx.graph_upload(row_ptr, col, W, NW)
    rng = np.random.default_rng(3)
    y = (rng.random((g.n, 1)) < 0.5).astype(np.float32)
    res = {}
    for mode in (pkg.MODE_EXACT, pkg.MODE_FAST):
        tr = capi.Trainer(ctx, layers)
        out = tr.predict(x, s, mode)
        loss = tr.mse_backprop(y, mode)
        res[mode] = (out, loss, {i: tr.read(1, i) for i, (k, _, _) in enumerate(layers) if k == capi.LINEAR})
        tr.close()
    out64, _, _ = po.train_backward_numpy(layers, [s], row_ptr, col, W, NW, x, np.zeros_like(y))
    _, _, g64 = po.train_backward_numpy(layers, [s], row_ptr, col, W, NW, x, 2.0 * (out64 - y))
    assert rel_to_scale(res[pkg.MODE_EXACT][0], out64) < 1e-5 and rel_to_scale(res[pkg.MODE_FAST][0], out64) < 1e-5
    assert abs(res[pkg.MODE_EXACT][1] - res[pkg.MODE_FAST][1]) < 1e-6
    for i in g64:
        for part in (0, 1):
            assert rel_to_scale(res[pkg.MODE_FAST][2][i][part], g64[i][part]) < 2e-4, (i, part)
            assert rel_to_scale(res[pkg.MODE_EXACT][2][i][part], g64[i][part]) < 2e-3, (i, part)   # 65 536-term fp32 chains


def test_backprop_needs_a_predict_on_the_current_graph(ctx, z):
    name = "grid12"
    layers = golden_layers(z, name)
    row_ptr, col, W, NW = csr_of(z, name)
    ctx.graph_upload(row_ptr, col, W, NW)
    tr = capi.Trainer(ctx, layers)
    with pytest.raises(capi.GvcError, match="without a predict"):
        tr.backprop(np.zeros((len(W), 1), np.float32))
    tr.predict(z[f"{name}.x"], 200.0)
    row_ptr2, col2, W2, NW2 = csr_of(z, "er300")
    ctx.graph_upload(row_ptr2, col2, W2, NW2)                   # another graph: the saved activations are stale
    with pytest.raises(capi.GvcError, match="without a predict"):
        tr.backprop(np.zeros((len(W2), 1), np.float32))
    with pytest.raises(capi.GvcError, match="35 x 32"):
        capi.Trainer(ctx, [(capi.LINEAR, np.zeros((40, 8), np.float32), np.zeros(8, np.float32))])
    tr.close()


def test_trainer_on_an_empty_graph_and_a_generic_model(ctx):
    """n = 0 is a no-op everywhere (the solver calls predict on an empty graph); a model that is not
    the GNN_VC architecture (graph layer of width 3, one linear layer, sigmoid) trains through the same kernels"""
    layers = [(capi.GRAPH, None, None), (capi.LINEAR, np.full((9, 2), 0.1, np.float32), np.zeros(2, np.float32)), (capi.SIGMOID, None, None)]
    ctx.graph_upload(np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    tr = capi.Trainer(ctx, layers)
    assert (tr.in_w, tr.out_w) == (3, 2)
    assert tr.predict(np.zeros((0, 3), np.float32), 200.0).shape == (0, 2)
    assert tr.backprop(np.zeros((0, 2), np.float32)).shape == (0, 3)
    assert tr.mse_backprop(np.zeros((0, 2), np.float32)) == 0.0
    g = graphs.er_graph(500, 1500, seed=3)
    row_ptr, col, W, NW, _, s = inputs_of(g)
    ctx.graph_upload(row_ptr, col, W, NW)
    rng = np.random.default_rng(8)
    x = rng.random((g.n, 3)).astype(np.float32)
    y = rng.random((g.n, 2)).astype(np.float32)
    out = tr.predict(x, s)
    loss = tr.mse_backprop(y)
    out64, _, _ = po.train_backward_numpy(layers, [s], row_ptr, col, W, NW, x, np.zeros_like(y))
    _, gx64, g64 = po.train_backward_numpy(layers, [s], row_ptr, col, W, NW, x, (out64 - y))     # MSE_grad: 2 (x - y) / width, width 2
    assert rel_to_scale(out, out64) < 1e-5
    assert abs(loss - float(np.mean(np.sum((out64 - y) ** 2, axis=1) / 2))) < 1e-5
    gW, gb = tr.read(1, 1)
    assert rel_to_scale(gW, g64[1][0]) < 1e-4 and rel_to_scale(gb, g64[1][1]) < 1e-4
    tr.predict(x, s)
    gx = tr.backprop((out64 - y).astype(np.float32))
    assert rel_to_scale(gx, gx64) < 1e-4
    tr.close()


def test_single_layer_entry_points_vs_the_live_reference(ctx, z):
    if not po.TRAIN_REF_SO.exists():
        pytest.skip("oracle/_ref/libgnntrainref.so did not travel")
    ref = po.TrainHarness(threads=1)
    rng = np.random.default_rng(17)
    lib, h = ctx.lib, ctx.h
    f32p = capi._f32p
    for n, K, N in ((257, 35, 32), (1000, 5, 32), (64, 16, 1), (3, 32, 16)):
        Wm = rng.standard_normal((K, N)).astype(np.float32) * 0.3
        b = rng.standard_normal(N).astype(np.float32)
        x = rng.standard_normal((n, K)).astype(np.float32)
        g = rng.standard_normal((n, N)).astype(np.float32)
        _, gW, gb, gi = ref.linear_layer(Wm, b, x, g)
        gW2, gb2, gi2 = np.zeros((K, N), np.float32), np.zeros(N, np.float32), np.empty((n, K), np.float32)
        ctx._check(lib.gvc_linear_backward_host(h, n, K, N, capi._ptr(x, f32p), capi._ptr(g, f32p), capi._ptr(Wm, f32p), capi._ptr(gW2, f32p),
                                                capi._ptr(gb2, f32p), capi._ptr(gi2, f32p), pkg.MODE_EXACT))
        assert_bit_equal(gW2, gW, f"grad_W {n}x{K}x{N}")
        assert_bit_equal(gb2, gb, f"grad_bias {n}x{K}x{N}")
        assert_bit_equal(gi2, gi, f"grad_in {n}x{K}x{N}")
        gW3, gb3, gi3 = np.zeros((K, N), np.float32), np.zeros(N, np.float32), np.empty((n, K), np.float32)
        ctx._check(lib.gvc_linear_backward_host(h, n, K, N, capi._ptr(x, f32p), capi._ptr(g, f32p), capi._ptr(Wm, f32p), capi._ptr(gW3, f32p),
                                                capi._ptr(gb3, f32p), capi._ptr(gi3, f32p), pkg.MODE_FAST))
        assert rel_to_scale(gW3, gW) < 1e-5 and rel_to_scale(gb3, gb) < 1e-5 and rel_to_scale(gi3, gi) < 1e-5
    # graph layer backward on a golden graph, widths 1 and 16
    name = "rmat9"
    row_ptr, col, W, NW = csr_of(z, name)
    ctx.graph_upload(row_ptr, col, W, NW)
    hr = ref.create([(po.GRAPH, None, None)])
    ref.set_graph(hr, len(W), z[f"{name}.eu"], z[f"{name}.ev"], W)
    for w in (1, 16):
        x = rng.standard_normal((len(W), w)).astype(np.float32)
        g = rng.standard_normal((len(W), 2 * w + 3)).astype(np.float32)
        _, gi = ref.graph_layer(hr, x, g, 200.0)
        gi2 = np.empty((len(W), w), np.float32)
        ctx._check(lib.gvc_graph_backward_host(h, capi._ptr(g, f32p), w, capi._ptr(gi2, f32p)))
        assert_bit_equal(gi2, gi, f"graph backward w={w}")
    ref.destroy(hr)
    # activations and the loss
    zz = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 4, np.float32([0.0, -0.0, 30.0, -30.0, np.inf, -np.inf])])
    g = rng.standard_normal(zz.size).astype(np.float32)
    for kind, fn in ((po.RELU, lambda o: lib.gvc_relu_backward_host(h, zz.size, capi._ptr(zz, f32p), capi._ptr(g, f32p), capi._ptr(o, f32p))),
                     (po.SIGMOID, lambda o: lib.gvc_sigmoid_backward_host(h, zz.size, capi._ptr(zz, f32p), capi._ptr(g, f32p), capi._ptr(o, f32p), pkg.MODE_EXACT))):
        _, gi = ref.activation(kind, zz, g)
        o = np.empty_like(zz)
        ctx._check(fn(o))
        assert_bit_equal(o, gi, f"activation {kind} backward")
    xx, yy = rng.random((777, 3)).astype(np.float32), rng.random((777, 3)).astype(np.float32)
    loss, grad = ref.mse(xx, yy)
    l2, g2 = np.zeros(1, np.float32), np.empty_like(xx)
    ctx._check(lib.gvc_mse_host(h, 777, 3, capi._ptr(xx, f32p), capi._ptr(yy, f32p), capi._ptr(l2, f32p), capi._ptr(g2, f32p)))
    assert_bit_equal(g2, grad, "MSE_grad")
    assert abs(float(l2[0]) - loss) < 1e-6


CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, "{root}")
sys.path.insert(0, "{root}/tests")
import gnn_mwvc_b200  # noqa: F401
from helpers import assert_bit_equal
from oracle import pyoracle as po
from test_train_cpu import CASES, golden_layers

z = np.load("{root}/tests/golden/train_vectors.npz")
d = po.TrainHarness(dropin=True)
for name in CASES:
    layers = golden_layers(z, name)
    scale = float(z[name + ".scale"])
    h = d.create(layers, scales=np.full(len(layers), scale, np.float32))
    d.set_graph(h, len(z[name + ".w"]), z[name + ".eu"], z[name + ".ev"], z[name + ".w"])
    x, y = z[name + ".x"], z[name + ".y"]
    assert_bit_equal(d.predict(h, x), z[name + ".out"], "predict")
    loss = d.mse_step(h, y)
    assert abs(loss - float(z[name + ".loss"])) < 1e-6, (loss, float(z[name + ".loss"]))
    d.predict(h, x)
    d.mse_step(h, y)
    d.predict(h, x)
    assert_bit_equal(d.backprop(h, z[name + ".g_dir"]), z[name + ".grad_x"], "grad_x")
    for i, (k, W, _) in enumerate(layers):
        if k == po.LINEAR:
            gW, gb = d.read(h, 1, i, W.shape)
            assert_bit_equal(gW, z[name + f".gW{{i}}"], f"grad_W {{i}}")
            assert_bit_equal(gb, z[name + f".gb{{i}}"].reshape(-1), f"grad_bias {{i}}")
    d.sgd_step(h, 3 * len(x), lr=0.05, momentum=0.9, wd=0.001)
    d.zero_grad(h)
    for i, (k, W, _) in enumerate(layers):
        if k == po.LINEAR:
            W1, b1 = d.read(h, 0, i, W.shape)
            assert_bit_equal(W1, z[name + f".W1_{{i}}"], f"W after SGD {{i}}")
            assert_bit_equal(b1, z[name + f".b1_{{i}}"].reshape(-1), f"bias after SGD {{i}}")
    assert_bit_equal(d.predict(h, x), z[name + ".out1"], "predict after the step")
    # the layer structs' own forward / backward (host buffers per call)
    rng = np.random.default_rng(4)
    Wm, b = layers[1][1], layers[1][2]
    xx = rng.standard_normal((len(x), 5)).astype(np.float32)
    gg = rng.standard_normal((len(x), 32)).astype(np.float32)
    out, gW, gb, gi = d.linear_layer(Wm, b, xx, gg)
    if po.TRAIN_REF_SO.exists():
        r = po.TrainHarness(threads=1)
        for a, bb, what in zip((out, gW, gb, gi), r.linear_layer(Wm, b, xx, gg), ("out", "grad_W", "grad_bias", "grad_in")):
            assert_bit_equal(a, bb, "linear layer struct: " + what)
    text = d.text(h)
    h2 = d.parse(text)
    assert d.text(h2) == text
    d.destroy(h)
    d.destroy(h2)
    print("training drop-in", name, "ok", flush=True)
'''


def test_dropin_through_the_reference_training_interface():
    if not po.TRAIN_DROPIN_SO.exists():
        pytest.skip("oracle/_ref/libgnntraindropin.so not built (needs /root/reference's headers at build time)")
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == len(CASES)
