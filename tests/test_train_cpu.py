"""CPU: the checker of the TRAINING path.  The reference's own
old_files/src/lib/gnn_training.cpp, compiled unmodified behind oracle/train_harness.cpp, (a) reproduces the
committed golden vectors bit for bit, (b) agrees with the numpy restatement of forward + backward
(oracle/pyoracle.py train_backward_numpy, float64) within fp32 accuracy, and (c) round-trips its text format.
No GPU involved."""
import numpy as np
import pytest

import gnn_mwvc_b200  # noqa: F401
from conftest import GOLDEN
from helpers import assert_bit_equal
from oracle import pyoracle as po

PATTERN = [po.GRAPH, po.LINEAR, po.RELU, po.LINEAR, po.RELU, po.LINEAR, po.RELU,
           po.GRAPH, po.LINEAR, po.RELU, po.LINEAR, po.RELU, po.LINEAR, po.RELU,
           po.GRAPH, po.LINEAR, po.RELU, po.LINEAR, po.RELU, po.LINEAR, po.SIGMOID]
CASES = ("er300", "rmat9", "grid12")


def golden_layers(z, name, prefix=("W", "b")):
    return [(k, z[f"{name}.{prefix[0]}{i}"], z[f"{name}.{prefix[1]}{i}"]) if k == po.LINEAR else (k, None, None)
            for i, k in enumerate(PATTERN)]


def csr_of(z, name):
    from gnn_mwvc_b200 import graphs
    import torch
    w = z[f"{name}.w"]
    g = graphs.graph_from_edges(len(w), torch.from_numpy(z[f"{name}.eu"].astype(np.int64)), torch.from_numpy(z[f"{name}.ev"].astype(np.int64)),
                                torch.from_numpy(w.astype(np.int64)), name=name)
    return g.numpy()


@pytest.fixture(scope="module")
def ref():
    if not po.TRAIN_REF_SO.exists():
        pytest.skip("oracle/_ref/libgnntrainref.so not built (needs /root/reference at build time)")
    return po.TrainHarness(threads=1)


@pytest.fixture(scope="module")
def z():
    return np.load(GOLDEN / "train_vectors.npz")


@pytest.mark.parametrize("name", CASES)
def test_reference_reproduces_the_golden_training_vectors(ref, z, name):
    layers = golden_layers(z, name)
    scale = float(z[f"{name}.scale"])
    h = ref.create(layers, scales=np.full(len(layers), scale, np.float32))
    ref.set_graph(h, len(z[f"{name}.w"]), z[f"{name}.eu"], z[f"{name}.ev"], z[f"{name}.w"])
    x, y = z[f"{name}.x"], z[f"{name}.y"]
    assert_bit_equal(ref.predict(h, x), z[f"{name}.out"], "predict")
    loss = ref.mse_step(h, y)
    assert np.float32(loss) == z[f"{name}.loss"]
    ref.predict(h, x)
    ref.mse_step(h, y)
    ref.predict(h, x)
    assert_bit_equal(ref.backprop(h, z[f"{name}.g_dir"]), z[f"{name}.grad_x"], "grad_x")
    for i, (k, W, _) in enumerate(layers):
        if k == po.LINEAR:
            gW, gb = ref.read(h, 1, i, W.shape)
            assert_bit_equal(gW, z[f"{name}.gW{i}"], f"grad_W {i}")
            assert_bit_equal(gb, z[f"{name}.gb{i}"], f"grad_bias {i}")
    ref.sgd_step(h, 3 * len(x), lr=0.05, momentum=0.9, wd=0.001)
    ref.zero_grad(h)
    for i, (k, W, _) in enumerate(layers):
        if k == po.LINEAR:
            W1, b1 = ref.read(h, 0, i, W.shape)
            assert_bit_equal(W1, z[f"{name}.W1_{i}"], f"W after SGD {i}")
            assert_bit_equal(b1, z[f"{name}.b1_{i}"], f"bias after SGD {i}")
            gW, gb = ref.read(h, 1, i, W.shape)
            assert not gW.any() and not gb.any()
    assert_bit_equal(ref.predict(h, x), z[f"{name}.out1"], "predict after the step")
    ref.destroy(h)


def rel_to_scale(got, want):
    """max |got - want| relative to the largest |want| (gradients cancel: elementwise relative error means nothing)"""
    want = np.asarray(want, np.float64)
    return float(np.max(np.abs(np.asarray(got, np.float64) - want)) / max(np.max(np.abs(want)), 1e-30))


@pytest.mark.parametrize("name", CASES)
def test_numpy_restatement_agrees_with_the_golden_vectors(z, name):
    """the float64 restatement of predict + backprop against what the reference computed in fp32"""
    layers = golden_layers(z, name)
    row_ptr, col, W, NW = csr_of(z, name)
    scale = float(z[f"{name}.scale"])
    x, y = z[f"{name}.x"], z[f"{name}.y"]
    out, _, _ = po.train_backward_numpy(layers, [scale], row_ptr, col, W, NW, x, np.zeros_like(y))
    assert rel_to_scale(out, z[f"{name}.out"]) < 1e-5
    # the recorded gradients: two MSE passes + one pass with g_dir, accumulated
    g_mse = 2.0 * (out - y.astype(np.float64))                    # MSE_grad :184-190, width 1
    _, _, grads_mse = po.train_backward_numpy(layers, [scale], row_ptr, col, W, NW, x, g_mse)
    _, gx, grads_dir = po.train_backward_numpy(layers, [scale], row_ptr, col, W, NW, x, z[f"{name}.g_dir"])
    assert rel_to_scale(gx, z[f"{name}.grad_x"]) < 1e-4
    for i, (k, _, _) in enumerate(layers):
        if k == po.LINEAR:
            assert rel_to_scale(2 * grads_mse[i][0] + grads_dir[i][0], z[f"{name}.gW{i}"]) < 1e-4, i
            assert rel_to_scale(2 * grads_mse[i][1] + grads_dir[i][1], z[f"{name}.gb{i}"]) < 1e-4, i
    loss = float(np.mean((out - y) ** 2))
    assert abs(loss - float(z[f"{name}.loss"])) < 1e-6
    # SGD_step :192-224 on the recorded gradients
    for i, (k, Wm, b) in enumerate(layers):
        if k == po.LINEAR:
            for p0, g0, p1, v1 in ((Wm, z[f"{name}.gW{i}"], z[f"{name}.W1_{i}"], z[f"{name}.vW{i}"]),
                                   (b, z[f"{name}.gb{i}"], z[f"{name}.b1_{i}"], z[f"{name}.vb{i}"])):
                gg = g0.astype(np.float64) + 2 * 0.001 * p0.astype(np.float64)
                vel = gg / (3 * len(x))
                assert rel_to_scale(vel, v1.reshape(vel.shape)) < 1e-5
                assert rel_to_scale(p0 - 0.05 * vel, p1.reshape(vel.shape)) < 1e-6


def test_training_model_text_round_trip(ref, z):
    """operator<< / operator>> of model_training (:131-173): the same records as the inference model"""
    layers = golden_layers(z, "er300")
    h = ref.create(layers, scales=np.full(len(layers), 200.0, np.float32))
    text = ref.text(h)
    assert text.split()[1:3] == ["21", "Layers"] and text.count("Linear_Layer") == 9 and text.count("Graph_Layer") == 3
    h2 = ref.parse(text)
    assert ref.text(h2) == text
    ref.destroy(h)
    ref.destroy(h2)
