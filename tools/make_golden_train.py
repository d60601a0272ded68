#!/usr/bin/env python
"""Golden vectors of the TRAINING path, written by the reference itself:
old_files/src/lib/gnn_training.cpp, unmodified, compiled into oracle/_ref/libgnntrainref.so behind
oracle/train_harness.cpp (OpenBLAS 0.3.15 kernel set "Prescott", one thread -- the pin of every checker
here).  Run in the build container (needs /root/reference):
    python tools/make_golden_train.py        ->  tests/golden/train_vectors.npz
For every case: the model (GNN_VC architecture, the reference's seeded random init), a graph, x, labels y;
then what the reference computes: predict, MSE loss, the gradients after backprop (two accumulated passes),
parameters and velocities after SGD_step with weight decay, and predict again with the new parameters."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200  # noqa: E402,F401
from gnn_mwvc_b200 import graphs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

PATTERN = [po.GRAPH, po.LINEAR, po.RELU, po.LINEAR, po.RELU, po.LINEAR, po.RELU,
           po.GRAPH, po.LINEAR, po.RELU, po.LINEAR, po.RELU, po.LINEAR, po.RELU,
           po.GRAPH, po.LINEAR, po.RELU, po.LINEAR, po.RELU, po.LINEAR, po.SIGMOID]
DIMS = [(5, 32), (32, 32), (32, 16), (35, 32), (32, 32), (32, 16), (35, 32), (32, 16), (16, 1)]


def architecture():
    it = iter(DIMS)
    return [(k, next(it), None) if k == po.LINEAR else (k, None, None) for k in PATTERN]


def main():
    po.build_train_harness()
    ref = po.TrainHarness(threads=1)
    out = {}
    cases = {"er300": graphs.er_graph(300, 900, seed=5), "rmat9": graphs.rmat_graph(9, 6, seed=6), "grid12": graphs.grid_graph(12, 13, seed=7)}
    for case_no, (name, g) in enumerate(cases.items()):
        eu, ev = g.edges_numpy()
        W = g.numpy()[2]
        scale = 200.0
        layers = architecture()
        seeds = [100 + i for i in range(len(layers))]
        h = ref.create(layers, scales=np.full(len(layers), scale, np.float32), seeds=seeds)
        ref.set_graph(h, g.n, eu, ev, W)
        rng = np.random.default_rng(1000 + case_no)
        x = (W.astype(np.float32) / np.float32(scale)).reshape(-1, 1)
        y = (rng.random((g.n, 1)) < 0.4).astype(np.float32)
        out[f"{name}.eu"], out[f"{name}.ev"], out[f"{name}.w"] = eu.astype(np.uint32), ev.astype(np.uint32), W.astype(np.uint32)
        out[f"{name}.scale"], out[f"{name}.x"], out[f"{name}.y"] = np.float32(scale), x, y
        for i, (k, shape, _) in enumerate(layers):
            if k == po.LINEAR:
                out[f"{name}.W{i}"], out[f"{name}.b{i}"] = ref.read(h, 0, i, shape)
        out[f"{name}.out"] = ref.predict(h, x)
        out[f"{name}.loss"] = np.float32(ref.mse_step(h, y))            # MSE_loss + MSE_grad + backprop
        ref.predict(h, x)
        ref.mse_step(h, y)                                               # gradients accumulate over two passes
        g_dir = (rng.standard_normal((g.n, 1)) * 0.1).astype(np.float32)
        ref.predict(h, x)
        out[f"{name}.g_dir"] = g_dir
        out[f"{name}.grad_x"] = ref.backprop(h, g_dir)                   # a third pass with an arbitrary output gradient
        for i, (k, shape, _) in enumerate(layers):
            if k == po.LINEAR:
                out[f"{name}.gW{i}"], out[f"{name}.gb{i}"] = ref.read(h, 1, i, shape)
        ref.sgd_step(h, 3 * g.n, lr=0.05, momentum=0.9, wd=0.001)
        ref.zero_grad(h)
        for i, (k, shape, _) in enumerate(layers):
            if k == po.LINEAR:
                out[f"{name}.W1_{i}"], out[f"{name}.b1_{i}"] = ref.read(h, 0, i, shape)
                out[f"{name}.vW{i}"], out[f"{name}.vb{i}"] = ref.read(h, 2, i, shape)
        out[f"{name}.out1"] = ref.predict(h, x)
        ref.destroy(h)
        print(name, g.n, "vertices, loss", float(out[f"{name}.loss"]))
    np.savez_compressed(ROOT / "tests" / "golden" / "train_vectors.npz", **out)
    print("wrote tests/golden/train_vectors.npz")


if __name__ == "__main__":
    main()
