#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference built in oracle/_ref.

Run in the container that has /root/reference (after `make -C oracle ref`):
    python tools/make_golden.py

The reference has no golden vectors of its own for the GNN forward (SURVEY.md
section 4), so the pins are outputs of the reference itself:
  mwvc_model.npz        the trained model of reference src/GNN_VC.cpp:23 as parsed
                        fp32 arrays (binary, so no reference text is committed)
  predict_vectors.npz   graphs (edge lists + weights) with the reference's scores,
                        OpenBLAS pinned to the Prescott kernel, ONE thread
  layer_vectors.npz     single-layer inputs/outputs (graph layer incl. its column
                        quirk, every linear shape, ReLU, sigmoid)
  er10k_run.json        cost / md5 of the result file of
                        `GNN_VC er10k.graph out 0 -1 0` (deterministic, time = 0)
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200  # noqa: E402,F401
from gnn_mwvc_b200 import graphs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    ref = po.Reference(threads=1)
    orc = po.Oracle()
    text = po.reference_model_text()

    # ---- model -----------------------------------------------------------------
    ho = orc.parse(text)
    layers = orc.layers(ho)
    hr = ref.model(text)
    # the reference's own parser, seen through its operator<< (6 significant digits,
    # which is what the source literal carries): re-parsing must give the same floats
    relayers = orc.layers(orc.parse(ref.model_text(hr)))
    assert len(relayers) == len(layers) == 21
    for (k, W, b), (k2, W2, b2) in zip(layers, relayers):
        assert k == k2
        if k == po.LINEAR:
            assert np.array_equal(W, W2) and np.array_equal(b, b2)
    arrs = {"kinds": np.array([k for k, _, _ in layers], np.int32)}
    for i, (k, W, b) in enumerate(layers):
        if k == po.LINEAR:
            arrs[f"W{i}"] = W
            arrs[f"b{i}"] = b
    np.savez(GOLD / "mwvc_model.npz", **arrs)
    print("model:", sum(W.size + b.size for k, W, b in layers if k == po.LINEAR), "parameters")

    # ---- whole-forward vectors -----------------------------------------------------
    cases = {
        "readme": graphs.graph_from_edges(3, torch.tensor([0, 1]), torch.tensor([2, 2]), torch.tensor([15, 15, 20])),
        "isolated": graphs.graph_from_edges(1, torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), torch.tensor([7])),
        "pair": graphs.graph_from_edges(2, torch.tensor([0]), torch.tensor([1]), torch.tensor([3, 200])),
        "er607": graphs.er_graph(607, 900, seed=2),
        "er1000": graphs.er_graph(1000, 5000, seed=3),
        "er3157": graphs.er_graph(3157, 6754, seed=4),
        "rmat10": graphs.rmat_graph(10, 8, seed=5),
        "rmat11_odd": graphs.rmat_graph(11, 16, seed=6, n_limit=2001),
        "grid9x11": graphs.grid_graph(9, 11, seed=7),
        "star": graphs.graph_from_edges(400, torch.zeros(399, dtype=torch.int64), torch.arange(1, 400), graphs.random_weights(400, 8)),
    }
    scales = {"readme": 20.0, "isolated": 20.0}
    vec = {}
    for name, g in cases.items():
        row_ptr, col, W, NW = g.numpy()
        eu, ev = g.edges_numpy()
        s = scales.get(name, float(W.max()))
        x = W.astype(np.float32) / np.float32(s)
        scores = ref.predict(hr, g.n, eu, ev, W, x, s)
        mine = orc.predict(ho, row_ptr, col, W, NW, x, s)[:, 0]
        assert np.array_equal(scores.view(np.uint32), mine.view(np.uint32)), name
        vec[f"{name}.eu"], vec[f"{name}.ev"] = eu, ev
        vec[f"{name}.w"] = W
        vec[f"{name}.scale"] = np.float32(s)
        vec[f"{name}.scores"] = scores
        print(f"{name}: n={g.n} E={g.n_edges} scores[:3]={scores[:3]}")
    np.savez_compressed(GOLD / "predict_vectors.npz", **vec)

    # ---- single-layer vectors --------------------------------------------------------
    rng = np.random.default_rng(0)
    lay = {}
    g = cases["readme"]
    eu, ev = g.edges_numpy()
    W = g.numpy()[2]
    xin = (100.0 * (np.arange(3)[:, None] + 1) + np.arange(16)[None, :]).astype(np.float32)
    lay["graph16.in"] = xin
    lay["graph16.out"] = ref.graph_layer(3, eu, ev, W, 20.0, xin)
    g = cases["er607"]
    eu, ev = g.edges_numpy()
    W = g.numpy()[2]
    for w in (1, 16, 3):
        xin = rng.standard_normal((g.n, w)).astype(np.float32)
        lay[f"graph_er607_w{w}.in"] = xin
        lay[f"graph_er607_w{w}.out"] = ref.graph_layer(g.n, eu, ev, W, 200.0, xin)
    for (n, K, N) in [(64, 5, 32), (64, 32, 32), (64, 32, 16), (64, 35, 32), (64, 16, 1),
                      (67, 32, 32), (67, 35, 32), (67, 32, 16), (67, 16, 1), (1, 32, 32), (2, 35, 32), (3, 16, 1)]:
        xin = rng.standard_normal((n, K)).astype(np.float32)
        Wm = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32)
        key = f"linear_{n}_{K}_{N}"
        lay[key + ".in"], lay[key + ".W"], lay[key + ".b"] = xin, Wm, b
        lay[key + ".out"] = ref.linear_layer(xin, Wm, b)
    xs = np.concatenate([rng.standard_normal(2000).astype(np.float32) * 8,
                         np.array([0.0, -0.0, 1e-30, -1e-30, 88.5, -88.5, 104.0, -104.0, 20.0, -20.0], np.float32)])
    lay["relu.in"], lay["relu.out"] = xs, ref.relu(xs)
    lay["sigmoid.in"], lay["sigmoid.out"] = xs, ref.sigmoid(xs)
    np.savez_compressed(GOLD / "layer_vectors.npz", **lay)

    # ---- deterministic end-to-end run (config 1) ---------------------------------------
    g = graphs.er10k_fixture()
    with tempfile.TemporaryDirectory() as td:
        gp = Path(td) / "er10k.graph"
        graphs.write_metis(g, gp)
        md5_graph = hashlib.md5(gp.read_bytes()).hexdigest()
        env = dict(os.environ, OPENBLAS_CORETYPE="Prescott", OPENBLAS_NUM_THREADS="1")
        out = subprocess.run([str(po.REF_BIN), str(gp), str(Path(td) / "out"), "0", "-1", "0"],
                             capture_output=True, text=True, env=env, check=True).stdout.strip()
        res = (Path(td) / "out").read_bytes()
    fields = out.split(",")
    info = {"graph_md5": md5_graph, "stdout_name": fields[0], "cost": int(fields[1]), "best_seen": int(fields[2]),
            "result_md5": hashlib.md5(res).hexdigest(), "cover_size": int(res.count(b"1")),
            "command": "GNN_VC er10k.graph out 0 -1 0", "openblas": ref.blas_config(), "threads": 1}
    (GOLD / "er10k_run.json").write_text(json.dumps(info, indent=1) + "\n")
    print(info)


if __name__ == "__main__":
    main()
