#!/usr/bin/env python
"""Round-2 additions to tests/golden/, again outputs of the UNMODIFIED reference (oracle/_ref, OpenBLAS
0.3.15 Prescott, ONE thread).  Run where /root/reference exists, after `make -C oracle ref`:
    python tools/make_golden_r2.py
  dot_vectors.npz      dot() (src/matrix.cpp:106-122) over every OpenBLAS block class (rows 4/2/1 x
                       columns 8/4/2/1), both transposes, beta, and k long enough to be cut into blocks
  mixed_scales.npz     predict of a model assembled with add_layer whose three graph layers carry
                       DIFFERENT WEIGHT_SCALEs (include/gnn_inference.hpp:25; src/gnn_inference.cpp:38-40)
  reduced_graphs.npz   graphs after real reduction_graph mutators (remove_node, remove_neighborhood,
                       fold_neighborhood, fold_isolated, relable_graph; include/reduction_graph.hpp:248-587):
                       the mutation script, the CSR predict then reads, and the reference's scores
  sort_order.npz       the order src/GNN_VC.cpp:194-206 puts the vertices in after predict (std::sort with
                       the eps-tolerance comparator on min(out, 1-out)), for two graphs
The existing files of round 1 are not touched.
"""
from __future__ import annotations

import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gnn_mwvc_b200  # noqa: E402,F401
from gnn_mwvc_b200 import capi, graphs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

GOLD = ROOT / "tests" / "golden"

DOT_CASES = [  # m, n, k, at, bt, beta
    (7, 15, 47, 0, 0, 0.0), (6, 12, 300, 0, 0, 0.0), (5, 3, 41, 1, 0, 0.0), (9, 1, 35, 0, 0, 0.0),
    (1, 7, 19, 0, 1, 0.0), (3, 2, 600, 1, 1, 0.5), (64, 35, 32, 0, 1, 0.0), (35, 32, 1001, 1, 0, 1.0),
    (13, 8, 9, 0, 0, -0.75), (2, 6, 130, 0, 0, 0.0), (11, 14, 16, 1, 1, 0.0), (1, 1, 40, 0, 0, 0.0),
]


def wide(rng, shape):
    """values over many binades and both signs: summation orders then differ in the last bits"""
    return (np.exp(rng.uniform(-6, 6, shape)) * rng.choice([-1, 1], shape)).astype(np.float32)


def mutation_script(ref, gh, seed, steps):
    """Apply random mutators the way reductions do; return the list of (op, u) that were accepted."""
    rnd = random.Random(seed)
    done = []
    for _ in range(steps):
        op = rnd.choice([ref.REMOVE_NODE, ref.REMOVE_NEIGHBORHOOD, ref.FOLD_NEIGHBORHOOD, ref.FOLD_NEIGHBORHOOD,
                         ref.FOLD_ISOLATED, ref.REMOVE_NODE])
        u = rnd.randrange(ref.graph_size(gh))
        if ref.graph_mutate(gh, op, u):
            done.append((op, u))
    return done


def main():
    ref = po.Reference(threads=1)
    orc = po.Oracle()
    layers = capi.load_model_npz(GOLD / "mwvc_model.npz")
    rng = np.random.default_rng(2)

    # ---- dot --------------------------------------------------------------------------------
    d = {}
    for i, (m, n, k, at, bt, beta) in enumerate(DOT_CASES):
        A, B, C0 = wide(rng, (m, k)), wide(rng, (k, n)), wide(rng, (m, n))
        As = A.T.copy() if at else A
        Bs = B.T.copy() if bt else B
        out = ref.dot(As, Bs, C0, bool(at), bool(bt), beta)
        assert np.array_equal(out.view(np.uint32), orc.dot(As, Bs, C0, bool(at), bool(bt), beta).view(np.uint32)), i
        d[f"c{i}.A"], d[f"c{i}.B"], d[f"c{i}.C0"], d[f"c{i}.out"] = As, Bs, C0, out
        d[f"c{i}.flags"] = np.array([at, bt], np.int32)
        d[f"c{i}.beta"] = np.float32(beta)
    np.savez_compressed(GOLD / "dot_vectors.npz", **d)
    print("dot:", len(DOT_CASES), "cases")

    # ---- different WEIGHT_SCALE per graph layer -------------------------------------------------
    g = graphs.er_graph(1501, 6000, seed=17)
    rp, col, W, NW = g.numpy()
    eu, ev = g.edges_numpy()
    scales3 = [20.0, 200.0, 57.0]
    per_layer, it = [], iter(scales3)
    for k, _, _ in layers:
        per_layer.append(next(it) if k == po.GRAPH else 0.0)
    hm = ref.model_build(layers, per_layer)
    gh = ref.graph_create(g.n, eu, ev, W)
    x = W.astype(np.float32) / np.float32(200.0)
    want = ref.predict_on_as_is(hm, gh, x)
    ho = orc.parse(po.layers_to_text(layers))
    for i, s in enumerate(scales3):
        orc.set_graph_layer_scale(ho, i, s)
    mine = orc.predict(ho, rp, col, W, NW, x)[:, 0]
    assert np.array_equal(want.view(np.uint32), mine.view(np.uint32))
    np.savez_compressed(GOLD / "mixed_scales.npz", eu=eu, ev=ev, w=W, x=x, scales=np.array(scales3, np.float32), scores=want)
    ref.graph_destroy(gh)
    print("mixed scales:", want[:3])

    # ---- graphs after real reductions ---------------------------------------------------------------
    hr = ref.model(po.layers_to_text(layers))
    ho = orc.parse(po.layers_to_text(layers))
    red = {}
    for name, g, seed, steps in (("er3000", graphs.er_graph(3000, 9000, seed=31), 5, 500),
                                 ("er800_dense", graphs.er_graph(800, 6000, seed=32), 6, 120),
                                 ("grid40", graphs.grid_graph(40, 41, seed=33), 7, 300)):
        rp, col, W, NW = g.numpy()
        eu, ev = g.edges_numpy()
        gh = ref.graph_create(g.n, eu, ev, W)
        script = []
        for rnd in range(3):                       # three rounds of reductions + relabel + predict, as gnn_solve does
            script += mutation_script(ref, gh, seed * 10 + rnd, steps)
            assert ref.graph_mutate(gh, ref.RELABEL)
            script.append((ref.RELABEL, 0))
            rp2, col2, w2, nw2, act = ref.graph_csr(gh)
            assert act.all()
            n2 = ref.graph_size(gh)
            x = w2.astype(np.float32) / np.float32(200.0)
            scores = ref.predict_on(hr, gh, x, 200.0)
            mine = orc.predict(ho, rp2, col2, w2, nw2, x, 200.0)[:, 0] if n2 else np.zeros(0, np.float32)
            assert np.array_equal(scores.view(np.uint32), mine.view(np.uint32)), (name, rnd)
            key = f"{name}.r{rnd}"
            red[key + ".row_ptr"], red[key + ".col"], red[key + ".w"], red[key + ".nw"] = rp2, col2, w2, nw2
            red[key + ".scores"] = scores
            red[key + ".script_len"] = np.int64(len(script))
            print(key, "n", n2, "nnz", len(col2))
        red[f"{name}.eu"], red[f"{name}.ev"], red[f"{name}.w0"] = eu, ev, W
        red[f"{name}.script"] = np.array(script, np.int64).reshape(-1, 2)
        ref.graph_destroy(gh)
    np.savez_compressed(GOLD / "reduced_graphs.npz", **red)


if __name__ == "__main__":
    main()
