"""GPU: the drop-in THROUGH THE REFERENCE'S OWN C++ INTERFACE.  oracle/ref_harness.cpp is the C
harness that produced tests/golden/ from the unmodified reference; `make -C oracle dropin_harness`
builds the very same harness over gnn-mwvc_b200/host/*.cpp + libgvc instead.  So these calls are
gnn::model::predict(in, out, reduction_graph), graph_layer::forward, linear_layer::forward,
ReLU::forward and sigmoid::forward exactly as src/GNN_VC.cpp would make them -- a real
reduction_graph built by its constructor, walked through begin(u)/end(u)/W/NW by the drop-in -- and
the outputs must equal the reference's recorded ones bit for bit.

The calls run in a child process: the reference interface is `void`, so the drop-in reports a CUDA
problem by aborting, and that must fail one test, not take the test session down."""
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, "{root}")
sys.path.insert(0, "{root}/tests")
import gnn_mwvc_b200  # noqa: F401
from gnn_mwvc_b200 import capi
from helpers import assert_bit_equal, golden_names
from oracle import pyoracle as po

GOLDEN = "{root}/tests/golden/"
dropin = po.Reference(so=po.DROPIN_SO)
vec = np.load(GOLDEN + "predict_vectors.npz")
lay = np.load(GOLDEN + "layer_vectors.npz")
which = sys.argv[1]
if which == "predict":
    h = dropin.model(po.layers_to_text(capi.load_model_npz(GOLDEN + "mwvc_model.npz")))
    for name in golden_names(vec):
        w = vec[name + ".w"].astype(np.uint32)
        scale = float(vec[name + ".scale"])
        x = w.astype(np.float32) / np.float32(scale)
        got = dropin.predict(h, len(w), vec[name + ".eu"], vec[name + ".ev"], w, x, scale)
        assert_bit_equal(got, np.asarray(vec[name + ".scores"]).reshape(-1), name)
        print("predict", name, len(w), flush=True)
    dropin.destroy(h)
elif which == "reduced":
    # graphs as the solver hands them to predict: the committed mutation scripts replayed on a real
    # reduction_graph through its own mutators (holes, rotated lists, fold vertices appended at the end of
    # the edge array, relabelled ids), predict through the drop-in -- i.e. through the streamed upload
    # that reads begin(u)/end(u) -- against the reference's recorded scores
    z = np.load(GOLDEN + "reduced_graphs.npz")
    h = dropin.model(po.layers_to_text(capi.load_model_npz(GOLDEN + "mwvc_model.npz")))
    for name in ("er3000", "er800_dense", "grid40"):
        gh = dropin.graph_create(len(z[name + ".w0"]), z[name + ".eu"], z[name + ".ev"], z[name + ".w0"])
        script, done = z[name + ".script"], 0
        for rnd in range(3):
            upto = int(z["%s.r%d.script_len" % (name, rnd)])
            for op, u in script[done:upto]:
                assert dropin.graph_mutate(gh, int(op), int(u)), (name, rnd, op, u)
            done = upto
            w = z["%s.r%d.w" % (name, rnd)]
            assert dropin.graph_size(gh) == len(w)
            x = w.astype(np.float32) / np.float32(200.0)
            got = dropin.predict_on(h, gh, x, 200.0)
            assert_bit_equal(got, z["%s.r%d.scores" % (name, rnd)], "%s round %d" % (name, rnd))
            print("reduced", name, rnd, len(w), flush=True)
        dropin.graph_destroy(gh)
    # and a model whose graph layers carry different WEIGHT_SCALEs, assembled with add_layer
    m = np.load(GOLDEN + "mixed_scales.npz")
    layers = capi.load_model_npz(GOLDEN + "mwvc_model.npz")
    it = iter(m["scales"])
    hm = dropin.model_build(layers, [float(next(it)) if k == po.GRAPH else 0.0 for k, _, _ in layers])
    gh = dropin.graph_create(len(m["w"]), m["eu"], m["ev"], m["w"])
    assert_bit_equal(dropin.predict_on_as_is(hm, gh, m["x"]), m["scores"], "mixed scales through add_layer")
    dropin.graph_destroy(gh)
else:
    w = vec["er607.w"].astype(np.uint32)
    for width in (1, 16, 3):
        got = dropin.graph_layer(len(w), vec["er607.eu"], vec["er607.ev"], w, 200.0, lay["graph_er607_w%d.in" % width])
        assert_bit_equal(got, lay["graph_er607_w%d.out" % width], "graph layer w=%d" % width)
    for k in sorted(k[:-3] for k in lay.files if k.startswith("linear_") and k.endswith(".in")):
        assert_bit_equal(dropin.linear_layer(lay[k + ".in"], lay[k + ".W"], lay[k + ".b"]), lay[k + ".out"], k)
    assert_bit_equal(dropin.relu(lay["relu.in"]), np.asarray(lay["relu.out"]).reshape(-1), "relu")
    assert_bit_equal(dropin.sigmoid(lay["sigmoid.in"]), np.asarray(lay["sigmoid.out"]).reshape(-1), "sigmoid")
print("REFERENCE_INTERFACE_OK", flush=True)
'''


def run_child(which):
    if not po.DROPIN_SO.exists():
        pytest.skip("oracle/_ref/libgnndropin.so not built (needs the reference's headers at build time)")
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT), which], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "REFERENCE_INTERFACE_OK" in r.stdout, (r.stdout[-1500:] + "\n" + r.stderr[-3000:])


def test_predict_through_the_reference_interface():
    run_child("predict")


def test_single_layer_forwards_through_the_reference_interface():
    run_child("layers")


def test_predict_on_reduced_graphs_and_mixed_scales_through_the_reference_interface():
    run_child("reduced")
