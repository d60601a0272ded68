"""CPU: the drop-in host units (gnn-mwvc_b200/host/*.cpp) behind the reference's own C++ interface,
against the unmodified reference, for everything that does not compute: the text format both ways
(operator>>, operator<<, matrix I/O) and the constructors' random
initialisation.  Both sides are reached through the same C harness (oracle/ref_harness.cpp), built
once over the reference's src/*.cpp and once over the replacement units."""
import numpy as np
import pytest

from oracle import pyoracle as po


@pytest.fixture(scope="module")
def both():
    if not po.REF_SO.exists():
        pytest.skip("compiled reference not available")
    so = po.build_dropin_harness()
    if so is None or not so.exists():
        pytest.skip("drop-in harness not built (needs /root/reference headers)")
    return po.Reference(threads=1), po.Reference(so=so)


def test_trained_model_text_round_trip(both):
    ref, ours = both
    text = po.reference_model_text()
    a, b = ref.model(text), ours.model(text)
    ta, tb = ref.model_text(a), ours.model_text(b)
    assert ta == tb and len(ta) > 50_000                 # byte for byte what the reference prints
    c = ours.model(ta)                                    # and our parser reads the reference's print-out
    assert ours.model_text(c) == ta
    for h, side in ((a, ref), (b, ours), (c, ours)):
        side.destroy(h)


def test_parser_follows_the_reference_on_odd_input(both):
    ref, ours = both
    rng = np.random.default_rng(5)

    def mat(r, c):
        rows = [" ".join(repr(float(np.float32(v))) for v in rng.standard_normal(c)) for _ in range(r)]
        return f"{r} {c}\\n" + "\\n".join(rows) + "\\n"
    texts = [
        "tiny\\n3 Layers\\nGraph_Layer\\n\\nLinear_Layer\\nWeights: " + mat(5, 1) + "Bias: " + mat(1, 1) + "\\nSigmoid_Activation\\n\\n",
        # an unknown layer name is skipped by the reference's parser (it still counts as one of the n)
        "odd\\n4 Layers\\nReLU_Activation\\n\\nDropout_Layer\\n\\nGraph_Layer\\n\\nSigmoid_Activation\\n\\n",
        "empty\\n0 Layers\\n",
        "sci\\n1 Layers\\nLinear_Layer\\nWeights: 2 2\\n1e-3 -2.5E+2\\n3 4\\nBias: 1 2\\n0.125 -0\\n\\n",
    ]
    for t in texts:
        a, b = ref.model(t), ours.model(t)
        assert ours.model_text(b) == ref.model_text(a), t[:20]
        ref.destroy(a); ours.destroy(b)


@pytest.mark.parametrize("dims", [(5, 32, 0), (35, 32, 7), (16, 1, 123456789), (1, 1, 1)])
def test_linear_layer_random_init(both, dims):
    ref, ours = both
    K, N, seed = dims
    a = np.empty(K * N + N, np.float32)
    b = np.empty_like(a)
    ref.L.ref_linear_init(K, N, seed, po._p(a, po._f32p))
    ours.L.ref_linear_init(K, N, seed, po._p(b, po._f32p))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))     # same mt19937 stream, weights then bias
    assert np.abs(a).max() <= 1.0 / np.sqrt(K + 1) + 1e-7


def test_matrix_container_behaves_like_the_reference(both):
    """class matrix through its public interface: construction, element access,
    the stateful row selection of operator[] / raw(), iterator pairs, resize, text form both ways."""
    import ctypes as C
    ref, ours = both
    got = []
    for side in (ref, ours):
        side.L.ref_matrix_probe.restype = C.c_size_t
        side.L.ref_matrix_probe.argtypes = [po._f32p, C.c_size_t, C.c_char_p, C.c_size_t]
        out = np.full(4096, np.nan, np.float32)
        text = C.create_string_buffer(1 << 16)
        n = side.L.ref_matrix_probe(po._p(out, po._f32p), out.size, text, len(text))
        assert 50 < n <= out.size
        got.append((out[:n].copy(), text.value.decode()))
    (a, ta), (b, tb) = got
    assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert ta == tb and ta.count("|") == 4


def test_model_built_in_code_prints_like_the_reference(both):
    """model(name), add_layer, the layer constructors, copying a model, the empty model
     -- printed through operator<< on both sides."""
    import ctypes as C
    ref, ours = both
    texts = []
    for side in (ref, ours):
        side.L.ref_model_build_probe.restype = C.c_size_t
        side.L.ref_model_build_probe.argtypes = [C.c_char_p, C.c_size_t]
        buf = C.create_string_buffer(1 << 16)
        n = side.L.ref_model_build_probe(buf, len(buf))
        assert 100 < n < len(buf)
        texts.append(buf.value.decode())
    assert texts[0] == texts[1] and texts[0].count("|") == 3 and "built_in_code" in texts[0]
