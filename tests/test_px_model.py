"""CPU: the algorithm behind the exact-mode hub sums (gvc_px.cuh), restated in numpy (tests/px_model.py),
against the element-wise sequential fp32 chain it must reproduce -- friendly and adversarial inputs."""
import numpy as np
import pytest

from px_model import px_sum, seq_sum

f32 = np.float32


def _cases(rng, n):
    yield "uniform", rng.random(n).astype(f32)
    yield "x=W/200", (rng.integers(1, 201, n).astype(f32) / f32(200))
    yield "relu", np.maximum(rng.standard_normal(n), 0).astype(f32)
    yield "wide", np.exp(rng.uniform(-20, 5, n)).astype(f32)
    yield "few-bits (ties everywhere)", (rng.integers(0, 8, n) * 0.125).astype(f32)
    yield "halves", np.full(n, 0.5, f32)
    yield "tiny then big", np.concatenate([np.full(n // 2 + 1, 1e-30, f32), rng.random(n // 2 + 1).astype(f32)])
    yield "outliers", np.where(rng.random(n) < 0.01, 1e6, rng.random(n)).astype(f32)
    yield "negatives", rng.standard_normal(n).astype(f32)
    yield "NaN inside", np.where(np.arange(n) == n // 2, np.nan, rng.random(n)).astype(f32)
    yield "inf inside", np.where(np.arange(n) == n // 3, np.inf, rng.random(n)).astype(f32)
    yield "subnormal", (rng.random(n) * 1e-39).astype(f32)
    yield "overflowing", (rng.random(n) * 1e37).astype(f32)
    yield "all zero", np.zeros(n, f32)
    yield "power-of-two crossings", np.full(n, 1.0, f32)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 20001])
@pytest.mark.parametrize("batch", [64, 256])
def test_parallel_sum_equals_the_sequential_chain(n, batch):
    rng = np.random.default_rng(n + batch)
    with np.errstate(all="ignore"):
        for name, v in _cases(rng, n):
            st = []
            a, b = seq_sum(v), px_sum(v, BATCH=batch, stats=st)
            assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), (name, n, a, b)


def test_friendly_inputs_take_the_fast_way():
    rng = np.random.default_rng(1)
    st = []
    px_sum(rng.random(70001).astype(f32), stats=st)
    nb, slow = st[0]
    assert slow < nb // 20                      # a few dozen of ~1100 batches go element by element
