"""CPU: the reference arm of bench.py honours the output contract -- exactly one JSON line on stdout
(library chatter goes to stderr), the keys the driver reads, a bounded sample of the workload."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_json_line():
    if not (ROOT / "oracle" / "_ref" / "libgnnref.so").exists() and not (ROOT / "oracle" / "_build" / "libgnnoracle.so").exists():
        pytest.skip("neither the compiled reference nor the oracle is built")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scale", "14"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "gnn_forward_edges_per_sec" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "sample" in d["cpu_baseline"] and "workload" in d["config"]
    assert d["config"]["workload"] == "rmat_scale14_ef16" and d["config"]["edges"] > 0   # the arm's own config, not a stand-in
    assert d["cpu_baseline"]["blas_kernel"]                                              # which OpenBLAS kernel set was timed


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert r.stdout.strip() == ""                      # no number without the CUDA path
    assert "no CPU path" in r.stderr or "B200" in r.stderr
