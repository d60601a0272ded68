import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def model_layers():
    import gnn_mwvc_b200  # noqa: F401
    from gnn_mwvc_b200 import capi
    return capi.load_model_npz(GOLDEN / "mwvc_model.npz")


@pytest.fixture(scope="session")
def oracle_model(oracle, model_layers):
    from oracle import pyoracle
    return oracle.parse(pyoracle.layers_to_text(model_layers))


@pytest.fixture(scope="session")
def reference():
    """The compiled, unmodified reference (oracle/_ref); absent on a box that never had /root/reference."""
    from oracle import pyoracle
    if not pyoracle.REF_SO.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return pyoracle.Reference(threads=1)
