"""Import shim: the package directory is ``gnn-mwvc_b200/`` (hyphen), which Python
cannot import by name.  ``import gnn_mwvc_b200`` loads that directory as a package."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "gnn-mwvc_b200"
_spec = importlib.util.spec_from_file_location(
    "gnn_mwvc_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gnn_mwvc_b200"] = _mod
_spec.loader.exec_module(_mod)
