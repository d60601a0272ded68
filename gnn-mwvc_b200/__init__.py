"""gnn-mwvc_b200: B200-native GNN forward for GNN_VC (reference: KennethLangedal/GNN-MWVC).

Only what the hot path needs:
  csrc/     hand-written sm_100a kernels + the C ABI (include/gvc.h) -> libgvc.so
  host/     drop-in replacements for the reference's src/gnn_inference.cpp and
            src/matrix.cpp that keep include/gnn_inference.hpp untouched
  capi.py   ctypes binding of the C ABI
  graphs.py synthetic weighted graphs in the reference's input format
  dist.py   vertex-range sharding over several GPUs (one process per GPU)
  build.py  nvcc / g++ recipes

The directory name carries a hyphen; import it as ``gnn_mwvc_b200`` (repo-root shim).
"""
from . import capi, graphs  # noqa: F401
from .capi import Context, GvcError, MODE_EXACT, MODE_FAST  # noqa: F401

__all__ = ["capi", "graphs", "Context", "GvcError", "MODE_EXACT", "MODE_FAST"]
