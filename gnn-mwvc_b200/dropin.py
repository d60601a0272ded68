"""The drop-in through the reference's own C++ interface, from Python.

``host/_build/libgvc_dropin.so`` (host/gvc_dropin_capi.cpp) wraps the calls src/GNN_VC.cpp makes --
parse the model text, build a ``reduction_graph``, mutate it as the reductions do, and
``gnn::model::predict(in, out, g)`` -- so that bench.py can time predict() end to end (CSR extraction
from the reduction_graph, upload, forward, scores back in the host matrix) and tools can replay a
solver's sequence of shrinking graphs.  Compute happens in libgvc; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "host" / "_build" / "libgvc_dropin.so"

_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)

REMOVE_NODE, REMOVE_NEIGHBORHOOD, FOLD_NEIGHBORHOOD, FOLD_TWIN, FOLD_ISOLATED, RELABEL, UNDO = range(7)


def _p(a, ty):
    return a.ctypes.data_as(ty)


class Dropin:
    def __init__(self, path: Path | None = None):
        p = Path(path) if path else LIB_PATH
        if not p.exists():
            raise FileNotFoundError(f"{p} missing: built by __graft_entry__.build() where the reference's headers exist")
        L = C.CDLL(str(p))
        L.gvcd_model_create.restype = C.c_void_p
        L.gvcd_model_create.argtypes = [C.c_char_p]
        L.gvcd_model_destroy.argtypes = [C.c_void_p]
        L.gvcd_model_set_weight_scale.argtypes = [C.c_void_p, C.c_float]
        L.gvcd_graph_create.restype = C.c_void_p
        L.gvcd_graph_create.argtypes = [C.c_uint32, C.c_uint64, _u32p, _u32p, _u32p]
        L.gvcd_graph_destroy.argtypes = [C.c_void_p]
        L.gvcd_graph_size.restype = C.c_uint32
        L.gvcd_graph_size.argtypes = [C.c_void_p]
        L.gvcd_graph_mutate.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32]
        L.gvcd_predict.argtypes = [C.c_void_p, C.c_void_p, _f32p, _f32p, C.POINTER(C.c_double)]
        L.gvcd_predict_order.argtypes = [C.c_void_p, C.c_void_p, _f32p, _f32p, _u32p, C.POINTER(C.c_double)]
        self.L = L
        self.last_seconds = 0.0

    def model(self, text: str):
        return self.L.gvcd_model_create(text.encode())

    def model_destroy(self, m):
        self.L.gvcd_model_destroy(m)

    def graph(self, n: int, eu, ev, weights):
        eu = np.ascontiguousarray(eu, np.uint32)
        ev = np.ascontiguousarray(ev, np.uint32)
        weights = np.ascontiguousarray(weights, np.uint32)
        return self.L.gvcd_graph_create(n, len(eu), _p(eu, _u32p), _p(ev, _u32p), _p(weights, _u32p))

    def graph_destroy(self, g):
        self.L.gvcd_graph_destroy(g)

    def graph_size(self, g) -> int:
        return int(self.L.gvcd_graph_size(g))

    def mutate(self, g, op: int, u: int = 0, v: int = 0) -> bool:
        return self.L.gvcd_graph_mutate(g, op, u, v) == 0

    def predict(self, m, g, x, weight_scale: float) -> np.ndarray:
        """gnn::model::predict on the resident graph; wall time of the call in ``last_seconds``."""
        self.L.gvcd_model_set_weight_scale(m, float(weight_scale))
        n = self.graph_size(g)
        x = np.ascontiguousarray(x, np.float32).ravel()
        if x.size != n:
            raise ValueError(f"x has {x.size} entries, the graph {n} vertices")
        out = np.empty(n, np.float32)
        sec = C.c_double()
        rc = self.L.gvcd_predict(m, g, _p(x, _f32p), _p(out, _f32p), C.byref(sec))
        if rc != 0 and n:
            raise RuntimeError("predict left `out` with an unexpected shape")
        self.last_seconds = sec.value
        return out


    def predict_order(self, m, g, x, weight_scale: float):
        """(scores, nodes): predict plus the vertex order the driver derives from it (src/GNN_VC.cpp:186-206),
        sorted from the selection keys the stage-2 kernel left on the device."""
        self.L.gvcd_model_set_weight_scale(m, float(weight_scale))
        n = self.graph_size(g)
        x = np.ascontiguousarray(x, np.float32).ravel()
        out = np.empty(n, np.float32)
        nodes = np.empty(n, np.uint32)
        sec = C.c_double()
        self.L.gvcd_predict_order(m, g, _p(x, _f32p), _p(out, _f32p), _p(nodes, _u32p), C.byref(sec))
        self.last_seconds = sec.value
        return out, nodes


def model_text(layers, name="MWVC_Model") -> str:
    """[(kind, W, bias)] -> the reference's model text (src/gnn_inference.cpp:92-118), %.9g so that
    fp32 weights survive the round trip."""
    from . import capi
    parts = [name, f"{len(layers)} Layers"]
    for kind, W, b in layers:
        if kind == capi.LINEAR:
            parts.append("Linear_Layer")
            parts.append(f"Weights: {W.shape[0]} {W.shape[1]}")
            for row in np.asarray(W, np.float32):
                parts.append(" ".join(f"{float(v):.9g}" for v in row) + " ")
            parts.append("")
            parts.append(f"Bias: 1 {len(b)}")
            parts.append(" ".join(f"{float(v):.9g}" for v in np.asarray(b, np.float32).ravel()) + " ")
            parts.append("")
        elif kind == capi.GRAPH:
            parts.append("Graph_Layer")
        elif kind == capi.RELU:
            parts.append("ReLU_Activation")
        else:
            parts.append("Sigmoid_Activation")
        parts.append("")
    return "\n".join(parts) + "\n"
