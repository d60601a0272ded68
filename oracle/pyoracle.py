"""ctypes bindings for the oracle -- TEST INFRASTRUCTURE ONLY.

Two checkers live here:
  * ``Oracle``    -- oracle/_build/libgnnoracle.so, the plain-C restatement
                     (oracle/gnn_oracle.c), single-threaded; builds anywhere.
  * ``Reference`` -- oracle/_ref/libgnnref.so, the UNMODIFIED reference forward
                     (reference src/gnn_inference.cpp + src/matrix.cpp + OpenBLAS
                     0.3.15), built in the container that has /root/reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "_build" / "libgnnoracle.so"
REF_SO = HERE / "_ref" / "libgnnref.so"
REF_BIN = HERE / "_ref" / "GNN_VC_ref"
REF_MODEL_INC = HERE / "_ref" / "model_text.inc"

LINEAR, GRAPH, RELU, SIGMOID = 0, 1, 2, 3

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def build_oracle(force: bool = False) -> Path:
    """Compile the C restatement (gcc only; works on the GPU box too)."""
    src = HERE / "gnn_oracle.c"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(HERE), "oracle"])
    return ORACLE_SO


def build_ref() -> Path | None:
    """Compile the unmodified reference when /root/reference is present."""
    if Path("/root/reference/src/gnn_inference.cpp").exists():
        subprocess.check_call(["make", "-s", "-C", str(HERE), "ref"])
    return REF_SO if REF_SO.exists() else None


def c_unescape(lit: str) -> str:
    """Decode the C string literal of oracle/_ref/model_text.inc."""
    lit = lit.strip()
    assert lit.startswith('"') and lit.endswith('"')
    return lit[1:-1].replace("\\n", "\n").replace('\\"', '"').replace("\\\\", "\\")


def reference_model_text() -> str:
    return c_unescape(REF_MODEL_INC.read_text())


class Oracle:
    """The C restatement."""

    def __init__(self):
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.gvo_model_parse.restype = C.c_void_p
        L.gvo_model_parse.argtypes = [C.c_char_p]
        L.gvo_model_free.argtypes = [C.c_void_p]
        L.gvo_model_num_layers.argtypes = [C.c_void_p]
        L.gvo_model_layer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(_f32p), C.POINTER(_f32p)]
        L.gvo_model_set_weight_scale.argtypes = [C.c_void_p, C.c_float]
        L.gvo_graph_forward.argtypes = [C.c_uint32, _u64p, _u32p, _u32p, _u32p, C.c_float, _f32p,
                                        C.c_int, _f32p]
        L.gvo_linear_forward.argtypes = [C.c_size_t, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p]
        L.gvo_relu_forward.argtypes = [C.c_size_t, _f32p, _f32p]
        L.gvo_sigmoid_forward.argtypes = [C.c_size_t, _f32p, _f32p]
        L.gvo_predict.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u32p, _u32p, _u32p, _f32p,
                                  C.c_int, _f32p, C.POINTER(C.c_int)]
        L.gvo_model_out_width.argtypes = [C.c_void_p, C.c_int]
        L.gvo_dot.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_float, _f32p]
        L.gvo_model_set_graph_layer_scale.argtypes = [C.c_void_p, C.c_int, C.c_float]
        self.L = L

    # -- model ---------------------------------------------------------------
    def parse(self, text: str):
        h = self.L.gvo_model_parse(text.encode())
        assert h, "gvo_model_parse failed"
        return h

    def free(self, h):
        self.L.gvo_model_free(h)

    def layers(self, h):
        """[(kind, W (K x Nout) | None, bias (Nout) | None)]"""
        out = []
        for i in range(self.L.gvo_model_num_layers(h)):
            r, c = C.c_int(), C.c_int()
            W, b = _f32p(), _f32p()
            kind = self.L.gvo_model_layer(h, i, C.byref(r), C.byref(c), C.byref(W), C.byref(b))
            if kind == LINEAR:
                Wm = np.ctypeslib.as_array(W, shape=(r.value, c.value)).copy()
                bv = np.ctypeslib.as_array(b, shape=(c.value,)).copy()
                out.append((kind, Wm, bv))
            else:
                out.append((kind, None, None))
        return out

    def model_from_layers(self, layers) -> str:
        """Serialise [(kind, W, bias)] into the reference text format with
        enough digits (%.9g) to round-trip fp32 exactly."""
        return layers_to_text(layers)

    # -- layers ----------------------------------------------------------------
    def graph_forward(self, row_ptr, col, W, NW, scale, x):
        n, w = x.shape
        out = np.empty((n, 2 * w + 3), np.float32)
        x = np.ascontiguousarray(x, np.float32)
        self.L.gvo_graph_forward(n, _p(row_ptr, _u64p), _p(col, _u32p), _p(W, _u32p), _p(NW, _u32p),
                                 float(scale), _p(x, _f32p), w, _p(out, _f32p))
        return out

    def linear_forward(self, x, Wm, bias):
        n, K = x.shape
        Nout = Wm.shape[1]
        out = np.empty((n, Nout), np.float32)
        x = np.ascontiguousarray(x, np.float32)
        Wm = np.ascontiguousarray(Wm, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        self.L.gvo_linear_forward(n, K, Nout, _p(x, _f32p), _p(Wm, _f32p), _p(bias, _f32p), _p(out, _f32p))
        return out

    def relu(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self.L.gvo_relu_forward(x.size, _p(x, _f32p), _p(out, _f32p))
        return out

    def sigmoid(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self.L.gvo_sigmoid_forward(x.size, _p(x, _f32p), _p(out, _f32p))
        return out

    def dot(self, A, B, C0=None, at=False, bt=False, beta=0.0):
        """dot() of matrix.cpp:106-122 in the reference's operation order (A, B as stored)."""
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        m, k = (A.shape[1], A.shape[0]) if at else A.shape
        n = B.shape[0] if bt else B.shape[1]
        out = np.zeros((m, n), np.float32) if C0 is None else np.ascontiguousarray(C0, np.float32).copy()
        self.L.gvo_dot(int(at), int(bt), m, n, k, _p(A, _f32p), _p(B, _f32p), float(beta), _p(out, _f32p))
        return out

    def set_graph_layer_scale(self, h, which: int, scale: float):
        assert self.L.gvo_model_set_graph_layer_scale(h, which, float(scale)) == 0

    def predict(self, h, row_ptr, col, W, NW, x, scale=None):
        if scale is not None:
            self.L.gvo_model_set_weight_scale(h, float(scale))
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim == 1:
            x = x[:, None]
        n, w = x.shape
        ow = self.L.gvo_model_out_width(h, w)
        out = np.empty((n, ow), np.float32)
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        rc = self.L.gvo_predict(h, n, _p(row_ptr, _u64p), _p(col, _u32p), _p(W, _u32p), _p(NW, _u32p),
                                _p(x, _f32p), w, _p(out, _f32p), None)
        assert rc == 0
        return out


DROPIN_SO = HERE / "_ref" / "libgnndropin.so"


def build_dropin_harness() -> Path | None:
    """oracle/_ref/libgnndropin.so: ref_harness.cpp over the drop-in host units (needs /root/reference
    for the headers and gnn-mwvc_b200/libgvc.so)."""
    if not Path("/root/reference/include/gnn_inference.hpp").exists():
        return DROPIN_SO if DROPIN_SO.exists() else None
    subprocess.check_call(["make", "-s", "-C", str(HERE), "dropin_harness"])
    return DROPIN_SO


class Reference:
    """The unmodified reference (OpenBLAS kernel pinned to Prescott)."""

    def __init__(self, threads: int | None = None, so: Path | None = None, coretype: str | None = "Prescott"):
        """coretype: the OpenBLAS kernel set.  "Prescott" (default) is the one the oracle restates and
        the golden vectors were made with -- every CHECKER use keeps it.  None leaves the choice to
        OpenBLAS' own CPU detection (the TIMED reference arm of bench.py: the kernel a stock build
        would run on this host).  Fixed at the first load of libopenblas in a process."""
        REF_SO = so or globals()["REF_SO"]
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        # must be set before libopenblas' constructor runs
        if coretype is None:
            os.environ.pop("OPENBLAS_CORETYPE", None)
        else:
            os.environ.setdefault("OPENBLAS_CORETYPE", coretype)
        if threads:
            os.environ["OPENBLAS_NUM_THREADS"] = str(threads)
        L = C.CDLL(str(REF_SO))
        L.ref_blas_config.restype = C.c_char_p
        L.ref_blas_threads.argtypes = [C.c_int]
        L.ref_model_create.restype = C.c_void_p
        L.ref_model_create.argtypes = [C.c_char_p]
        L.ref_model_destroy.argtypes = [C.c_void_p]
        L.ref_model_build.restype = C.c_void_p
        L.ref_model_build.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(_f32p), C.POINTER(_f32p), _f32p]
        L.ref_model_set_weight_scale.argtypes = [C.c_void_p, C.c_float]
        L.ref_model_text.restype = C.c_size_t
        L.ref_model_text.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ref_predict.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, _u32p, _u32p, _u32p, _f32p, _f32p,
                                  _u64p, _u32p, _u32p, C.c_int, C.POINTER(C.c_double)]
        L.ref_graph_layer.argtypes = [C.c_uint32, C.c_uint64, _u32p, _u32p, _u32p, C.c_float, _f32p,
                                      C.c_int, _f32p]
        L.ref_linear_layer.argtypes = [C.c_size_t, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p]
        L.ref_relu.argtypes = [C.c_size_t, _f32p, _f32p]
        L.ref_sigmoid.argtypes = [C.c_size_t, _f32p, _f32p]
        L.ref_linear_init.argtypes = [C.c_int, C.c_int, C.c_size_t, _f32p]
        L.ref_dot.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_float, _f32p]
        if hasattr(L, "ref_parse_graph"):          # only in the build over the reference's sources
            L.ref_parse_graph.argtypes = [C.c_char_p, _u64p, _u64p, _u32p, _u32p, _u32p]
        L.ref_graph_create.restype = C.c_void_p
        L.ref_graph_create.argtypes = [C.c_uint32, C.c_uint64, _u32p, _u32p, _u32p]
        L.ref_graph_destroy.argtypes = [C.c_void_p]
        L.ref_graph_size.restype = C.c_uint32
        L.ref_graph_size.argtypes = [C.c_void_p]
        L.ref_graph_mutate.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32]
        L.ref_graph_csr.restype = C.c_uint64
        L.ref_graph_csr.argtypes = [C.c_void_p, _u64p, _u32p, _u32p, _u32p, C.POINTER(C.c_uint8)]
        L.ref_predict_on.argtypes = [C.c_void_p, C.c_void_p, _f32p, _f32p, C.c_int, C.POINTER(C.c_double)]
        L.ref_selection_order.argtypes = [C.c_void_p, C.c_void_p, _f32p, _f32p, _u32p]
        self.L = L
        if threads:
            L.ref_blas_threads(threads)

    def blas_config(self) -> str:
        return self.L.ref_blas_config().decode()

    def model(self, text: str):
        return self.L.ref_model_create(text.encode())

    def model_build(self, layers, scales=None):
        """A model assembled with add_layer: layers = [(kind, W, bias)], scales[i] = WEIGHT_SCALE of
        layer i where it is a graph layer (default 120, the header's)."""
        n = len(layers)
        kinds = (C.c_int * n)(*[int(k) for k, _, _ in layers])
        rows, cols = (C.c_int * n)(), (C.c_int * n)()
        Wp, bp = (_f32p * n)(), (_f32p * n)()
        sc = np.ascontiguousarray(scales if scales is not None else [120.0] * n, np.float32)
        keep = []
        for i, (k, W, b) in enumerate(layers):
            if k == LINEAR:
                W = np.ascontiguousarray(W, np.float32)
                b = np.ascontiguousarray(b, np.float32).ravel()
                rows[i], cols[i] = W.shape
                Wp[i], bp[i] = W.ctypes.data_as(_f32p), b.ctypes.data_as(_f32p)
                keep += [W, b]
        return self.L.ref_model_build(n, kinds, rows, cols, Wp, bp, _p(sc, _f32p))

    def predict_on_as_is(self, h, gh, x):
        """predict on a resident graph WITHOUT touching the model's weight scales."""
        n = self.graph_size(gh)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(n, np.float32)
        sec = C.c_double()
        rc = self.L.ref_predict_on(h, gh, _p(x, _f32p), _p(out, _f32p), 1, C.byref(sec))
        assert rc == 0 or n == 0
        self.last_seconds = sec.value
        return out

    def model_text(self, h) -> str:
        n = self.L.ref_model_text(h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.ref_model_text(h, buf, n + 1)
        return buf.value.decode()

    def destroy(self, h):
        self.L.ref_model_destroy(h)

    def predict(self, h, n, eu, ev, weights, x, scale, want_csr=False, reps=1):
        """eu<ev sorted unique undirected edges.  Returns scores[, (row_ptr, col, NW)][, seconds]."""
        self.L.ref_model_set_weight_scale(h, float(scale))
        eu = np.ascontiguousarray(eu, np.uint32)
        ev = np.ascontiguousarray(ev, np.uint32)
        weights = np.ascontiguousarray(weights, np.uint32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(n, np.float32)
        rp = col = nw = None
        if want_csr:
            rp = np.empty(n + 1, np.uint64)
            col = np.empty(2 * len(eu), np.uint32)
            nw = np.empty(n, np.uint32)
        sec = C.c_double()
        rc = self.L.ref_predict(h, n, len(eu), _p(eu, _u32p), _p(ev, _u32p), _p(weights, _u32p),
                                _p(x, _f32p), _p(out, _f32p), _p(rp, _u64p), _p(col, _u32p),
                                _p(nw, _u32p), reps, C.byref(sec))
        assert rc == 0 or n == 0, "reference predict returned an unexpected shape"
        self.last_seconds = sec.value
        if want_csr:
            return out, (rp, col, nw)
        return out

    # -- a graph that lives across calls (as GNN_VC's does) ---------------------------------------
    REMOVE_NODE, REMOVE_NEIGHBORHOOD, FOLD_NEIGHBORHOOD, FOLD_TWIN, FOLD_ISOLATED, RELABEL, UNDO = range(7)

    def graph_create(self, n, eu, ev, weights):
        eu = np.ascontiguousarray(eu, np.uint32)
        ev = np.ascontiguousarray(ev, np.uint32)
        weights = np.ascontiguousarray(weights, np.uint32)
        return self.L.ref_graph_create(n, len(eu), _p(eu, _u32p), _p(ev, _u32p), _p(weights, _u32p))

    def graph_destroy(self, gh):
        self.L.ref_graph_destroy(gh)

    def graph_size(self, gh) -> int:
        return int(self.L.ref_graph_size(gh))

    def graph_mutate(self, gh, op: int, u: int = 0, v: int = 0) -> bool:
        """One reduction_graph mutator (see REMOVE_NODE ...); False if its precondition does not hold."""
        return self.L.ref_graph_mutate(gh, op, u, v) == 0

    def graph_csr(self, gh):
        """(row_ptr u64, col u32, W u32, NW u32, active u8) as predict reads them: begin/end/W/NW."""
        n = self.graph_size(gh)
        nnz = int(self.L.ref_graph_csr(gh, None, None, None, None, None))
        rp, col = np.empty(n + 1, np.uint64), np.empty(max(nnz, 1), np.uint32)
        w, nw, act = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint8)
        self.L.ref_graph_csr(gh, _p(rp, _u64p), _p(col, _u32p), _p(w, _u32p), _p(nw, _u32p),
                             act.ctypes.data_as(C.POINTER(C.c_uint8)))
        return rp, col[:nnz], w[:n], nw[:n], act[:n]

    def predict_on(self, h, gh, x, scale, reps=1):
        """model::predict on the resident graph; wall time of predict() alone in self.last_seconds."""
        self.L.ref_model_set_weight_scale(h, float(scale))
        n = self.graph_size(gh)
        x = np.ascontiguousarray(x, np.float32)
        assert x.size == n
        out = np.empty(n, np.float32)
        sec = C.c_double()
        rc = self.L.ref_predict_on(h, gh, _p(x, _f32p), _p(out, _f32p), reps, C.byref(sec))
        assert rc == 0 or n == 0, "predict returned an unexpected shape"
        self.last_seconds = sec.value
        return out

    def parse_graph(self, path):
        """The reference's own parse_graph (src/GNN_VC.cpp:34-91): (n, weights, eu, ev)."""
        n, e = C.c_uint64(), C.c_uint64()
        self.L.ref_parse_graph(str(path).encode(), C.byref(n), C.byref(e), None, None, None)
        w = np.zeros(max(n.value, 1), np.uint32)
        eu, ev = np.zeros(max(e.value, 1), np.uint32), np.zeros(max(e.value, 1), np.uint32)
        self.L.ref_parse_graph(str(path).encode(), None, None, _p(w, _u32p), _p(eu, _u32p), _p(ev, _u32p))
        return int(n.value), w[:n.value], eu[:e.value], ev[:e.value]

    def selection_order(self, h, gh, x, scale):
        """(scores, nodes): predict and the driver's sort of the vertices (src/GNN_VC.cpp:186-206)."""
        self.L.ref_model_set_weight_scale(h, float(scale))
        n = self.graph_size(gh)
        x = np.ascontiguousarray(x, np.float32)
        out, nodes = np.empty(n, np.float32), np.empty(n, np.uint32)
        self.L.ref_selection_order(h, gh, _p(x, _f32p), _p(out, _f32p), _p(nodes, _u32p))
        return out, nodes

    def graph_layer(self, n, eu, ev, weights, scale, x):
        x = np.ascontiguousarray(x, np.float32)
        w = x.shape[1]
        out = np.empty((n, 2 * w + 3), np.float32)
        eu = np.ascontiguousarray(eu, np.uint32)
        ev = np.ascontiguousarray(ev, np.uint32)
        weights = np.ascontiguousarray(weights, np.uint32)
        self.L.ref_graph_layer(n, len(eu), _p(eu, _u32p), _p(ev, _u32p), _p(weights, _u32p),
                               float(scale), _p(x, _f32p), w, _p(out, _f32p))
        return out

    def linear_layer(self, x, Wm, bias):
        x = np.ascontiguousarray(x, np.float32)
        Wm = np.ascontiguousarray(Wm, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        n, K = x.shape
        out = np.empty((n, Wm.shape[1]), np.float32)
        self.L.ref_linear_layer(n, K, Wm.shape[1], _p(x, _f32p), _p(Wm, _f32p), _p(bias, _f32p),
                                _p(out, _f32p))
        return out

    def relu(self, x):
        x = np.ascontiguousarray(x, np.float32).ravel()
        out = np.empty_like(x)
        self.L.ref_relu(x.size, _p(x, _f32p), _p(out, _f32p))
        return out

    def sigmoid(self, x):
        x = np.ascontiguousarray(x, np.float32).ravel()
        out = np.empty_like(x)
        self.L.ref_sigmoid(x.size, _p(x, _f32p), _p(out, _f32p))
        return out

    def dot(self, A, B, C0=None, at=False, bt=False, beta=0.0):
        """The reference's dot(): C = op(A) op(B) + beta C0 (A, B as stored: transposed when the flag is set)."""
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        m, k = (A.shape[1], A.shape[0]) if at else A.shape
        n = B.shape[0] if bt else B.shape[1]
        out = np.zeros((m, n), np.float32) if C0 is None else np.ascontiguousarray(C0, np.float32).copy()
        self.L.ref_dot(int(at), int(bt), m, n, k, _p(A, _f32p), _p(B, _f32p), float(beta), _p(out, _f32p))
        return out

    def linear_init(self, K, Nout, seed):
        out = np.empty(K * Nout + Nout, np.float32)
        self.L.ref_linear_init(K, Nout, seed, _p(out, _f32p))
        return out[: K * Nout].reshape(K, Nout), out[K * Nout:]


def layers_to_text(layers, name="MWVC_Model") -> str:
    """[(kind, W, bias)] -> reference text format (gnn_inference.cpp:92-118 layout)."""
    parts = [name, f"{len(layers)} Layers"]
    for kind, W, b in layers:
        if kind == LINEAR:
            parts.append("Linear_Layer")
            parts.append(f"Weights: {W.shape[0]} {W.shape[1]}")
            for row in np.asarray(W, np.float32):
                parts.append(" ".join(f"{float(v):.9g}" for v in row) + " ")
            parts.append("")
            parts.append(f"Bias: 1 {len(b)}")
            parts.append(" ".join(f"{float(v):.9g}" for v in np.asarray(b, np.float32)) + " ")
            parts.append("")
        elif kind == GRAPH:
            parts.append("Graph_Layer")
        elif kind == RELU:
            parts.append("ReLU_Activation")
        else:
            parts.append("Sigmoid_Activation")
        parts.append("")
    return "\n".join(parts) + "\n"


# ---- training path -------------------------------------------------------------------
TRAIN_REF_SO = HERE / "_ref" / "libgnntrainref.so"
TRAIN_DROPIN_SO = HERE / "_ref" / "libgnntraindropin.so"


def build_train_harness() -> tuple[Path | None, Path | None]:
    """oracle/_ref/libgnntrainref.so (the reference's old_files/src/lib/gnn_training.cpp, unmodified) and
    libgnntraindropin.so (the drop-in host units), both behind oracle/train_harness.cpp."""
    if Path("/root/reference/old_files/src/lib/gnn_training.cpp").exists():
        subprocess.check_call(["make", "-s", "-C", str(HERE), "train_ref", "train_dropin"])
    return (TRAIN_REF_SO if TRAIN_REF_SO.exists() else None, TRAIN_DROPIN_SO if TRAIN_DROPIN_SO.exists() else None)


class TrainHarness:
    """The reference's training interface (old_files/include/gnn/gnn_training.hpp) through
    oracle/train_harness.cpp: over the reference's own sources (default) or over the drop-in (`dropin=True`)."""

    def __init__(self, dropin: bool = False, threads: int | None = 1):
        so = TRAIN_DROPIN_SO if dropin else TRAIN_REF_SO
        if not so.exists():
            raise FileNotFoundError(f"{so} missing: run `make -C oracle train_ref train_dropin` where /root/reference exists")
        if not dropin:
            os.environ.setdefault("OPENBLAS_CORETYPE", "Prescott")     # the kernel set the checker is pinned to (App. B)
        L = C.CDLL(str(so))
        pp = C.POINTER(C.POINTER(C.c_float))
        L.trn_create.restype = C.c_void_p
        L.trn_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), pp, pp, _f32p, C.POINTER(C.c_size_t)]
        L.trn_parse.restype = C.c_void_p
        L.trn_parse.argtypes = [C.c_char_p]
        L.trn_destroy.argtypes = [C.c_void_p]
        L.trn_text.restype = C.c_size_t
        L.trn_text.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.trn_set_graph.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, _u32p, _u32p, _u32p]
        L.trn_predict.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, C.c_int]
        L.trn_backprop.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p]
        L.trn_mse_step.restype = C.c_float
        L.trn_mse_step.argtypes = [C.c_void_p, _f32p, C.c_int]
        L.trn_sgd_step.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float]
        L.trn_zero_grad.argtypes = [C.c_void_p]
        L.trn_read.restype = C.c_size_t
        L.trn_read.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]
        L.trn_linear_layer.argtypes = [C.c_size_t, C.c_int, C.c_int] + [_f32p] * 8
        L.trn_graph_layer.argtypes = [C.c_void_p, C.c_int, C.c_float] + [_f32p] * 4
        L.trn_activation.argtypes = [C.c_int, C.c_size_t] + [_f32p] * 4
        L.trn_mse.restype = C.c_float
        L.trn_mse.argtypes = [C.c_size_t, C.c_int, _f32p, _f32p, _f32p]
        L.trn_blas_threads.argtypes = [C.c_int]
        if threads:
            L.trn_blas_threads(threads)
        self.L = L

    def create(self, layers, scales=None, seeds=None):
        """layers: [(kind, W | (rows, cols), bias | None)]; linear layers without W keep the seeded random init."""
        n = len(layers)
        kinds = (C.c_int * n)(*[int(k) for k, _, _ in layers])
        rows, cols = (C.c_int * n)(), (C.c_int * n)()
        Wp, bp = (C.POINTER(C.c_float) * n)(), (C.POINTER(C.c_float) * n)()
        keep = []
        for i, (k, W, b) in enumerate(layers):
            if k != LINEAR:
                continue
            if isinstance(W, tuple):
                rows[i], cols[i] = W
                continue
            W = np.ascontiguousarray(W, np.float32)
            b = np.ascontiguousarray(b, np.float32).ravel()
            rows[i], cols[i] = W.shape
            keep += [W, b]
            Wp[i], bp[i] = _p(W, _f32p), _p(b, _f32p)
        sc = np.ascontiguousarray(scales if scales is not None else np.full(n, 1200.0), np.float32)
        sd = (C.c_size_t * n)(*[int(s) for s in (seeds if seeds is not None else range(n))])
        h = self.L.trn_create(n, kinds, rows, cols, Wp, bp, _p(sc, _f32p), sd)
        self._shapes = {i: (int(rows[i]), int(cols[i])) for i in range(n) if layers[i][0] == LINEAR}
        return h

    def parse(self, text: str):
        return self.L.trn_parse(text.encode())

    def destroy(self, h):
        self.L.trn_destroy(h)

    def text(self, h) -> str:
        n = self.L.trn_text(h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.trn_text(h, buf, n + 1)
        return buf.value.decode()

    def set_graph(self, h, n, eu, ev, W):
        eu, ev, W = (np.ascontiguousarray(a, np.uint32) for a in (eu, ev, W))
        self._n = n
        self.L.trn_set_graph(h, n, len(eu), _p(eu, _u32p), _p(ev, _u32p), _p(W, _u32p))

    def predict(self, h, x, out_w=1):
        x = np.ascontiguousarray(x, np.float32).reshape(self._n, -1)
        out = np.empty((self._n, out_w), np.float32)
        rc = self.L.trn_predict(h, _p(x, _f32p), x.shape[1], _p(out, _f32p), out_w)
        if rc:
            raise RuntimeError("predict left another shape")
        return out

    def backprop(self, h, grad, in_w=1):
        grad = np.ascontiguousarray(grad, np.float32).reshape(self._n, -1)
        gx = np.empty((self._n, in_w), np.float32)
        w = self.L.trn_backprop(h, _p(grad, _f32p), grad.shape[1], _p(gx, _f32p))
        assert w == in_w, (w, in_w)
        return gx

    def mse_step(self, h, y):
        y = np.ascontiguousarray(y, np.float32).reshape(self._n, -1)
        return float(self.L.trn_mse_step(h, _p(y, _f32p), y.shape[1]))

    def sgd_step(self, h, batch, lr=0.1, momentum=0.9, wd=0.0):
        self.L.trn_sgd_step(h, batch, lr, momentum, wd)

    def zero_grad(self, h):
        self.L.trn_zero_grad(h)

    def read(self, h, what, layer, shape):
        W, b = np.zeros(shape, np.float32), np.zeros(shape[1], np.float32)
        k = self.L.trn_read(h, what, layer, _p(W, _f32p), _p(b, _f32p))
        return (W, b) if k else (None, None)

    def linear_layer(self, W, bias, x, grad):
        W, x, grad = (np.ascontiguousarray(a, np.float32) for a in (W, x, grad))
        bias = np.ascontiguousarray(bias, np.float32).ravel()
        n, (K, N) = x.shape[0], W.shape
        out, gW, gb, gi = np.empty((n, N), np.float32), np.empty((K, N), np.float32), np.empty(N, np.float32), np.empty((n, K), np.float32)
        self.L.trn_linear_layer(n, K, N, *(_p(a, _f32p) for a in (W, bias, x, grad, out, gW, gb, gi)))
        return out, gW, gb, gi

    def graph_layer(self, h, x, grad, scale):
        x, grad = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(grad, np.float32)
        w = x.shape[1]
        out, gi = np.empty((self._n, 2 * w + 3), np.float32), np.empty((self._n, w), np.float32)
        self.L.trn_graph_layer(h, w, scale, *(_p(a, _f32p) for a in (x, grad, out, gi)))
        return out, gi

    def activation(self, kind, z, grad):
        z, grad = np.ascontiguousarray(z, np.float32).ravel(), np.ascontiguousarray(grad, np.float32).ravel()
        out, gi = np.empty_like(z), np.empty_like(z)
        self.L.trn_activation(kind, z.size, *(_p(a, _f32p) for a in (z, grad, out, gi)))
        return out, gi

    def mse(self, x, y):
        x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
        g = np.empty_like(x)
        loss = self.L.trn_mse(x.shape[0], x.shape[1], _p(x, _f32p), _p(y, _f32p), _p(g, _f32p))
        return float(loss), g


def train_backward_numpy(layers, scales, row_ptr, col, W, NW, x, grad_out, dtype=np.float64):
    """CPU restatement of model_training::predict + ::backprop (old_files/src/lib/gnn_training.cpp:17-129) in
    numpy -- TEST INFRASTRUCTURE.  Forward and backward in `dtype` (float64: what the fp32 results are
    compared with under a tolerance; the fp32 forward itself is checked bit for bit elsewhere).
    Returns (out, grad_x, {layer index: (grad_W, grad_bias)})."""
    n = len(row_ptr) - 1
    deg = np.diff(row_ptr).astype(np.int64)
    src = np.repeat(np.arange(n), deg)                        # row of every adjacency entry
    a = np.asarray(x, dtype).reshape(n, -1)
    saved = []
    gi = 0
    for kind, Wm, b in layers:
        saved.append(a)
        if kind == LINEAR:
            a = a @ np.asarray(Wm, dtype) + np.asarray(b, dtype).reshape(1, -1)
        elif kind == GRAPH:
            w = a.shape[1]
            s = dtype(np.float32(scales[min(gi, len(scales) - 1)]))
            gi += 1
            out = np.zeros((n, 2 * w + 3), dtype)
            np.add.at(out[:, :w], src, a[col])                # :33-36
            out[:, w:2 * w] = a                               # :37
            out[:, w + 1] = deg                               # :38-40 (the column quirk)
            out[:, w + 2] = W.astype(dtype) / s
            out[:, w + 3] = NW.astype(dtype) / s
            a = out
        elif kind == RELU:
            a = np.maximum(a, 0)
        else:
            a = 1.0 / (1.0 + np.exp(-a))
    out = a
    g = np.asarray(grad_out, dtype).reshape(n, -1)
    grads = {}
    for i in range(len(layers) - 1, -1, -1):
        kind, Wm, b = layers[i]
        z = saved[i]
        if kind == LINEAR:
            grads[i] = (z.T @ g, g.sum(axis=0))               # :19-22
            g = g @ np.asarray(Wm, dtype).T                   # :25
        elif kind == GRAPH:
            w = z.shape[1]
            gn = np.zeros((n, w), dtype)
            np.add.at(gn, src, g[col][:, :w])                 # :36-39: grad_out[u] += grad_in[v][0..w) over v in N(u)
            gn += g[:, w:2 * w]                               # :40
            g = gn
        elif kind == RELU:
            g = np.where(z >= 0, g, 0)                        # :52
        else:
            f = 1.0 / (1.0 + np.exp(-z))
            g = f * (1.0 - f) * g                             # :64
    return out, g, grads
