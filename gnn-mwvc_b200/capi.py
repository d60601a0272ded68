"""ctypes binding of libgvc.so (include/gvc.h) -- the product's only compute path.

There is deliberately no fallback: if the shared library is missing, or no B200
is visible, importing works but every compute call raises.  Nothing here touches
``oracle/``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["GVC_LIB"]) if os.environ.get("GVC_LIB") else PKG_DIR / "libgvc.so"     # GVC_LIB: an experimental build

LINEAR, GRAPH, RELU, SIGMOID = 0, 1, 2, 3
MODE_EXACT, MODE_FAST = 0, 1

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int)

# every symbol include/gvc.h declares: (restype, argtypes)
SIGNATURES = {
    "gvc_last_error": (C.c_char_p, []),
    "gvc_abi_version": (C.c_int, []),
    "gvc_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "gvc_ctx_destroy": (None, [C.c_void_p]),
    "gvc_model_upload": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, C.POINTER(_f32p), C.POINTER(_f32p)]),
    "gvc_model_is_fused": (C.c_int, [C.c_void_p]),
    "gvc_model_weight_scales": (C.c_int, [C.c_void_p, C.c_int, _f32p]),
    "gvc_graph_upload": (C.c_int, [C.c_void_p, C.c_uint32, _u64p, _u32p, _u32p, _u32p]),
    "gvc_graph_upload_shard": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _u32p, _u32p, _u32p]),
    "gvc_graph_staging": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(_u64p), C.POINTER(_u32p),
                                    C.POINTER(_u32p), C.POINTER(_u32p)]),
    "gvc_ctx_warm": (C.c_int, [C.c_void_p, C.c_uint64]),
    # training path
    "gvc_trainer_create": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, C.POINTER(_f32p), C.POINTER(_f32p), C.POINTER(C.c_void_p)]),
    "gvc_trainer_destroy": (None, [C.c_void_p]),
    "gvc_trainer_input_width": (C.c_int, [C.c_void_p]),
    "gvc_trainer_output_width": (C.c_int, [C.c_void_p]),
    "gvc_trainer_predict": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int, _f32p, C.c_int]),
    "gvc_trainer_backprop": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int]),
    "gvc_trainer_mse_backprop": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int]),
    "gvc_trainer_sgd_step": (C.c_int, [C.c_void_p, C.c_uint64, C.c_float, C.c_float, C.c_float]),
    "gvc_trainer_zero_grad": (C.c_int, [C.c_void_p]),
    "gvc_trainer_read": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]),
    "gvc_trainer_write": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]),
    "gvc_linear_backward_host": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int]),
    "gvc_graph_backward_host": (C.c_int, [C.c_void_p, _f32p, C.c_int, _f32p]),
    "gvc_relu_backward_host": (C.c_int, [C.c_void_p, C.c_uint64, _f32p, _f32p, _f32p]),
    "gvc_sigmoid_backward_host": (C.c_int, [C.c_void_p, C.c_uint64, _f32p, _f32p, _f32p, C.c_int]),
    "gvc_mse_host": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, _f32p, _f32p, _f32p, _f32p]),
    "gvc_sgd_host": (C.c_int, [C.c_void_p, C.c_uint64, _f32p, _f32p, _f32p, C.c_uint64, C.c_float, C.c_float, C.c_float]),
    "gvc_graph_upload_stream": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "gvc_graph_upload_stream_x": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gvc_graph_adopt_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gvc_graph_set_tail": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32]),
    "gvc_peer_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    "gvc_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "gvc_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gvc_peer_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gvc_stage_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "gvc_peer_owners": (C.c_int, [C.c_void_p, C.c_int, _u32p, _i32p]),
    "gvc_forward": (C.c_int, [C.c_void_p, _f32p, C.c_float, _f32p, C.c_int]),
    "gvc_forward_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int]),
    "gvc_forward_keys": (C.c_int, [C.c_void_p, _f32p, C.c_float, _f32p, _f32p, C.POINTER(C.c_ubyte), C.c_int]),
    "gvc_last_keys": (C.c_int, [C.c_void_p, _f32p, C.POINTER(C.c_ubyte)]),
    "gvc_forward_device_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "gvc_stage_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int]),
    "gvc_graph_layer_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float]),
    "gvc_linear_layer_device": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "gvc_relu_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "gvc_sigmoid_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "gvc_graph_layer_host": (C.c_int, [C.c_void_p, _f32p, C.c_int, _f32p, C.c_float]),
    "gvc_linear_host": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_int]),
    "gvc_relu_host": (C.c_int, [C.c_void_p, C.c_uint64, _f32p, _f32p]),
    "gvc_sigmoid_host": (C.c_int, [C.c_void_p, C.c_uint64, _f32p, _f32p, C.c_int]),
    "gvc_sgemm_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, _f32p, C.c_uint64,
                                 _f32p, C.c_uint64, C.c_float, _f32p, C.c_uint64]),
    "gvc_group_create": (C.c_int, [C.POINTER(C.c_void_p), _i32p, C.c_int]),
    "gvc_group_destroy": (None, [C.c_void_p]),
    "gvc_group_size": (C.c_int, [C.c_void_p]),
    "gvc_group_model_upload": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, C.POINTER(_f32p), C.POINTER(_f32p)]),
    "gvc_group_model_weight_scales": (C.c_int, [C.c_void_p, C.c_int, _f32p]),
    "gvc_group_graph_upload": (C.c_int, [C.c_void_p, C.c_uint32, _u64p, _u32p, _u32p, _u32p]),
    "gvc_group_bounds": (C.c_int, [C.c_void_p, _u32p]),
    "gvc_group_forward": (C.c_int, [C.c_void_p, _f32p, C.c_float, _f32p, C.c_int]),
    "gvc_metis_parse": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "gvc_metis_free": (None, [C.c_void_p]),
    "gvc_metis_vertices": (C.c_uint64, [C.c_void_p]),
    "gvc_metis_edges": (C.c_uint64, [C.c_void_p]),
    "gvc_metis_weights": (_u32p, [C.c_void_p]),
    "gvc_metis_edge_u": (_u32p, [C.c_void_p]),
    "gvc_metis_edge_v": (_u32p, [C.c_void_p]),
    "gvc_metis_csr": (C.c_int, [C.c_void_p, _u64p, _u32p, _u32p]),
    "gvc_stream": (C.c_void_p, [C.c_void_p]),
    "gvc_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gvc_sync": (C.c_int, [C.c_void_p]),
    "gvc_launch_count": (C.c_uint64, [C.c_void_p]),
    "gvc_debug_h": (C.c_void_p, [C.c_void_p, C.c_int]),
    "gvc_debug_row_order": (C.c_int, [C.c_void_p, _u32p]),
    "gvc_debug_px": (C.c_int, [C.c_void_p, _u32p]),
}


class GvcError(RuntimeError):
    pass


_lib = None


def load_library(path: Path | None = None):
    """dlopen libgvc.so and type every symbol of include/gvc.h.  Raises if missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise GvcError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, ty):
    return a.ctypes.data_as(ty) if a is not None and a.size else None


def _dptr(t):
    """Device pointer of a torch tensor (or a raw int)."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class Context:
    """One gvc_ctx: a device, its stream, one model and one graph shard."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        self._check(self.lib.gvc_ctx_create(C.byref(h), device))
        self.h = h
        self.device = device
        self.n_global = 0
        self.v_begin = 0
        self.v_end = 0
        self._keep = []      # adopted torch tensors must outlive the context's use of them

    def _check(self, rc):
        if rc != 0:
            raise GvcError(f"libgvc error {rc}: {self.lib.gvc_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.gvc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- model ---------------------------------------------------------------
    def model_upload(self, layers):
        """layers: [(kind, W (K x Nout) | None, bias (Nout) | None)]"""
        n = len(layers)
        kinds = (C.c_int * n)(*[int(k) for k, _, _ in layers])
        rows = (C.c_int * n)()
        cols = (C.c_int * n)()
        Wp = (_f32p * n)()
        bp = (_f32p * n)()
        keep = []
        for i, (k, W, b) in enumerate(layers):
            if k == LINEAR:
                W = _np(W, np.float32)
                b = _np(b, np.float32).ravel()
                rows[i], cols[i] = W.shape
                Wp[i] = W.ctypes.data_as(_f32p)
                bp[i] = b.ctypes.data_as(_f32p)
                keep += [W, b]
        self._check(self.lib.gvc_model_upload(self.h, n, kinds, rows, cols, Wp, bp))
        self.layers = layers

    @property
    def fused(self) -> bool:
        return bool(self.lib.gvc_model_is_fused(self.h))

    def weight_scales(self, scales):
        """Per-graph-layer WEIGHT_SCALE (None / empty: every layer uses the forward call's scalar)."""
        a = np.ascontiguousarray(scales if scales is not None else [], np.float32)
        self._check(self.lib.gvc_model_weight_scales(self.h, a.size, _ptr(a, _f32p)))

    # -- graph ---------------------------------------------------------------
    def graph_upload(self, row_ptr, col, W, NW, n_global=None, v_begin=0, v_end=None):
        row_ptr = _np(row_ptr, np.uint64)
        col = _np(col, np.uint32)
        W = _np(W, np.uint32)
        NW = _np(NW, np.uint32)
        nl = len(row_ptr) - 1
        if n_global is None:
            n_global = nl
        if v_end is None:
            v_end = v_begin + nl
        self._check(self.lib.gvc_graph_upload_shard(self.h, n_global, v_begin, v_end, _ptr(row_ptr, _u64p),
                                                    _ptr(col, _u32p), _ptr(W, _u32p), _ptr(NW, _u32p)))
        self.n_global, self.v_begin, self.v_end = n_global, v_begin, v_end

    def graph_upload_ranges(self, span, begin, end, W, NW, n_threads: int = 0, x=None):
        """gvc_graph_upload_stream(_x) from numpy arrays: `span` is the raw edge array (holes allowed),
        begin/end the per-vertex ranges into it.  The callbacks copy slices into libgvc's pinned slots
        (possibly from several of its worker threads).  With `x` the forward's input travels along
        (then ``forward(None, ...)`` uses it)."""
        span = _np(span, np.uint32)
        begin, end = _np(begin, np.uint32), _np(end, np.uint32)
        W, NW = _np(W, np.uint32), _np(NW, np.uint32)
        n = len(begin)
        FV = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, _u32p, _u32p, _u32p, _u32p)
        FS = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, _u32p)

        def fill_vertices(_user, first, count, b, e, w, nw):
            for dst, src in ((b, begin), (e, end), (w, W), (nw, NW)):
                np.ctypeslib.as_array(dst, shape=(count,))[:] = src[first:first + count]

        def fill_span(_user, offset, count, dst):
            np.ctypeslib.as_array(dst, shape=(count,))[:] = span[offset:offset + count]
        fv, fs = FV(fill_vertices), FS(fill_span)
        if x is None:
            self._check(self.lib.gvc_graph_upload_stream(self.h, n, len(span), C.cast(fv, C.c_void_p), C.cast(fs, C.c_void_p),
                                                         None, n_threads))
        else:
            x = _np(x, np.float32).ravel()
            if x.size != n:
                raise GvcError(f"x has {x.size} entries, the graph {n} vertices")
            self._check(self.lib.gvc_graph_upload_stream_x(self.h, n, len(span), C.cast(fv, C.c_void_p), C.cast(fs, C.c_void_p),
                                                           None, n_threads, _ptr(x, _f32p)))
        self.n_global, self.v_begin, self.v_end = n, 0, n

    def graph_staging(self, n_local: int, nnz: int):
        """Pinned host buffers of the context as numpy views (row_ptr u64, col, W, NW u32): fill
        them and hand them to graph_upload for DMA-speed uploads."""
        rp, col, W, NW = _u64p(), _u32p(), _u32p(), _u32p()
        self._check(self.lib.gvc_graph_staging(self.h, n_local, nnz, C.byref(rp), C.byref(col), C.byref(W), C.byref(NW)))
        view = lambda p, k: np.ctypeslib.as_array(p, shape=(max(k, 1),))[:k]
        return view(rp, n_local + 1), view(col, nnz), view(W, n_local), view(NW, n_local)

    def graph_adopt(self, row_ptr_i32, col_i32, W_i32, NW_i32, n_global=None, v_begin=0, v_end=None):
        """Device-resident shard: torch int32 CUDA tensors (bit patterns of uint32)."""
        nl = row_ptr_i32.numel() - 1
        if n_global is None:
            n_global = nl
        if v_end is None:
            v_end = v_begin + nl
        self._keep = [row_ptr_i32, col_i32, W_i32, NW_i32]
        if hasattr(row_ptr_i32, "device") and row_ptr_i32.device.type == "cuda":
            # the tensors may still be being written on torch's streams; libgvc reads them on its own
            import torch
            torch.cuda.synchronize(row_ptr_i32.device)
        self._check(self.lib.gvc_graph_adopt_device(self.h, n_global, v_begin, v_end, _dptr(row_ptr_i32),
                                                    _dptr(col_i32), _dptr(W_i32), _dptr(NW_i32)))
        self.n_global, self.v_begin, self.v_end = n_global, v_begin, v_end

    def graph_set_tail(self, global_vertex):
        """Where the reference's last vertex of an odd-sized graph went after a relabelling
        (None: the graph has an even vertex count, or the vertex lives on another shard)."""
        if global_vertex is not None and self.v_begin <= global_vertex < self.v_end:
            self._check(self.lib.gvc_graph_set_tail(self.h, 1, global_vertex - self.v_begin))
        else:
            self._check(self.lib.gvc_graph_set_tail(self.h, 0, 0))

    # -- peer memory (multi-GPU row exchange inside the stage kernels) -------------
    def peer_alloc(self, nbytes: int):
        """(device pointer, 64-byte IPC handle) of a fresh zeroed buffer other processes can map."""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._check(self.lib.gvc_peer_alloc(self.h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self.lib.gvc_peer_open(self.h, buf, C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int):
        self._check(self.lib.gvc_peer_close(self.h, C.c_void_p(ptr)))

    def peer_free(self, ptr: int):
        self._check(self.lib.gvc_peer_free(self.h, C.c_void_p(ptr)))

    def stage_peers(self, stage: int, ptrs):
        """Mirror the output rows of `stage` (0 or 1) into these mapped buffers of the other ranks."""
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        self._check(self.lib.gvc_stage_peers(self.h, stage, len(ptrs), arr))

    def peer_owners(self, bounds, peer_of_part):
        """Vertex ranges of the parts and, per part, the index of its owner in the stage_peers tables
        (-1: this rank): rows then only travel to the peers that own a neighbour."""
        b = np.ascontiguousarray(bounds, np.uint32)
        q = np.ascontiguousarray(peer_of_part, np.int32)
        self._check(self.lib.gvc_peer_owners(self.h, len(q), _ptr(b, _u32p), q.ctypes.data_as(_i32p)))

    # -- forward ---------------------------------------------------------------
    def forward(self, x, weight_scale: float, mode: int = MODE_EXACT) -> np.ndarray:
        """Host in, host out: gnn::model::predict.  x = None: the input that came with the graph
        (graph_upload_ranges(..., x=...))."""
        out = np.empty(self.n_global, np.float32)
        if x is None:
            self._check(self.lib.gvc_forward(self.h, None, float(weight_scale), _ptr(out, _f32p), mode))
            return out
        x = _np(x, np.float32).ravel()
        if x.size != self.n_global:
            raise GvcError(f"x has {x.size} entries, graph has {self.n_global} vertices")
        self._check(self.lib.gvc_forward(self.h, _ptr(x, _f32p), float(weight_scale), _ptr(out, _f32p), mode))
        return out

    def forward_keys(self, x, weight_scale: float, mode: int = MODE_EXACT):
        """(scores, keys = min(s, 1 - s), side = s > 0.5): the inputs of the caller's selection order."""
        x = _np(x, np.float32).ravel()
        if x.size != self.n_global:
            raise GvcError(f"x has {x.size} entries, graph has {self.n_global} vertices")
        out = np.empty(self.n_global, np.float32)
        keys = np.empty(self.n_global, np.float32)
        side = np.empty(self.n_global, np.uint8)
        self._check(self.lib.gvc_forward_keys(self.h, _ptr(x, _f32p), float(weight_scale), _ptr(out, _f32p), _ptr(keys, _f32p),
                                              side.ctypes.data_as(C.POINTER(C.c_ubyte)), mode))
        return out, keys, side

    def row_order(self) -> np.ndarray:
        """vertex_of_row: which vertex every row of h1 / h2 belongs to (identity unless renumbered, gvc_debug_row_order)"""
        out = np.empty(self.n_global, np.uint32)
        if self.lib.gvc_debug_row_order(self.h, _ptr(out, _u32p)) < 0:
            raise GvcError("gvc_debug_row_order failed")
        return out

    def forward_device(self, d_x, weight_scale: float, d_scores, mode: int = MODE_EXACT):
        self._check(self.lib.gvc_forward_device(self.h, _dptr(d_x), float(weight_scale), _dptr(d_scores), mode))

    def stage_device(self, stage: int, d_in, d_out, weight_scale: float, mode: int = MODE_EXACT):
        self._check(self.lib.gvc_stage_device(self.h, stage, _dptr(d_in), _dptr(d_out), float(weight_scale), mode))

    def graph_layer_device(self, d_in, width: int, d_out, weight_scale: float):
        self._check(self.lib.gvc_graph_layer_device(self.h, _dptr(d_in), width, _dptr(d_out), float(weight_scale)))

    def linear_layer_device(self, layer_index: int, n: int, d_in, d_out):
        self._check(self.lib.gvc_linear_layer_device(self.h, layer_index, n, _dptr(d_in), _dptr(d_out)))

    def relu_device(self, count: int, d_in, d_out):
        self._check(self.lib.gvc_relu_device(self.h, count, _dptr(d_in), _dptr(d_out)))

    def sigmoid_device(self, count: int, d_in, d_out, mode: int = MODE_EXACT):
        self._check(self.lib.gvc_sigmoid_device(self.h, count, _dptr(d_in), _dptr(d_out), mode))

    # -- host-buffer single layers (what the drop-in layer structs call) -----------
    def graph_layer_host(self, x, weight_scale: float) -> np.ndarray:
        x = _np(x, np.float32)
        n, w = x.shape
        out = np.empty((n, 2 * w + 3), np.float32)
        self._check(self.lib.gvc_graph_layer_host(self.h, _ptr(x, _f32p), w, _ptr(out, _f32p), float(weight_scale)))
        return out

    def linear_host(self, x, W, bias, mode: int = MODE_EXACT) -> np.ndarray:
        x, W, bias = _np(x, np.float32), _np(W, np.float32), _np(bias, np.float32).ravel()
        n, K = x.shape
        out = np.empty((n, W.shape[1]), np.float32)
        self._check(self.lib.gvc_linear_host(self.h, n, K, W.shape[1], _ptr(x, _f32p), _ptr(W, _f32p),
                                             _ptr(bias, _f32p), _ptr(out, _f32p), mode))
        return out

    def relu_host(self, x) -> np.ndarray:
        x = _np(x, np.float32)
        out = np.empty_like(x)
        self._check(self.lib.gvc_relu_host(self.h, x.size, _ptr(x, _f32p), _ptr(out, _f32p)))
        return out

    def sigmoid_host(self, x, mode: int = MODE_EXACT) -> np.ndarray:
        x = _np(x, np.float32)
        out = np.empty_like(x)
        self._check(self.lib.gvc_sigmoid_host(self.h, x.size, _ptr(x, _f32p), _ptr(out, _f32p), mode))
        return out

    def sgemm_host(self, A, B, C0=None, trans_a=False, trans_b=False, beta: float = 0.0) -> np.ndarray:
        A, B = _np(A, np.float32), _np(B, np.float32)
        m = A.shape[1] if trans_a else A.shape[0]
        k = A.shape[0] if trans_a else A.shape[1]
        n = B.shape[0] if trans_b else B.shape[1]
        out = np.zeros((m, n), np.float32) if C0 is None else _np(C0, np.float32).copy()
        self._check(self.lib.gvc_sgemm_host(self.h, int(trans_a), int(trans_b), m, n, k, _ptr(A, _f32p), A.shape[1],
                                            _ptr(B, _f32p), B.shape[1], float(beta), _ptr(out, _f32p), n))
        return out

    # -- plumbing ----------------------------------------------------------------
    @property
    def stream_ptr(self) -> int:
        return int(self.lib.gvc_stream(self.h) or 0)

    def use_torch_stream(self, stream=None):
        """Enqueue on a torch stream (default: a new one) and return it; torch copies, NCCL
        collectives and events issued under ``torch.cuda.stream(s)`` then order with the forward."""
        import torch
        if stream is None:
            stream = torch.cuda.Stream(device=torch.device("cuda", self.device))
        self._check(self.lib.gvc_set_stream(self.h, C.c_void_p(stream.cuda_stream)))
        self._torch_stream = stream            # keep it alive while it is set
        return stream

    def torch_stream(self):
        return self.use_torch_stream(getattr(self, "_torch_stream", None))

    def sync(self):
        self._check(self.lib.gvc_sync(self.h))

    def px_stats(self):
        """{hubs, chunks, slow batches} of the parallel exact hub sums in the last stage launched."""
        out = np.zeros(8, np.uint32)
        self._check(self.lib.gvc_debug_px(self.h, _ptr(out, _u32p)))
        return {"hubs": int(out[0]), "chunks": int(out[1]), "slow_batches": int(out[3]),
                "walk_wait_kcycles": int(out[5]), "walk_kcycles": int(out[6]), "quantiser_dirty": int(out[7])}

    @property
    def launches(self) -> int:
        return int(self.lib.gvc_launch_count(self.h))


class Group:
    """gvc_group: one context per device in THIS process, the forward sharded over them."""

    def __init__(self, devices):
        self.lib = load_library()
        h = C.c_void_p()
        d = (C.c_int * len(devices))(*devices)
        self._check(self.lib.gvc_group_create(C.byref(h), d, len(devices)))
        self.h = h
        self.n = 0

    def _check(self, rc):
        if rc != 0:
            raise GvcError(f"libgvc error {rc}: {self.lib.gvc_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.gvc_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def model_upload(self, layers):
        n = len(layers)
        kinds = (C.c_int * n)(*[int(k) for k, _, _ in layers])
        rows, cols = (C.c_int * n)(), (C.c_int * n)()
        Wp, bp = (_f32p * n)(), (_f32p * n)()
        keep = []
        for i, (k, W, b) in enumerate(layers):
            if k == LINEAR:
                W = _np(W, np.float32)
                b = _np(b, np.float32).ravel()
                rows[i], cols[i] = W.shape
                Wp[i], bp[i] = W.ctypes.data_as(_f32p), b.ctypes.data_as(_f32p)
                keep += [W, b]
        self._check(self.lib.gvc_group_model_upload(self.h, n, kinds, rows, cols, Wp, bp))

    def graph_upload(self, row_ptr, col, W, NW):
        row_ptr, col = _np(row_ptr, np.uint64), _np(col, np.uint32)
        W, NW = _np(W, np.uint32), _np(NW, np.uint32)
        self.n = len(row_ptr) - 1
        self._check(self.lib.gvc_group_graph_upload(self.h, self.n, _ptr(row_ptr, _u64p), _ptr(col, _u32p), _ptr(W, _u32p), _ptr(NW, _u32p)))

    @property
    def bounds(self):
        b = np.zeros(self.lib.gvc_group_size(self.h) + 1, np.uint32)
        self._check(self.lib.gvc_group_bounds(self.h, _ptr(b, _u32p)))
        return b.tolist()

    def forward(self, x, weight_scale: float, mode: int = MODE_EXACT) -> np.ndarray:
        x = _np(x, np.float32).ravel()
        if x.size != self.n:
            raise GvcError(f"x has {x.size} entries, graph has {self.n} vertices")
        out = np.empty(self.n, np.float32)
        self._check(self.lib.gvc_group_forward(self.h, _ptr(x, _f32p), float(weight_scale), _ptr(out, _f32p), mode))
        return out


class Trainer:
    """gvc_trainer: a model_training on the device (reference old_files/src/lib/gnn_training.cpp).  The graph is
    the context's current graph."""

    def __init__(self, ctx: "Context", layers):
        self.ctx, self.lib, self.layers = ctx, ctx.lib, layers
        n = len(layers)
        kinds = (C.c_int * n)(*[int(k) for k, _, _ in layers])
        rows, cols = (C.c_int * n)(), (C.c_int * n)()
        Wp, bp = (_f32p * n)(), (_f32p * n)()
        keep = []
        for i, (k, W, b) in enumerate(layers):
            if k == LINEAR:
                W = _np(W, np.float32)
                b = _np(b, np.float32).ravel()
                rows[i], cols[i] = W.shape
                Wp[i], bp[i] = W.ctypes.data_as(_f32p), b.ctypes.data_as(_f32p)
                keep += [W, b]
        h = C.c_void_p()
        ctx._check(self.lib.gvc_trainer_create(ctx.h, n, kinds, rows, cols, Wp, bp, C.byref(h)))
        self.h = h
        self.in_w, self.out_w = self.lib.gvc_trainer_input_width(h), self.lib.gvc_trainer_output_width(h)

    def close(self):
        if self.h:
            self.lib.gvc_trainer_destroy(self.h)
            self.h = None

    def predict(self, x, scales, mode: int = MODE_EXACT, want_out: bool = True):
        n = self.ctx.n_global
        x = _np(x, np.float32)
        if x.size != n * self.in_w:
            raise GvcError(f"x has {x.size} values, the model takes {n} x {self.in_w}")
        x = x.reshape(n, self.in_w)
        sc = _np(np.atleast_1d(scales), np.float32)
        out = np.empty((n, self.out_w), np.float32) if want_out else None
        self.ctx._check(self.lib.gvc_trainer_predict(self.h, _ptr(x, _f32p), _ptr(sc, _f32p), sc.size,
                                                     _ptr(out, _f32p) if want_out else None, mode))
        return out

    def backprop(self, grad, mode: int = MODE_EXACT, want_grad_x: bool = True):
        n = self.ctx.n_global
        grad = _np(grad, np.float32).reshape(n, self.out_w)
        gx = np.empty((n, self.in_w), np.float32) if want_grad_x else None
        self.ctx._check(self.lib.gvc_trainer_backprop(self.h, _ptr(grad, _f32p), _ptr(gx, _f32p) if want_grad_x else None, mode))
        return gx

    def mse_backprop(self, y, mode: int = MODE_EXACT) -> float:
        y = _np(y, np.float32).reshape(self.ctx.n_global, self.out_w)
        loss = C.c_float()
        self.ctx._check(self.lib.gvc_trainer_mse_backprop(self.h, _ptr(y, _f32p), C.cast(C.byref(loss), _f32p), mode))
        return float(loss.value)

    def sgd_step(self, batch_size: int, lr=0.1, momentum=0.9, weight_decay=0.0):
        self.ctx._check(self.lib.gvc_trainer_sgd_step(self.h, batch_size, lr, momentum, weight_decay))

    def zero_grad(self):
        self.ctx._check(self.lib.gvc_trainer_zero_grad(self.h))

    def read(self, what: int, layer: int):
        """(W-shaped, bias-shaped) arrays: what = 0 parameters, 1 gradients, 2 velocities."""
        W = np.empty(np.asarray(self.layers[layer][1]).shape, np.float32)
        b = np.empty(W.shape[1], np.float32)
        self.ctx._check(self.lib.gvc_trainer_read(self.h, what, layer, _ptr(W, _f32p), _ptr(b, _f32p)))
        return W, b

    def write(self, what: int, layer: int, W, b):
        W, b = _np(W, np.float32), _np(b, np.float32).ravel()
        self.ctx._check(self.lib.gvc_trainer_write(self.h, what, layer, _ptr(W, _f32p), _ptr(b, _f32p)))


def parse_metis(path, n_threads: int = 0, csr: bool = False):
    """gvc_metis_parse: (n, weights u32, eu u32, ev u32[, (row_ptr u64, col u32, nw u32)]) of a METIS file as
    the reference's parse_graph reads it (src/GNN_VC.cpp:34-91).  Host code, needs no GPU."""
    lib = load_library()
    h = C.c_void_p()
    rc = lib.gvc_metis_parse(str(path).encode(), n_threads, C.byref(h))
    if rc != 0:
        raise GvcError(f"libgvc error {rc}: {lib.gvc_last_error().decode()}")
    try:
        n, e = int(lib.gvc_metis_vertices(h)), int(lib.gvc_metis_edges(h))
        w = np.ctypeslib.as_array(lib.gvc_metis_weights(h), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        eu = np.ctypeslib.as_array(lib.gvc_metis_edge_u(h), shape=(e,)).copy() if e else np.zeros(0, np.uint32)
        ev = np.ctypeslib.as_array(lib.gvc_metis_edge_v(h), shape=(e,)).copy() if e else np.zeros(0, np.uint32)
        if not csr:
            return n, w, eu, ev
        rp, col, nw = np.zeros(n + 1, np.uint64), np.zeros(max(2 * e, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
        rc = lib.gvc_metis_csr(h, _ptr(rp, _u64p), _ptr(col, _u32p), _ptr(nw, _u32p))
        if rc != 0:
            raise GvcError(f"libgvc error {rc}: {lib.gvc_last_error().decode()}")
        return n, w, eu, ev, (rp, col[:2 * e], nw[:n])
    finally:
        lib.gvc_metis_free(h)


def load_model_npz(path):
    """tests/golden/mwvc_model.npz -> [(kind, W, bias)]"""
    z = np.load(path)
    kinds = z["kinds"]
    layers = []
    for i, k in enumerate(kinds):
        if int(k) == LINEAR:
            layers.append((LINEAR, z[f"W{i}"], z[f"b{i}"]))
        else:
            layers.append((int(k), None, None))
    return layers


def random_model(seed: int = 0):
    """The GNN_VC architecture with random weights,
    uniform(+-1/sqrt(K+1)) like linear_layer's ctor (reference src/gnn_inference.cpp:7-18)."""
    rng = np.random.default_rng(seed)
    dims = [(5, 32), (32, 32), (32, 16), (35, 32), (32, 32), (32, 16), (35, 32), (32, 16), (16, 1)]
    it = iter(dims)
    layers = []
    pattern = [GRAPH, LINEAR, RELU, LINEAR, RELU, LINEAR, RELU,
               GRAPH, LINEAR, RELU, LINEAR, RELU, LINEAR, RELU,
               GRAPH, LINEAR, RELU, LINEAR, RELU, LINEAR, SIGMOID]
    for k in pattern:
        if k == LINEAR:
            K, N = next(it)
            lim = 1.0 / np.sqrt(K + 1)
            layers.append((LINEAR, rng.uniform(-lim, lim, (K, N)).astype(np.float32),
                           rng.uniform(-lim, lim, N).astype(np.float32)))
        else:
            layers.append((k, None, None))
    return layers
