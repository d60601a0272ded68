/*
 * gvc.h -- C ABI of libgvc.so, the B200 (sm_100a) GNN forward for GNN_VC.
 *
 * This is the drop-in boundary for the hot path gnn::model::predict of the
 * reference (KennethLangedal/GNN-MWVC).  Plain pointers and sizes only; no C++,
 * torch or CUDA types.  The reference-side binding (a replacement
 * src/gnn_inference.cpp that keeps include/gnn_inference.hpp untouched) lives in
 * gnn-mwvc_b200/host/ and is described in INTEGRATION.md.
 *
 * Every entry point names the reference interface it stands in for
 * (file:line under the reference tree).
 *
 * Conventions: all functions return 0 on success or a non-zero code and leave a
 * message retrievable with gvc_last_error(); none of them throws or aborts.
 * One context belongs to one CUDA device and one caller thread (the reference's
 * predict is not re-entrant either: include/gnn_inference.hpp:44 `mutable`).
 * There is no CPU fallback: without a usable CUDA device gvc_ctx_create fails.
 */
#ifndef GVC_H
#define GVC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gvc_ctx gvc_ctx;

/* Layer kinds, in the order of gnn::component's variant alternatives
 * (include/gnn_inference.hpp:38). */
enum { GVC_LINEAR = 0, GVC_GRAPH = 1, GVC_RELU = 2, GVC_SIGMOID = 3 };

/* Arithmetic modes of the forward.
 *  GVC_MODE_EXACT reproduces the reference's fp32 operation order (OpenBLAS
 *    0.3.15 "Prescott" sgemm run single-threaded, sequential neighbour sums,
 *    glibc expf): scores are bit-identical to the reference.
 *  GVC_MODE_FAST uses fused multiply-adds and the device expf: scores agree to
 *    1e-4 relative (typically ~1e-6), not bit for bit. */
enum { GVC_MODE_EXACT = 0, GVC_MODE_FAST = 1 };

/* Error codes (besides CUDA's own, passed through as 1000 + cudaError_t). */
enum {
    GVC_OK = 0,
    GVC_ERR_ARG = 1,        /* bad argument */
    GVC_ERR_STATE = 2,      /* model or graph not uploaded yet */
    GVC_ERR_NO_DEVICE = 3,  /* no CUDA device / wrong architecture */
    GVC_ERR_ALLOC = 4,
    GVC_ERR_UNSUPPORTED = 5
};

const char *gvc_last_error(void);

/* Version of this ABI (bumped when entry points are added or change meaning; 4 = round 2). */
int gvc_abi_version(void);

/* ---- context ------------------------------------------------------------ */

/* One context per device; stands in for the gnn::model object's life time
 * (src/GNN_VC.cpp:250).  `device` is a CUDA ordinal. */
int gvc_ctx_create(gvc_ctx **out, int device);
void gvc_ctx_destroy(gvc_ctx *ctx);

/* ---- model: gnn::operator>> (src/gnn_inference.cpp:120-139) --------------- */

/* Upload a parsed model: n_layers records; kinds[i] in GVC_*; for GVC_LINEAR
 * rows[i] x cols[i] row-major weights W[i] and cols[i] bias values bias[i]
 * (host pointers, copied; ignored for other kinds).  Any layer sequence the
 * reference parser accepts is accepted.  The 21-layer GNN_VC architecture
 *  is detected structurally and runs on the fused kernels; every
 * other sequence runs on per-layer kernels. */
int gvc_model_upload(gvc_ctx *ctx, int n_layers, const int *kinds, const int *rows,
                     const int *cols, const float *const *W, const float *const *bias);

/* 1 if the uploaded model runs on the fused three-kernel path, else 0. */
int gvc_model_is_fused(const gvc_ctx *ctx);

/* graph_layer::WEIGHT_SCALE is a member of EVERY graph layer (include/gnn_inference.hpp:25, read at
 * src/gnn_inference.cpp:39-40); model::set_weight_scale (:83-90) makes them equal, but a model built
 * with add_layer may carry different ones.  scales[i] is the scale of the i-th graph layer of the
 * uploaded model (n_graph_layers must match it).  While set, these take precedence over the
 * weight_scale argument of the forward calls; n_graph_layers = 0 forgets them (and so does a model
 * upload). */
int gvc_model_weight_scales(gvc_ctx *ctx, int n_graph_layers, const float *scales);

/* ---- graph: what predict reads through reduction_graph's accessors ---------
 * size() include/reduction_graph.hpp:141, begin(u)/end(u) :692-704, D :144,
 * W :147-151, NW :153-158, as used at src/gnn_inference.cpp:32-40.
 *
 * CSR with n vertices: row_ptr[n+1] (row_ptr[0] == 0), col[row_ptr[n]]
 * 0-indexed neighbour ids in the order begin(u)..end(u) yields them, W[n] and
 * NW[n] the integer vertex and neighbourhood weights.  Host pointers, copied.
 * Checked on the device during the upload: row_ptr monotone, every neighbour id
 * < n (GVC_ERR_ARG otherwise, and the context is left without a graph). */
int gvc_graph_upload(gvc_ctx *ctx, uint32_t n, const uint64_t *row_ptr, const uint32_t *col,
                     const uint32_t *W, const uint32_t *NW);

/* Vertex-range shard for multi-GPU runs (one process per GPU): this context
 * owns vertices [v_begin, v_end) of an n_global-vertex graph.  row_ptr has
 * v_end - v_begin + 1 entries starting at 0; col holds GLOBAL neighbour ids;
 * W/NW are the shard's slices.  Host pointers, copied. */
int gvc_graph_upload_shard(gvc_ctx *ctx, uint32_t n_global, uint32_t v_begin, uint32_t v_end,
                           const uint64_t *row_ptr, const uint32_t *col, const uint32_t *W,
                           const uint32_t *NW);

/* Pinned host buffers owned by the context, big enough for a shard of n_local vertices and nnz
 * adjacency entries: row_ptr[n_local + 1], col[nnz], W[n_local], NW[n_local].  A caller that
 * builds its CSR anyway (the drop-in's extraction from the reduction_graph,
 * gnn-mwvc_b200/host/gvc_gnn_inference.cpp) writes it here and passes these pointers to
 * gvc_graph_upload[_shard]: the copies then run as DMA straight from the buffers, the adjacency
 * on a second stream beside the schedule construction.  Any other host pointers stay valid
 * arguments of the upload calls, just slower.  The buffers are valid until the next
 * gvc_graph_staging call with larger sizes or gvc_ctx_destroy. */
int gvc_graph_staging(gvc_ctx *ctx, uint32_t n_local, uint64_t nnz, uint64_t **row_ptr,
                      uint32_t **col, uint32_t **W, uint32_t **NW);

/* The graph as the reduction_graph holds it (include/reduction_graph.hpp:29-33): ONE edge array with
 * holes -- removed neighbours are rotated to the front of a list and skipped, folds append new lists
 * at the end (:248-510) -- and a half-open range [begin, end) into it per vertex, which is exactly
 * what begin(u)/end(u) (:692-704) expose.  Instead of compacting that into a CSR on the host and
 * sending it (gvc_graph_upload), the caller describes it with two callbacks and libgvc streams it:
 * worker threads (n_threads, 0 = pick) call
 *     fill_vertices(user, first, count, begin, end, W, NW)   offsets into the span, W(u), NW(u)
 *     fill_span(user, offset, count, dst)                    copy span[offset, offset + count) to dst
 * for disjoint pieces, concurrently, writing straight into a small ring of pinned buffers whose
 * earlier slots are in flight to the device meanwhile; degrees, offsets, the packed adjacency
 * (lists in the order the span holds them), the checks of gvc_graph_upload and the degree schedule are
 * computed on the device.  The span is [0, span_len): the caller rebases begin/end to the smallest
 * begin.  Whole-graph contexts only (v_begin = 0, v_end = n).  This is SURVEY.md 8(f) item 2: the
 * host neither compacts nor copies the adjacency twice, and nothing is pinned per graph. */
typedef void (*gvc_fill_vertices_fn)(void *user, uint32_t first, uint32_t count, uint32_t *begin,
                                     uint32_t *end, uint32_t *W, uint32_t *NW);
typedef void (*gvc_fill_span_fn)(void *user, uint64_t offset, uint64_t count, uint32_t *dst);
int gvc_graph_upload_stream(gvc_ctx *ctx, uint32_t n, uint64_t span_len, gvc_fill_vertices_fn fill_vertices,
                            gvc_fill_span_fn fill_span, void *user, int n_threads);
/* The same, and the forward's input travels along: x[0..n) (the `in` of model::predict,
 * src/GNN_VC.cpp:189-191) is copied by the worker threads next to W and NW, so the single-threaded
 * staging copy of gvc_forward drops out of predict's critical path.  A following
 * gvc_forward(ctx, NULL, ...) uses it (any number of times, until the next graph upload or a forward
 * with an explicit x).  x == NULL: exactly gvc_graph_upload_stream. */
int gvc_graph_upload_stream_x(gvc_ctx *ctx, uint32_t n, uint64_t span_len, gvc_fill_vertices_fn fill_vertices,
                              gvc_fill_span_fn fill_span, void *user, int n_threads, const float *x);

/* Same shard description with DEVICE pointers that stay owned by the caller and
 * must outlive the context's use of them (no copy; row_ptr is uint32 here, the
 * layout the kernels read).  Used by the benchmark, which builds its synthetic
 * graphs on the GPU. */
int gvc_graph_adopt_device(gvc_ctx *ctx, uint32_t n_global, uint32_t v_begin, uint32_t v_end,
                           const uint32_t *d_row_ptr, const uint32_t *d_col, const uint32_t *d_W,
                           const uint32_t *d_NW);

/* Exact mode reproduces OpenBLAS' 1-row remainder kernel for the LAST vertex of a graph with an
 * odd vertex count (oracle/gnn_oracle.c).  By default that is vertex n_global - 1 and it is
 * handled by the shard that owns it.  A caller that renames vertices before sharding
 * (gnn-mwvc_b200/graphs.py balanced_relabel) says where that vertex went: has_tail = 0 -> this
 * shard owns no such vertex, 1 -> it is local vertex `local_index`.  Call after the graph
 * upload/adopt (which resets to the default). */
int gvc_graph_set_tail(gvc_ctx *ctx, int has_tail, uint32_t local_index);

/* Optional: pay the one-off costs of the first streamed upload NOW (pinning the ring of upload slots,
 * the first device allocation of about graph_bytes_hint bytes; 0 = ring only) instead of inside the first
 * predict.  The drop-in calls it from a helper thread while the solver still parses and reduces the graph. */
int gvc_ctx_warm(gvc_ctx *ctx, uint64_t graph_bytes_hint);

/* ---- forward: gnn::model::predict (src/gnn_inference.cpp:67-81) ------------ */

/* Whole forward with HOST buffers: x[n] (= in(u,0), src/GNN_VC.cpp:189-191),
 * weight_scale (= graph_layer::WEIGHT_SCALE, include/gnn_inference.hpp:25, set by
 * set_weight_scale src/gnn_inference.cpp:83-90), scores[n] (= out(u,0)).  Includes
 * the host->device copy of x and the device->host copy of the scores.  n == 0
 * is a no-op (the reference is called with an empty graph).
 * x == NULL: use the input that came with the graph (gvc_graph_upload_stream_x); an error if none did.
 * Single-shard contexts only. */
int gvc_forward(gvc_ctx *ctx, const float *x, float weight_scale, float *scores, int mode);

/* Same with DEVICE buffers (d_x[n_global], d_scores[v_end - v_begin]), enqueued
 * on the context's stream without synchronising.  Single-shard contexts only. */
int gvc_forward_device(gvc_ctx *ctx, const float *d_x, float weight_scale, float *d_scores,
                       int mode);

/* The forward plus what the caller's vertex selection is computed from.  After
 * predict, src/GNN_VC.cpp:194-206 orders the vertices by min(out, 1 - out) with an eps-tolerance
 * comparator that also looks at out > 0.5; the stage-2 kernel writes both beside the scores:
 *   keys[u] = std::min(out(u,0), 1.0f - out(u,0))    side[u] = out(u,0) > 0.5f
 * evaluated in fp32 exactly as the comparator evaluates them, so a sort fed with them gives the
 * reference's order (host/gvc_dropin_capi.cpp gvcd_predict_order).  GNN_VC architecture only.
 * Host buffers of n entries each / device buffers of v_end - v_begin entries each. */
int gvc_forward_keys(gvc_ctx *ctx, const float *x, float weight_scale, float *scores, float *keys,
                     unsigned char *side, int mode);
int gvc_forward_device_keys(gvc_ctx *ctx, const float *d_x, float weight_scale, float *d_scores,
                            float *d_keys, unsigned char *d_side, int mode);
/* gvc_forward (what model::predict calls) leaves the keys of its forward on the device as a
 * by-product; this fetches them (n entries each) without running anything again. */
int gvc_last_keys(gvc_ctx *ctx, float *keys, unsigned char *side);

/* One fused stage of the GNN_VC architecture on this context's shard, DEVICE
 * buffers, enqueued on the context's stream:
 *   stage 0: graph layer (w=1)  + 5->32->32->16   d_in = x  [n_global]      -> d_out = h1 [n_global x 16]
 *   stage 1: graph layer (w=16) + 35->32->32->16  d_in = h1 [n_global x 16] -> d_out = h2 [n_global x 16]
 *   stage 2: graph layer (w=16) + 35->32->16->1 + sigmoid
 *                                                 d_in = h2 [n_global x 16] -> d_out = scores [v_end - v_begin]
 * Stages 0 and 1 write rows [v_begin, v_end) of the FULL d_out buffer; the
 * caller exchanges the other rows between shards before the next stage
 * (gnn-mwvc_b200/dist.py).  Reference: graph_layer::forward :27-42,
 * linear_layer::forward :20-25, ReLU :44-47, sigmoid :49-52. */
int gvc_stage_device(gvc_ctx *ctx, int stage, const float *d_in, float *d_out,
                     float weight_scale, int mode);

/* ---- multi-GPU: row exchange through peer memory --------------------------
 * One process per GPU.  Instead of all-gathering the 16-float rows between two
 * stages (gnn-mwvc_b200/dist.py exchange_rows), the store epilogue of stages 0
 * and 1 can write every row that another shard may read straight into the other
 * ranks' copies of the output buffer over NVLink; between stages the ranks then
 * only need a barrier.  The buffers must be reachable from this device:
 * gvc_peer_alloc allocates one and returns its CUDA IPC handle (64 bytes, to be
 * sent to the other processes by any means), gvc_peer_open maps a buffer another
 * process allocated, gvc_stage_peers(stage, n, ptrs) registers the n <= 7 mapped
 * buffers that mirror d_out of that stage (n = 0 switches the mirroring off).
 * Rows of isolated vertices are not mirrored (nobody reads them). */
#define GVC_PEER_HANDLE_BYTES 64
int gvc_peer_alloc(gvc_ctx *ctx, uint64_t bytes, void **d_ptr, unsigned char *handle);
int gvc_peer_open(gvc_ctx *ctx, const unsigned char *handle, void **d_ptr);
int gvc_peer_close(gvc_ctx *ctx, void *d_ptr);
int gvc_peer_free(gvc_ctx *ctx, void *d_ptr);
int gvc_stage_peers(gvc_ctx *ctx, int stage, int n_peers, float *const *d_out_peers);
/* Optional: tell the context which rank owns which vertex range -- bounds[0..n_parts] ascending,
 * bounds[n_parts] == n_global; peer_of_part[k] = position of part k's owner in the tables given to
 * gvc_stage_peers, -1 for this rank's own part.  A row then only goes to the peers that own a
 * neighbour of its vertex (the adjacency is symmetric: nobody else reads it) instead of to all of
 * them.  The per-vertex peer lists are rebuilt with every graph upload; n_parts = 0 forgets them. */
int gvc_peer_owners(gvc_ctx *ctx, int n_parts, const uint32_t *bounds, const int *peer_of_part);

/* ---- several GPUs in ONE process: predict on a graph too large (or too slow) for one device -------
 * SURVEY.md 8(e) behind one call, for a caller like gnn::model::predict (include/gnn_inference.hpp:50)
 * that is a single thread of a single process: a group owns one context per device; the graph is cut
 * into contiguous vertex ranges of about equal work; a forward runs the three stage kernels on every
 * device, each storing the rows other devices read straight into their buffers (peer access over
 * NVLink), with "every device has finished its stage" between two stages expressed as CUDA events
 * the streams wait on.  Scores come back in vertex order, bit-identical to a one-device forward
 * (per-vertex arithmetic does not depend on the sharding).  GNN_VC architecture only; devices may
 * repeat (two shards on one GPU: how the path is tested where only one GPU is visible). */
typedef struct gvc_group gvc_group;
int gvc_group_create(gvc_group **out, const int *devices, int n_devices);
void gvc_group_destroy(gvc_group *group);
int gvc_group_size(const gvc_group *group);
int gvc_group_model_upload(gvc_group *group, int n_layers, const int *kinds, const int *rows,
                           const int *cols, const float *const *W, const float *const *bias);
int gvc_group_model_weight_scales(gvc_group *group, int n_graph_layers, const float *scales);
/* whole graph in (as gvc_graph_upload), shards out */
int gvc_group_graph_upload(gvc_group *group, uint32_t n, const uint64_t *row_ptr, const uint32_t *col,
                           const uint32_t *W, const uint32_t *NW);
/* bounds_out[0 .. size]: the vertex ranges the last upload chose */
int gvc_group_bounds(const gvc_group *group, uint32_t *bounds_out);
int gvc_group_forward(gvc_group *group, const float *x, float weight_scale, float *scores, int mode);

/* Single layers on device buffers, row counts explicit (generic path; also the
 * kernel-level parity tests).  in/out are row-major n x width. */
int gvc_graph_layer_device(gvc_ctx *ctx, const float *d_in, int width, float *d_out,
                           float weight_scale);                       /* :27-42  */
int gvc_linear_layer_device(gvc_ctx *ctx, int layer_index, uint64_t n, const float *d_in,
                            float *d_out);                            /* :20-25  */
int gvc_relu_device(gvc_ctx *ctx, uint64_t count, const float *d_in, float *d_out);      /* :44-47 */
int gvc_sigmoid_device(gvc_ctx *ctx, uint64_t count, const float *d_in, float *d_out, int mode); /* :49-52 */

/* Host-buffer forms of the single layers and of dot(): what the public
 * forward() methods of the layer structs (include/gnn_inference.hpp:11-36) and
 * dot (include/matrix.hpp:49, src/matrix.cpp:106-122) call in the drop-in host
 * code.  Copies in, runs the kernel, copies out, synchronises. */
int gvc_graph_layer_host(gvc_ctx *ctx, const float *in, int width, float *out, float weight_scale);
int gvc_linear_host(gvc_ctx *ctx, uint64_t n, int K, int Nout, const float *in, const float *W,
                    const float *bias, float *out, int mode);
int gvc_relu_host(gvc_ctx *ctx, uint64_t count, const float *in, float *out);
int gvc_sigmoid_host(gvc_ctx *ctx, uint64_t count, const float *in, float *out, int mode);
/* Row-major C[m x n] = op(A) * op(B) + beta * C with op = transpose when the flag
 * is non-zero; lda/ldb/ldc are the row lengths of the stored matrices. */
int gvc_sgemm_host(gvc_ctx *ctx, int trans_a, int trans_b, uint64_t m, uint64_t n, uint64_t k,
                   const float *A, uint64_t lda, const float *B, uint64_t ldb, float beta, float *C,
                   uint64_t ldc);

/* ---- the METIS-format reader: parse_graph (src/GNN_VC.cpp:34-91; format README.md:45-60) ------------
 * SURVEY.md 8(f) item 3.  The reference reads its input with one getline + stringstream per vertex;
 * this reader maps the file and parses it with n_threads threads (0 = all cores), with the same
 * result: the vertex weights and the sorted, de-duplicated undirected edges (u < v) that parse_graph
 * hands to the reduction_graph constructor -- including its quirks (only neighbours with a larger id
 * are kept; a header E larger than the file's edge count leaves one (0,0) self-loop; see
 * host/gvc_metis.cpp).  Where the reference would run into undefined behaviour (more edges than the
 * header says, ids out of range) an error is returned instead.  gvc_metis_csr builds the adjacency
 * exactly as the reduction_graph constructor does (include/reduction_graph.hpp:103-128), ready for
 * gvc_graph_upload.  Pure host code. */
typedef struct gvc_metis gvc_metis;
int gvc_metis_parse(const char *path, int n_threads, gvc_metis **out);
void gvc_metis_free(gvc_metis *m);
uint64_t gvc_metis_vertices(const gvc_metis *m);
uint64_t gvc_metis_edges(const gvc_metis *m);
const uint32_t *gvc_metis_weights(const gvc_metis *m);
const uint32_t *gvc_metis_edge_u(const gvc_metis *m);
const uint32_t *gvc_metis_edge_v(const gvc_metis *m);
int gvc_metis_csr(const gvc_metis *m, uint64_t *row_ptr, uint32_t *col, uint32_t *nw);

/* ---- plumbing ----------------------------------------------------------- */

/* The context's CUDA stream as an opaque handle (cudaStream_t). */
void *gvc_stream(gvc_ctx *ctx);
/* Make the context enqueue on a caller-owned stream (cudaStream_t of the context's device),
 * e.g. the host framework's current stream, so that its copies, collectives and events order
 * with the forward without extra synchronisation.  NULL returns to the context's own stream.
 * The caller keeps the stream alive while it is set. */
int gvc_set_stream(gvc_ctx *ctx, void *stream);
/* Block until everything enqueued on the context's stream has finished. */
int gvc_sync(gvc_ctx *ctx);
/* Number of kernels this context has launched so far. */
uint64_t gvc_launch_count(const gvc_ctx *ctx);
/* Statistics of the exact-mode parallel hub sums of the LAST stage launched (csrc/gvc_px.cuh):
 * out8 = {vertices handled that way, their 4096-entry chunks, -, batches that fell back to the
 * element-wise chain, -, SM cycles / 1024 the walks waited for their chunks, SM cycles / 1024 the walks
 * took, batches the quantiser marked}.  For tests and tools. */
int gvc_debug_px(gvc_ctx *ctx, uint32_t *out8);
/* Device buffers of the last forward (h1/h2: n_global x 16), for tests. */
/* Rows of the stage outputs h1 / h2 (gvc_stage_device with the caller's buffers, gvc_debug_h): row r belongs to
 * vertex vertex_of_row[r].  The identity, except for whole-graph contexts of GVC_ROW_ORDER_MIN_VERTICES (default
 * 500 000) vertices and more, which keep their rows in the order of the degree schedule so that the rows of
 * the high-degree vertices share cache lines; x, scores and selection keys are always in
 * the caller's numbering.  Returns 1 if renumbered, 0 if not, < 0 on error. */
int gvc_debug_row_order(gvc_ctx *ctx, uint32_t *vertex_of_row);
const float *gvc_debug_h(const gvc_ctx *ctx, int which);

/* ---- training path ---------------------------------------------------------
 * Reference: old_files/src/lib/gnn_training.cpp (+ old_files/include/gnn/gnn_training.hpp): the
 * backward pass of the four layer kinds (:17-65), model_training::predict / backprop (:81-129),
 * MSE_loss / MSE_grad (:175-190), SGD_step (:192-224), zero_grad (:226-235).
 *
 * gvc_trainer is a model_training on the device: parameters, gradients and velocities of every linear
 * layer and, between a predict and its backprop, every layer's input (the reference's in_copy).  The
 * graph is the context's current graph (any upload path; whole-graph contexts).  GVC_MODE_EXACT: the
 * reference's fp32 operation order everywhere, the three dot() calls of a linear layer in the order of
 * the OpenBLAS kernel the parity tests pin (bit-identical gradients; the sum over the rows is then a
 * sequential chain per weight -- a verification mode).  GVC_MODE_FAST: gradients reduced over the rows
 * in parallel (deterministic; within 1e-5 of exact). */
typedef struct gvc_trainer gvc_trainer;

/* Same layer description as gvc_model_upload; gradients and velocities start at zero
 * (linear_layer_training ctor, :7-9).  Linear layers of up to 35 x 32. */
int gvc_trainer_create(gvc_ctx *ctx, int n_layers, const int *kinds, const int *rows, const int *cols,
                       const float *const *W, const float *const *bias, gvc_trainer **out);
void gvc_trainer_destroy(gvc_trainer *t);
int gvc_trainer_input_width(const gvc_trainer *t);
int gvc_trainer_output_width(const gvc_trainer *t);

/* model_training::predict (:81-96).  x: n x input width (host); scales: WEIGHT_SCALE of every graph layer
 * (n_scales == 1: one for all); out: n x output width (host; may be NULL -- the output stays on the
 * device for gvc_trainer_mse_backprop). */
int gvc_trainer_predict(gvc_trainer *t, const float *x, const float *scales, int n_scales, float *out, int mode);

/* model_training::backprop (:98-129) of the last predict.  grad_in: n x output width; grad_out (may be
 * NULL): n x input width.  Accumulates into the gradients. */
int gvc_trainer_backprop(gvc_trainer *t, const float *grad_in, float *grad_out, int mode);

/* MSE_loss + MSE_grad + backprop without leaving the device: y goes up, the loss comes back
 * (run_model, old_files/src/apps/gnn_train.cpp:85-99). */
int gvc_trainer_mse_backprop(gvc_trainer *t, const float *y, float *loss, int mode);

int gvc_trainer_sgd_step(gvc_trainer *t, uint64_t batch_size, float lr, float momentum, float weight_decay);
int gvc_trainer_zero_grad(gvc_trainer *t);

/* what: 0 parameters, 1 gradients, 2 velocities of the linear layer at index `layer` (counted over ALL
 * layers); W: rows x cols, bias: cols; either may be NULL. */
int gvc_trainer_read(gvc_trainer *t, int what, int layer, float *W, float *bias);
int gvc_trainer_write(gvc_trainer *t, int what, int layer, const float *W, const float *bias);

/* Single layers with host buffers -- what the drop-in's *_training structs call, like gvc_linear_host
 * and gvc_graph_layer_host for the forward.
 *   linear   :17-26  in n x K (the layer's in_copy), grad_in n x Nout, W K x Nout; grad_W / grad_bias
 *                    are accumulated into; grad_out n x K
 *   graph    :32-42  on the context's graph: grad_in n x (2 width + 3), grad_out n x width
 *   ReLU     :50-53  z >= 0 ? g : 0;   sigmoid :61-65  f(z) (1 - f(z)) g   (z = the layer's input)
 *   MSE      :175-190  loss and / or grad (either may be NULL) of n x w matrices
 *   SGD      :192-224  one parameter array with its gradient and velocity, updated in place */
int gvc_linear_backward_host(gvc_ctx *ctx, uint64_t n, int K, int Nout, const float *in, const float *grad_in,
                             const float *W, float *grad_W, float *grad_bias, float *grad_out, int mode);
int gvc_graph_backward_host(gvc_ctx *ctx, const float *grad_in, int width, float *grad_out);
int gvc_relu_backward_host(gvc_ctx *ctx, uint64_t count, const float *z, const float *grad_in, float *grad_out);
int gvc_sigmoid_backward_host(gvc_ctx *ctx, uint64_t count, const float *z, const float *grad_in, float *grad_out, int mode);
int gvc_mse_host(gvc_ctx *ctx, uint64_t n, int w, const float *x, const float *y, float *loss, float *grad);
int gvc_sgd_host(gvc_ctx *ctx, uint64_t count, float *param, float *grad, float *vel, uint64_t batch_size, float lr,
                 float momentum, float weight_decay);

#ifdef __cplusplus
}
#endif
#endif /* GVC_H */
