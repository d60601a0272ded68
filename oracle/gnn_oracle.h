/*
 * gnn_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's GNN forward
 * (gnn::model::predict, /root/reference/src/gnn_inference.cpp:67-81 and the
 * layers it dispatches to, /root/reference/src/matrix.cpp:106-122 for the
 * dense product).  It is the checker the CUDA path is compared against; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * The product (libgvc.so) never links, loads or calls anything in oracle/.
 *
 * Pinning: the reference has no tests or golden vectors for this path
 * (SURVEY.md section 4), so the restatement is pinned against the reference
 * itself, compiled unmodified into oracle/_ref/ (see oracle/Makefile) and run
 * in this container; tests/golden/ holds the vectors that run produced.
 */
#ifndef GNN_ORACLE_H
#define GNN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layer kinds, in the order of the reference's std::variant
 * (/root/reference/include/gnn_inference.hpp:38). */
enum { GVO_LINEAR = 0, GVO_GRAPH = 1, GVO_RELU = 2, GVO_SIGMOID = 3 };

typedef struct gvo_model gvo_model;

/* Parse the reference's text model format (operator>>, gnn_inference.cpp:120-139;
 * matrix operator>>, matrix.cpp:97-104).  Returns NULL on allocation failure. */
gvo_model *gvo_model_parse(const char *text);
void gvo_model_free(gvo_model *m);

int gvo_model_num_layers(const gvo_model *m);
/* kind of layer i; for GVO_LINEAR also rows(K)/cols(Nout) and pointers to the
 * row-major K x Nout weights and the 1 x Nout bias (owned by the model). */
int gvo_model_layer(const gvo_model *m, int i, int *rows, int *cols,
                    const float **W, const float **bias);
/* set_weight_scale, gnn_inference.cpp:83-90 */
void gvo_model_set_weight_scale(gvo_model *m, float ws);
float gvo_model_weight_scale(const gvo_model *m);
/* WEIGHT_SCALE of the which-th graph layer alone (a model built with add_layer may mix them). */
int gvo_model_set_graph_layer_scale(gvo_model *m, int which, float ws);

/* graph_layer::forward, gnn_inference.cpp:27-42, on a CSR view of the graph
 * (row_ptr[n+1], col = neighbour ids in the order g.begin(u)..g.end(u) yields
 * them, W/NW = g.W(u)/g.NW(u)).  in: n x w, out: n x (2w+3), both row-major. */
void gvo_graph_forward(uint32_t n, const uint64_t *row_ptr, const uint32_t *col,
                       const uint32_t *W, const uint32_t *NW, float scale,
                       const float *in, int w, float *out);

/* linear_layer::forward, gnn_inference.cpp:20-25: out = in*W (the sgemm of
 * matrix.cpp:112 restated in the operation order that is bit-identical to
 * OpenBLAS 0.3.15 "Prescott" for these shapes) + bias. */
void gvo_linear_forward(size_t n, int K, int Nout, const float *in,
                        const float *Wm, const float *bias, float *out);

/* dot(), matrix.cpp:106-122 (cblas_sgemm with alpha = 1): C[m x n] = op(A) op(B) + beta C in the
 * operation order of OpenBLAS 0.3.15 "Prescott", one thread, for ANY shape (block classes, k
 * blocking; see gnn_oracle.c).  A stored k x m when at, B stored n x k when bt; row-major, dense. */
void gvo_dot(int at, int bt, size_t m, size_t n, size_t k, const float *A, const float *B, float beta, float *C);

/* ReLU::forward :44-47 and sigmoid::forward :49-52 (elementwise, count floats) */
void gvo_relu_forward(size_t count, const float *in, float *out);
void gvo_sigmoid_forward(size_t count, const float *in, float *out);

/* model::predict, gnn_inference.cpp:67-81.  x: n x in_width input, out gets
 * n x (width of last layer) floats; *out_width receives that width.
 * Returns 0, or -1 on allocation failure / empty model. */
int gvo_predict(const gvo_model *m, uint32_t n, const uint64_t *row_ptr,
                const uint32_t *col, const uint32_t *W, const uint32_t *NW,
                const float *x, int in_width, float *out, int *out_width);

/* Width of the output of predict for an input of width in_width. */
int gvo_model_out_width(const gvo_model *m, int in_width);

/* Caller-side ordering of GNN_VC.cpp:194-206 (eps-tolerance comparator fed to
 * std::sort) cannot be restated without libstdc++'s introsort; what the oracle
 * offers instead is the decision bit out>0.5f used at GNN_VC.cpp:213,220. */
void gvo_decisions(size_t n, const float *scores, uint8_t *take);

#ifdef __cplusplus
}
#endif
#endif
