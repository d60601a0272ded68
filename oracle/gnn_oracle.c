/*
 * gnn_oracle.c -- TEST INFRASTRUCTURE ONLY (see gnn_oracle.h).
 *
 * CPU restatement of the reference GNN forward.  Every function names the
 * reference lines it follows.  Must be compiled with -ffp-contract=off and
 * without -ffast-math: the whole point is the exact fp32 operation order
 * (SURVEY.md App. B); oracle/Makefile does that.
 *
 * Pinned against the unmodified reference built into oracle/_ref/ (reference
 * sources + OpenBLAS 0.3.15 "Prescott" sgemm): tests/test_oracle_vs_ref.py runs
 * both on seeded graphs and requires bit-equal scores; tests/golden/ holds the
 * vectors produced by that build (tools/make_golden.py).
 */
#include "gnn_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int kind;
    int rows, cols; /* linear: K x Nout */
    float *W;       /* rows*cols, row-major */
    float *bias;    /* cols */
    float weight_scale; /* graph layer: graph_layer::WEIGHT_SCALE, gnn_inference.hpp:25 */
} gvo_layer;

struct gvo_model {
    char name[128];
    int n_layers;
    gvo_layer *layers;
};

/* ---- tokenizer: istream >> std::string semantics (whitespace separated) ---- */
static const char *next_token(const char *p, char *buf, size_t cap) {
    while (*p && isspace((unsigned char)*p)) p++;
    if (!*p) { buf[0] = 0; return p; }
    size_t n = 0;
    while (*p && !isspace((unsigned char)*p)) {
        if (n + 1 < cap) buf[n++] = *p;
        p++;
    }
    buf[n] = 0;
    return p;
}

/* matrix operator>>, matrix.cpp:97-104: "h w" then h*w floats row-major. */
static const char *parse_matrix(const char *p, int *h, int *w, float **data) {
    char *end;
    long hh = strtol(p, &end, 10); p = end;
    long ww = strtol(p, &end, 10); p = end;
    if (hh < 0 || ww < 0) { hh = 0; ww = 0; }
    *h = (int)hh; *w = (int)ww;
    size_t cnt = (size_t)hh * (size_t)ww;
    *data = (float *)calloc(cnt ? cnt : 1, sizeof(float));
    if (!*data) return NULL;
    for (size_t i = 0; i < cnt; i++) {
        (*data)[i] = strtof(p, &end); /* iostream num_get -> strtof, correctly rounded */
        if (end == p) break;          /* stream failure: remaining stay 0 */
        p = end;
    }
    return p;
}

/* operator>>(istream&, model&), gnn_inference.cpp:120-139 */
gvo_model *gvo_model_parse(const char *text) {
    gvo_model *m = (gvo_model *)calloc(1, sizeof(*m));
    if (!m) return NULL;
    char tok[256];
    const char *p = text;
    p = next_token(p, m->name, sizeof(m->name));   /* is >> m.name */
    char *end;
    long n = strtol(p, &end, 10); p = end;          /* >> n */
    p = next_token(p, tok, sizeof(tok));            /* >> tmp ("Layers") */
    if (n < 0) n = 0;
    m->layers = (gvo_layer *)calloc((size_t)n ? (size_t)n : 1, sizeof(gvo_layer));
    if (!m->layers) { free(m); return NULL; }
    for (long i = 0; i < n; i++) {
        p = next_token(p, tok, sizeof(tok));
        gvo_layer *l = &m->layers[m->n_layers];
        l->weight_scale = 120.0f; /* gnn_inference.hpp:25 */
        if (!strcmp(tok, "Linear_Layer")) {
            l->kind = GVO_LINEAR;
            int bh, bw;
            p = next_token(p, tok, sizeof(tok));   /* "Weights:" */
            p = parse_matrix(p, &l->rows, &l->cols, &l->W);
            if (!p) { gvo_model_free(m); return NULL; }
            p = next_token(p, tok, sizeof(tok));   /* "Bias:" */
            p = parse_matrix(p, &bh, &bw, &l->bias);
            if (!p) { gvo_model_free(m); return NULL; }
            m->n_layers++;
        } else if (!strcmp(tok, "Graph_Layer")) {
            l->kind = GVO_GRAPH; m->n_layers++;
        } else if (!strcmp(tok, "ReLU_Activation")) {
            l->kind = GVO_RELU; m->n_layers++;
        } else if (!strcmp(tok, "Sigmoid_Activation")) {
            l->kind = GVO_SIGMOID; m->n_layers++;
        } /* unknown tokens are skipped but still count, :125-136 */
    }
    return m;
}

void gvo_model_free(gvo_model *m) {
    if (!m) return;
    for (int i = 0; i < m->n_layers; i++) { free(m->layers[i].W); free(m->layers[i].bias); }
    free(m->layers);
    free(m);
}

int gvo_model_num_layers(const gvo_model *m) { return m->n_layers; }

int gvo_model_layer(const gvo_model *m, int i, int *rows, int *cols,
                    const float **W, const float **bias) {
    const gvo_layer *l = &m->layers[i];
    if (rows) *rows = l->rows;
    if (cols) *cols = l->cols;
    if (W) *W = l->W;
    if (bias) *bias = l->bias;
    return l->kind;
}

/* model::set_weight_scale, gnn_inference.cpp:83-90 */
void gvo_model_set_weight_scale(gvo_model *m, float ws) {
    for (int i = 0; i < m->n_layers; i++)
        if (m->layers[i].kind == GVO_GRAPH) m->layers[i].weight_scale = ws;
}

/* graph_layer::WEIGHT_SCALE is a member of every graph layer (gnn_inference.hpp:25): a model built with
 * add_layer may carry a different one per layer.  which = 0, 1, ... counts graph layers. */
int gvo_model_set_graph_layer_scale(gvo_model *m, int which, float ws) {
    for (int i = 0; i < m->n_layers; i++)
        if (m->layers[i].kind == GVO_GRAPH && which-- == 0) { m->layers[i].weight_scale = ws; return 0; }
    return -1;
}

float gvo_model_weight_scale(const gvo_model *m) {
    for (int i = 0; i < m->n_layers; i++)
        if (m->layers[i].kind == GVO_GRAPH) return m->layers[i].weight_scale;
    return 120.0f;
}

/* graph_layer::forward, gnn_inference.cpp:27-42.
 * Note the layout quirk (:38-40): D, W/s, NW/s land at columns w+1..w+3, i.e.
 * on top of self features 1..3 when w>1, and the last 3 columns stay 0. */
void gvo_graph_forward(uint32_t n, const uint64_t *row_ptr, const uint32_t *col,
                       const uint32_t *W, const uint32_t *NW, float scale,
                       const float *in, int w, float *out) {
    const int ow = 2 * w + 3;
    for (size_t i = 0; i < (size_t)n * (size_t)ow; i++) out[i] = 0.0f;   /* :29 */
    for (uint32_t u = 0; u < n; u++) {
        float *o = out + (size_t)u * ow;
        for (uint64_t e = row_ptr[u]; e < row_ptr[u + 1]; e++) {            /* :32-36 */
            const float *r = in + (size_t)col[e] * w;
            for (int c = 0; c < w; c++) o[c] = r[c] + o[c];
        }
        const float *s = in + (size_t)u * w;
        for (int c = 0; c < w; c++) o[w + c] = s[c];                        /* :37 */
        o[w + 1] = (float)(uint32_t)(row_ptr[u + 1] - row_ptr[u]);          /* :38 */
        o[w + 2] = (float)W[u] / scale;                                     /* :39 */
        o[w + 3] = (float)NW[u] / scale;                                    /* :40 */
    }
}

/* ---- dot(), matrix.cpp:106-122: cblas_sgemm, alpha = 1 -------------------------------------
 *
 * The arithmetic lives in OpenBLAS (un-vendored; the reference pins no version).  What is
 * restated is the operation order of OpenBLAS 0.3.15, kernel set "Prescott" (sgemm 8x4 SSE3
 * micro-kernel: no FMA, product and sum rounded separately), run with ONE thread; established by
 * experiment against oracle/_ref (probes over every block class, transposes, K up to 2500) and
 * re-checked by tests/test_oracle_vs_ref.py.  In row-major terms, C[m x n] = op(A) op(B):
 *
 *   rows are taken 4 at a time, then 2 (if m & 2), then 1 (if m & 1);
 *   columns 8 at a time, then 4 (if n & 4), 2 (if n & 2), 1 (if n & 1);
 *   the micro-kernel of a (row class, column class) pair fixes how the k sum of each element is
 *   split over accumulators (transposes only change the packing, not the sums):
 *
 *                      8 cols   4 cols   2 cols   1 col
 *        4 rows        SEQ      SEQ      TWO      TWO
 *        2 rows        SEQ      TWO      TWO      TWO
 *        1 row         TWO      TWO      EIGHT    FOUR
 *
 *     SEQ    one accumulator, k ascending
 *     TWO    even / odd k below floor(K/8)*8, tail to the even one, even + odd
 *     FOUR   k mod 4 below floor(K/8)*8, tail to the first, (a0+a1)+(a2+a3)
 *     EIGHT  k mod 8 below floor(K/16)*16, tail to the first, ((a0+a2)+(a4+a6))+((a1+a3)+(a5+a7))
 *
 *   k is cut into blocks (GEMM_Q = 128: a remainder of >= 256 takes 128, one of 129..255 is halved
 *   and rounded up to a multiple of 8, the rest goes in one piece); every block is one kernel call
 *   with the sums above over its slice, added to C in turn; C is first scaled by beta (set to zero
 *   for beta == 0).
 * With several OpenBLAS threads the rows are cut into per-thread slices, each with its own 2-/1-row
 * tail, so a handful of rows per slice move by an ulp; parity is pinned at OPENBLAS_NUM_THREADS=1. */
enum { SUM_SEQ = 0, SUM_TWO = 1, SUM_FOUR = 2, SUM_EIGHT = 3 };

static int row_class(size_t i, size_t m) {
    size_t m4 = m & ~(size_t)3;
    if (i < m4) return 0;
    if ((m & 2) && i < m4 + 2) return 1;
    return 2;
}

static int col_class(size_t j, size_t n) {
    size_t p = n & ~(size_t)7;
    if (j < p) return 0;
    if (n & 4) { if (j < p + 4) return 1; p += 4; }
    if (n & 2) { if (j < p + 2) return 2; p += 2; }
    return 3;
}

static const int sum_scheme[3][4] = {{SUM_SEQ, SUM_SEQ, SUM_TWO, SUM_TWO},
                                     {SUM_SEQ, SUM_TWO, SUM_TWO, SUM_TWO},
                                     {SUM_TWO, SUM_TWO, SUM_EIGHT, SUM_FOUR}};

/* one k block [lo, hi) of one element: a[k * sa], b[k * sb] */
static float block_sum(const float *a, size_t sa, const float *b, size_t sb, size_t lo, size_t hi, int scheme) {
    const size_t K = hi - lo;
    a += lo * sa; b += lo * sb;
    size_t k = 0;
    if (scheme == SUM_SEQ) {
        float acc = 0.0f;
        for (; k < K; k++) acc = acc + a[k * sa] * b[k * sb];
        return acc;
    }
    float c[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const size_t w = scheme == SUM_TWO ? 2 : scheme == SUM_FOUR ? 4 : 8;
    const size_t body = scheme == SUM_EIGHT ? K / 16 * 16 : K / 8 * 8;
    for (; k < body; k++) c[k % w] = c[k % w] + a[k * sa] * b[k * sb];
    for (; k < K; k++) c[0] = c[0] + a[k * sa] * b[k * sb];
    if (scheme == SUM_TWO) return c[0] + c[1];
    if (scheme == SUM_FOUR) return (c[0] + c[1]) + (c[2] + c[3]);
    return ((c[0] + c[2]) + (c[4] + c[6])) + ((c[1] + c[3]) + (c[5] + c[7]));
}

/* dot(A, B, C, at, bt, beta), matrix.cpp:106-122.  A is stored k x m when at, else m x k; B n x k
 * when bt, else k x n; C m x n; all row-major and dense. */
void gvo_dot(int at, int bt, size_t m, size_t n, size_t k, const float *A, const float *B, float beta, float *C) {
    const size_t sa = at ? m : 1, sb = bt ? 1 : n;
    for (size_t i = 0; i < m; i++)
        for (size_t j = 0; j < n; j++) {
            const float *a = at ? A + i : A + i * k;
            const float *b = bt ? B + j * k : B + j;
            const int scheme = sum_scheme[row_class(i, m)][col_class(j, n)];
            float c = beta == 0.0f ? 0.0f : beta * C[i * n + j];
            size_t lo = 0;
            while (lo < k) {
                size_t len = k - lo;
                if (len >= 256) len = 128;
                else if (len > 128) len = (len / 2 + 7) / 8 * 8;
                c = c + block_sum(a, sa, b, sb, lo, lo + len, scheme);
                lo += len;
            }
            C[i * n + j] = c;
        }
}

/* linear_layer::forward, gnn_inference.cpp:20-25: dot(in, W, out) with beta = 0, then the
 * row-wise bias add (:22-24). */
void gvo_linear_forward(size_t n, int K, int Nout, const float *in,
                        const float *Wm, const float *bias, float *out) {
    gvo_dot(0, 0, n, (size_t)Nout, (size_t)K, in, Wm, 0.0f, out);
    for (size_t i = 0; i < n; i++)
        for (int j = 0; j < Nout; j++) out[i * (size_t)Nout + j] = out[i * (size_t)Nout + j] + bias[j];
}

/* ReLU::forward, gnn_inference.cpp:44-47: std::max(x, 0.0f) == (x < 0) ? 0 : x */
void gvo_relu_forward(size_t count, const float *in, float *out) {
    for (size_t i = 0; i < count; i++) out[i] = (in[i] < 0.0f) ? 0.0f : in[i];
}

/* sigmoid::forward, gnn_inference.cpp:49-52 (glibc expf) */
void gvo_sigmoid_forward(size_t count, const float *in, float *out) {
    for (size_t i = 0; i < count; i++) out[i] = 1.0f / (1.0f + expf(-in[i]));
}

int gvo_model_out_width(const gvo_model *m, int in_width) {
    int w = in_width;
    for (int i = 0; i < m->n_layers; i++) {
        if (m->layers[i].kind == GVO_LINEAR) w = m->layers[i].cols;
        else if (m->layers[i].kind == GVO_GRAPH) w = 2 * w + 3;
    }
    return w;
}

/* model::predict, gnn_inference.cpp:67-81: copy the input, run every layer in
 * order between two ping-pong buffers, leave the result in out. */
int gvo_predict(const gvo_model *m, uint32_t n, const uint64_t *row_ptr,
                const uint32_t *col, const uint32_t *W, const uint32_t *NW,
                const float *x, int in_width, float *out, int *out_width) {
    if (m->n_layers == 0) return -1;                                       /* :68 */
    int maxw = in_width, w = in_width;
    for (int i = 0; i < m->n_layers; i++) {
        if (m->layers[i].kind == GVO_LINEAR) w = m->layers[i].cols;
        else if (m->layers[i].kind == GVO_GRAPH) w = 2 * w + 3;
        if (w > maxw) maxw = w;
    }
    size_t cap = (size_t)n * (size_t)maxw;
    float *a = (float *)malloc((cap ? cap : 1) * sizeof(float));
    float *b = (float *)malloc((cap ? cap : 1) * sizeof(float));
    if (!a || !b) { free(a); free(b); return -1; }
    memcpy(a, x, (size_t)n * in_width * sizeof(float));                    /* :70-71 */
    w = in_width;
    for (int i = 0; i < m->n_layers; i++) {                                /* :73-79 */
        const gvo_layer *l = &m->layers[i];
        int wo = w;
        switch (l->kind) {
        case GVO_LINEAR:
            wo = l->cols;
            gvo_linear_forward(n, l->rows, l->cols, a, l->W, l->bias, b);
            break;
        case GVO_GRAPH:
            wo = 2 * w + 3;
            gvo_graph_forward(n, row_ptr, col, W, NW, l->weight_scale, a, w, b);
            break;
        case GVO_RELU:
            gvo_relu_forward((size_t)n * w, a, b);
            break;
        default:
            gvo_sigmoid_forward((size_t)n * w, a, b);
            break;
        }
        float *t = a; a = b; b = t;
        w = wo;
    }
    memcpy(out, a, (size_t)n * w * sizeof(float));                         /* :80 */
    if (out_width) *out_width = w;
    free(a); free(b);
    return 0;
}

/* GNN_VC.cpp:213,220: the commit decision is out(u,0) > 0.5f */
void gvo_decisions(size_t n, const float *scores, uint8_t *take) {
    for (size_t i = 0; i < n; i++) take[i] = scores[i] > 0.5f;
}
