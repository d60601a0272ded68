// train_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's TRAINING path, compiled twice by
// oracle/Makefile:
//   oracle/_ref/libgnntrainref.so     with the reference's own, unmodified
//                                     old_files/src/lib/gnn_training.cpp + src/gnn_inference.cpp +
//                                     src/matrix.cpp (where they lie) and the wheel OpenBLAS
//   oracle/_ref/libgnntraindropin.so  with the drop-in host units (gnn-mwvc_b200/host/*.cpp) over libgvc
// so the tests drive both through the same calls -- the reference's C++ interface
// (old_files/include/gnn/gnn_training.hpp) -- and compare what comes back.  The header is the
// reference's; reduction_graph.hpp is taken from /root/reference/include (old_files' copy does not
// compile with g++ 13: std::swap on vector<bool> references, old_files/include/mwvc/reduction_graph.hpp:491;
// the two gnn_inference.hpp / matrix.hpp are identical).  Nothing in the product links this file.
//
// Calls into the reference (old_files/src/lib/gnn_training.cpp):
//   model_training::add_layer :70, ::predict :81-96, ::backprop :98-129
//   the layer structs' forward / backward :11-65
//   MSE_loss :175-182, MSE_grad :184-190, SGD_step :192-224, zero_grad :226-235
//   operator<< / operator>> :131-173
#include "gnn_training.hpp"

#include <cstring>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#ifndef GVC_HARNESS_DROPIN
extern "C" void openblas_set_num_threads(int n);
#endif

namespace {
struct train_state {
    gnn::model_training m;
    reduction_graph<uint32_t, uint32_t> g;
    matrix x, out, grad, grad_x;
    train_state() : m("trained"), g(std::vector<uint32_t>(), std::vector<std::pair<uint32_t, uint32_t>>()) {}
};

void fill(matrix &m, size_t r, size_t c, const float *src) {
    m.resize(r, c);
    for (size_t i = 0; i < r; ++i)
        for (size_t j = 0; j < c; ++j) m(i, j) = src[i * c + j];
}
void spill(const matrix &m, float *dst) {
    for (size_t i = 0; i < m.get_height(); ++i)
        for (size_t j = 0; j < m.get_width(); ++j) dst[i * m.get_width() + j] = m(i, j);
}
}  // namespace

extern "C" {

void trn_blas_threads(int n) {
#ifndef GVC_HARNESS_DROPIN
    openblas_set_num_threads(n);
#else
    (void)n;
#endif
}

// kinds: 0 linear, 1 graph, 2 ReLU, 3 sigmoid (the order of component_training's alternatives).  Linear
// layers are built with the reference's seeded constructor and, if W[i] is given, overwritten.
void *trn_create(int n, const int *kinds, const int *rows, const int *cols, const float *const *W, const float *const *bias,
                 const float *scales, const size_t *seeds) {
    auto *s = new train_state();
    for (int i = 0; i < n; ++i) {
        switch (kinds[i]) {
        case 0: {
            gnn::linear_layer_training l(rows[i], cols[i], seeds ? seeds[i] : 0);
            if (W && W[i]) fill(l.l.W, rows[i], cols[i], W[i]);
            if (bias && bias[i]) fill(l.l.bias, 1, cols[i], bias[i]);
            s->m.add_layer(l);
            break;
        }
        case 1: s->m.add_layer(gnn::graph_layer_training(scales ? scales[i] : 1200.0f)); break;
        case 2: s->m.add_layer(gnn::ReLU_training()); break;
        default: s->m.add_layer(gnn::sigmoid_training()); break;
        }
    }
    return s;
}
void *trn_parse(const char *text) {
    auto *s = new train_state();
    std::istringstream is{std::string(text)};
    is >> s->m;
    return s;
}
void trn_destroy(void *h) { delete static_cast<train_state *>(h); }

size_t trn_text(void *h, char *buf, size_t cap) {
    std::ostringstream os;
    os << static_cast<train_state *>(h)->m;
    const std::string t = os.str();
    if (buf && cap) {
        const size_t k = t.size() < cap - 1 ? t.size() : cap - 1;
        std::memcpy(buf, t.data(), k);
        buf[k] = 0;
    }
    return t.size();
}

// eu[i] < ev[i], sorted, unique: what parse_graph hands to the constructor (gnn_train.cpp:14-32)
void trn_set_graph(void *h, uint32_t n, uint64_t n_edges, const uint32_t *eu, const uint32_t *ev, const uint32_t *w) {
    auto *s = static_cast<train_state *>(h);
    std::vector<uint32_t> weights(w, w + n);
    std::vector<std::pair<uint32_t, uint32_t>> edges(n_edges);
    for (uint64_t i = 0; i < n_edges; ++i) edges[i] = {eu[i], ev[i]};
    s->g = reduction_graph<uint32_t, uint32_t>(weights, edges);
}

int trn_predict(void *h, const float *x, int in_w, float *out, int out_w) {
    auto *s = static_cast<train_state *>(h);
    fill(s->x, s->g.size(), in_w, x);
    s->m.predict(s->x, s->out, s->g);
    if (s->out.get_height() != s->g.size() || (int)s->out.get_width() != out_w) return -1;
    spill(s->out, out);
    return 0;
}

// backprop of the last predict; grad_x may be null.  Returns the width of the input gradient.
int trn_backprop(void *h, const float *grad, int out_w, float *grad_x) {
    auto *s = static_cast<train_state *>(h);
    fill(s->grad, s->g.size(), out_w, grad);
    s->m.backprop(s->grad, s->grad_x, s->g);
    if (grad_x) spill(s->grad_x, grad_x);
    return (int)s->grad_x.get_width();
}

// run_model's training step for one graph (gnn_train.cpp:85-99): loss, MSE_grad, backprop
float trn_mse_step(void *h, const float *y, int out_w) {
    auto *s = static_cast<train_state *>(h);
    matrix ym;
    fill(ym, s->g.size(), out_w, y);
    const float loss = gnn::MSE_loss(s->out, ym);
    gnn::MSE_grad(s->out, ym, s->grad);
    s->m.backprop(s->grad, s->grad_x, s->g);
    return loss;
}

void trn_sgd_step(void *h, size_t batch, float lr, float momentum, float wd) { gnn::SGD_step(static_cast<train_state *>(h)->m, batch, lr, momentum, wd); }
void trn_zero_grad(void *h) { gnn::zero_grad(static_cast<train_state *>(h)->m); }

// what: 0 W/bias, 1 grad_W/grad_bias, 2 vel_W/vel_bias of the linear layer at index `layer`; returns rows * cols
// (0: not a linear layer).  Either pointer may be null.
size_t trn_read(void *h, int what, int layer, float *W, float *bias) {
    auto *s = static_cast<train_state *>(h);
    if (layer < 0 || layer >= (int)s->m.layers.size()) return 0;
    auto *l = std::get_if<gnn::linear_layer_training>(&s->m.layers[layer]);
    if (!l) return 0;
    const matrix &Wm = what == 0 ? l->l.W : what == 1 ? l->grad_W : l->vel_W;
    const matrix &bm = what == 0 ? l->l.bias : what == 1 ? l->grad_bias : l->vel_bias;
    if (W) spill(Wm, W);
    if (bias) spill(bm, bias);
    return Wm.get_height() * Wm.get_width();
}

// ---- single layers through the structs' own forward / backward ---------------------------------------------
// linear: in n x K, grad n x Nout; out n x Nout, grad_W K x Nout, grad_bias Nout (from zero), grad_in n x K
void trn_linear_layer(size_t n, int K, int Nout, const float *W, const float *bias, const float *in, const float *grad, float *out,
                      float *grad_W, float *grad_bias, float *grad_in) {
    gnn::linear_layer_training l(K, Nout, 0);
    fill(l.l.W, K, Nout, W);
    fill(l.l.bias, 1, Nout, bias);
    matrix im, om, gm, gi;
    fill(im, n, K, in);
    fill(gm, n, Nout, grad);
    l.forward(im, om);
    l.backward(gm, gi);
    spill(om, out);
    spill(l.grad_W, grad_W);
    spill(l.grad_bias, grad_bias);
    spill(gi, grad_in);
}

// graph layer on the harness' graph: in n x w, grad n x (2w + 3); out n x (2w + 3), grad_in n x w
void trn_graph_layer(void *h, int w, float scale, const float *in, const float *grad, float *out, float *grad_in) {
    auto *s = static_cast<train_state *>(h);
    gnn::graph_layer_training l(scale);
    matrix im, om, gm, gi;
    fill(im, s->g.size(), w, in);
    fill(gm, s->g.size(), 2 * w + 3, grad);
    l.forward(im, om, s->g);
    l.backward(gm, gi, s->g);
    spill(om, out);
    spill(gi, grad_in);
}

// kind 2 ReLU, 3 sigmoid: z and grad of `count` values
void trn_activation(int kind, size_t count, const float *z, const float *grad, float *out, float *grad_in) {
    matrix zm, om, gm, gi;
    fill(zm, count, 1, z);
    fill(gm, count, 1, grad);
    if (kind == 2) {
        gnn::ReLU_training l;
        l.forward(zm, om);
        l.backward(gm, gi);
    } else {
        gnn::sigmoid_training l;
        l.forward(zm, om);
        l.backward(gm, gi);
    }
    spill(om, out);
    spill(gi, grad_in);
}

float trn_mse(size_t n, int w, const float *x, const float *y, float *grad) {
    matrix xm, ym, gm;
    fill(xm, n, w, x);
    fill(ym, n, w, y);
    gnn::MSE_grad(xm, ym, gm);
    spill(gm, grad);
    return gnn::MSE_loss(xm, ym);
}

}  // extern "C"
