// gvc_gnn_training.cpp -- drop-in translation unit for the reference's
// old_files/src/lib/gnn_training.cpp.
//
// Compiled against the reference's own old_files/include/gnn/gnn_training.hpp: every symbol declared
// there is defined here with the same meaning, so old_files/src/apps/gnn_train.cpp (run_model :68-110)
// builds against it unchanged -- with forward, backward, loss and optimiser step on a B200.
//
//   model_training::predict / backprop   a gvc_trainer (include/gvc.h): one predict = the layer kernels
//                                        with every layer's input kept on the device, one backprop = the
//                                        backward kernels; x / grad in, out / grad out
//   layer ::forward / ::backward         the single-layer host-buffer entry points of libgvc
//   MSE_loss / MSE_grad / SGD_step       gvc_mse_host / gvc_sgd_host
//   zero_grad, parse, print, add_layer   host code (no arithmetic)
//
// The host structs stay the source of truth for parameters, gradients and velocities (callers print
// and edit them): they are written to the device before a predict (when changed) or a backprop, the
// accumulated gradients are read back after it -- 6 209 floats for the GNN_VC model.  The per-layer
// in_copy matrices of the reference are only filled by the layers' own forward(); after
// model_training::predict the saved inputs live on the device.  No OpenBLAS, no CPU arithmetic.
#include "gnn_training.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "gvc.h"
#include "gvc_host_ctx.hpp"

using namespace gnn;

namespace gvc_host {
void upload_graph_of(gvc_ctx *ctx, const reduction_graph<gnn::Tn, gnn::Tw> &g);   // host/gvc_gnn_inference.cpp
}

namespace {

const float *cdata(const matrix &m) { return m.get_height() * m.get_width() ? &m(0, 0) : nullptr; }
float *mdata(matrix &m) { return m.get_height() * m.get_width() ? &m(0, 0) : nullptr; }

template <class... Ts>
struct visitor : Ts... { using Ts::operator()...; };
template <class... Ts>
visitor(Ts...) -> visitor<Ts...>;

int kind_of(const component_training &c) {
    return std::visit(visitor{[](const linear_layer_training &) { return (int)GVC_LINEAR; },
                              [](const graph_layer_training &) { return (int)GVC_GRAPH; },
                              [](const ReLU_training &) { return (int)GVC_RELU; },
                              [](const sigmoid_training &) { return (int)GVC_SIGMOID; }},
                      c);
}

// The device twin of one model_training (a process holds one at a time, like the inference drop-in).
struct twin {
    const model_training *owner = nullptr;
    gvc_trainer *t = nullptr;
    std::vector<std::pair<size_t, size_t>> shapes;       // per layer (rows, cols); (0, 0) for the others
    std::vector<int> kinds;
};
twin &the_twin() {
    static twin tw;
    return tw;
}

bool same_architecture(const twin &tw, const model_training &m) {
    if (tw.owner != &m || !tw.t || tw.kinds.size() != m.layers.size()) return false;
    for (size_t i = 0; i < m.layers.size(); ++i) {
        if (tw.kinds[i] != kind_of(m.layers[i])) return false;
        if (auto *l = std::get_if<linear_layer_training>(&m.layers[i]))
            if (tw.shapes[i] != std::make_pair(l->l.W.get_height(), l->l.W.get_width())) return false;
    }
    return true;
}

// (Re)build the device twin when the model or its architecture changed; then bring `what` (0 parameters,
// 1 gradients) of every linear layer up to date from the host structs.
gvc_trainer *twin_of(const model_training &m, int what) {
    twin &tw = the_twin();
    gvc_ctx *ctx = gvc_host::context();
    if (!same_architecture(tw, m)) {
        if (tw.t) gvc_trainer_destroy(tw.t);
        tw = twin{};
        const int n = (int)m.layers.size();
        std::vector<int> kinds(n), rows(n, 0), cols(n, 0);
        std::vector<const float *> W(n, nullptr), b(n, nullptr);
        tw.shapes.assign(n, {0, 0});
        for (int i = 0; i < n; ++i) {
            kinds[i] = kind_of(m.layers[i]);
            if (auto *l = std::get_if<linear_layer_training>(&m.layers[i])) {
                rows[i] = (int)l->l.W.get_height();
                cols[i] = (int)l->l.W.get_width();
                W[i] = cdata(l->l.W);
                b[i] = cdata(l->l.bias);
                tw.shapes[i] = {l->l.W.get_height(), l->l.W.get_width()};
            }
        }
        const int rc = gvc_trainer_create(ctx, n, kinds.data(), rows.data(), cols.data(), W.data(), b.data(), &tw.t);
        if (rc != 0) gvc_host::die("gvc_trainer_create", rc);
        tw.owner = &m;
        tw.kinds = kinds;
    }
    for (size_t i = 0; i < m.layers.size(); ++i)
        if (auto *l = std::get_if<linear_layer_training>(&m.layers[i])) {
            const matrix &Wm = what == 0 ? l->l.W : l->grad_W, &bm = what == 0 ? l->l.bias : l->grad_bias;
            if (Wm.get_height() * Wm.get_width() != tw.shapes[i].first * tw.shapes[i].second || bm.get_width() != tw.shapes[i].second)
                continue;                                   // gradient matrices of a parsed model are empty until sized below
            const int rc = gvc_trainer_write(tw.t, what, (int)i, cdata(Wm), cdata(bm));
            if (rc != 0) gvc_host::die("gvc_trainer_write", rc);
        }
    return tw.t;
}

std::vector<float> graph_scales(const model_training &m) {
    std::vector<float> s;
    for (auto &c : m.layers)
        if (auto *gl = std::get_if<graph_layer_training>(&c)) s.push_back(gl->l.WEIGHT_SCALE);
    if (s.empty()) s.push_back(1.0f);
    return s;
}

void keep(const matrix &in, matrix &copy) {                 // in_copy = in  (:12-13, :44-45, :55-56)
    copy.resize(in.get_height(), in.get_width());
    const size_t count = in.get_height() * in.get_width();
    if (count) std::memcpy(mdata(copy), cdata(in), count * sizeof(float));
}

// a model read with operator>> has default-constructed (empty) gradient and velocity matrices, exactly as
// in the reference (:157-173); the reference's dot() resizes grad_W on first use, the bias sum writes
// through raw() iterators of an EMPTY matrix (nothing happens).  We size all four to the layer's shape.
void size_state(const linear_layer_training &l) {
    const size_t r = l.l.W.get_height(), c = l.l.W.get_width();
    l.grad_W.resize(r, c);
    l.grad_bias.resize(1, c);
    l.vel_W.resize(r, c);
    l.vel_bias.resize(1, c);
}

}  // namespace

// ---- linear (old_files/src/lib/gnn_training.cpp:7-26) --------------------------------------------------
linear_layer_training::linear_layer_training(size_t dim_in, size_t dim_out, size_t seed)
    : l(dim_in, dim_out, seed), grad_W(dim_in, dim_out), grad_bias(1, dim_out), vel_W(dim_in, dim_out), vel_bias(1, dim_out) {}

void linear_layer_training::forward(const matrix &in, matrix &out) const {
    keep(in, in_copy);
    l.forward(in, out);
}

void linear_layer_training::backward(const matrix &grad_in, matrix &grad_out) const {
    const size_t n = grad_in.get_height(), K = l.W.get_height(), Nout = l.W.get_width();
    size_state(*this);
    grad_out.resize(n, K);
    const int rc = gvc_linear_backward_host(gvc_host::context(), n, (int)K, (int)Nout, cdata(in_copy), cdata(grad_in), cdata(l.W),
                                            mdata(grad_W), mdata(grad_bias), mdata(grad_out), gvc_host::mode());
    if (rc != 0) gvc_host::die("gvc_linear_backward_host", rc);
}

// ---- graph (:28-42) ----------------------------------------------------------------------------------------
void graph_layer_training::forward(const matrix &in, matrix &out, const reduction_graph<Tn, Tw> &g) const { l.forward(in, out, g); }

void graph_layer_training::backward(const matrix &grad_in, matrix &grad_out, const reduction_graph<Tn, Tw> &g) const {
    const size_t n = grad_in.get_height(), w = (grad_in.get_width() - 3) / 2;
    grad_out.resize(n, w);
    if (!n || !w) return;
    gvc_ctx *ctx = gvc_host::context();
    gvc_host::upload_graph_of(ctx, g);
    const int rc = gvc_graph_backward_host(ctx, cdata(grad_in), (int)w, mdata(grad_out));
    if (rc != 0) gvc_host::die("gvc_graph_backward_host", rc);
}

// ---- activations (:44-65) -------------------------------------------------------------------------------
void ReLU_training::forward(const matrix &in, matrix &out) const {
    keep(in, in_copy);
    l.forward(in, out);
}
void ReLU_training::backward(const matrix &grad_in, matrix &grad_out) const {
    grad_out.resize(grad_in.get_height(), grad_in.get_width());
    const size_t count = grad_in.get_height() * grad_in.get_width();
    const int rc = gvc_relu_backward_host(gvc_host::context(), count, cdata(in_copy), cdata(grad_in), mdata(grad_out));
    if (rc != 0) gvc_host::die("gvc_relu_backward_host", rc);
}
void sigmoid_training::forward(const matrix &in, matrix &out) const {
    keep(in, in_copy);
    l.forward(in, out);
}
void sigmoid_training::backward(const matrix &grad_in, matrix &grad_out) const {
    grad_out.resize(grad_in.get_height(), grad_in.get_width());
    const size_t count = grad_in.get_height() * grad_in.get_width();
    const int rc = gvc_sigmoid_backward_host(gvc_host::context(), count, cdata(in_copy), cdata(grad_in), mdata(grad_out), gvc_host::mode());
    if (rc != 0) gvc_host::die("gvc_sigmoid_backward_host", rc);
}

// ---- model (:67-129) -------------------------------------------------------------------------------------
model_training::model_training(std::string name) : name(name) {}

void model_training::add_layer(const component_training &c) { layers.push_back(c); }

void model_training::predict(const matrix &in, matrix &out, const reduction_graph<Tn, Tw> &g) const {
    const Tn n = g.size();
    if (layers.empty()) return;
    gvc_trainer *t = twin_of(*this, 0);
    out.resize(n, (size_t)gvc_trainer_output_width(t));
    gvc_host::upload_graph_of(gvc_host::context(), g);
    const std::vector<float> scales = graph_scales(*this);
    const int rc = gvc_trainer_predict(t, cdata(in), scales.data(), (int)scales.size(), mdata(out), gvc_host::mode());
    if (rc != 0) gvc_host::die("gvc_trainer_predict", rc);
}

void model_training::backprop(const matrix &grad_in, matrix &grad_out, const reduction_graph<Tn, Tw> &g) const {
    if (layers.empty()) return;
    for (auto &c : layers)
        if (auto *l = std::get_if<linear_layer_training>(&c)) size_state(*l);
    gvc_trainer *t = twin_of(*this, 1);                      // the gradients accumulated so far
    grad_out.resize(g.size(), (size_t)gvc_trainer_input_width(t));
    int rc = gvc_trainer_backprop(t, cdata(grad_in), mdata(grad_out), gvc_host::mode());
    if (rc != 0) gvc_host::die("gvc_trainer_backprop", rc);
    for (size_t i = 0; i < layers.size(); ++i)
        if (auto *l = std::get_if<linear_layer_training>(&layers[i])) {
            rc = gvc_trainer_read(t, 1, (int)i, mdata(l->grad_W), mdata(l->grad_bias));
            if (rc != 0) gvc_host::die("gvc_trainer_read", rc);
        }
}

// ---- text format (:131-173; the same records as the inference model) ------------------
std::ostream &gnn::operator<<(std::ostream &os, const model_training &m) {
    os << m.name << std::endl << m.layers.size() << " Layers" << std::endl;
    for (auto &c : m.layers) {
        switch (kind_of(c)) {
        case GVC_LINEAR: {
            const linear_layer_training &l = std::get<linear_layer_training>(c);
            os << "Linear_Layer" << std::endl
               << "Weights: " << l.l.W << std::endl
               << "Bias: " << l.l.bias << std::endl;
            break;
        }
        case GVC_GRAPH: os << "Graph_Layer" << std::endl; break;
        case GVC_RELU: os << "ReLU_Activation" << std::endl; break;
        default: os << "Sigmoid_Activation" << std::endl; break;
        }
        os << std::endl;
    }
    return os;
}

std::istream &gnn::operator>>(std::istream &is, model_training &m) {
    gvc_host::warm_start();
    size_t count = 0;
    std::string word;
    is >> m.name >> count >> word;
    for (size_t i = 0; i < count && is; ++i) {
        is >> word;
        if (word == "Graph_Layer") m.layers.emplace_back(graph_layer_training());
        else if (word == "ReLU_Activation") m.layers.emplace_back(ReLU_training());
        else if (word == "Sigmoid_Activation") m.layers.emplace_back(sigmoid_training());
        else if (word == "Linear_Layer") {
            linear_layer_training l;
            is >> word >> l.l.W >> word >> l.l.bias;
            m.layers.emplace_back(std::move(l));
        }
    }
    return is;
}

// ---- loss and optimiser (:175-235) ---------------------------------------------------------------------
float gnn::MSE_loss(const matrix &x, const matrix &y) {
    float loss = 0.0f;
    const int rc = gvc_mse_host(gvc_host::context(), x.get_height(), (int)x.get_width(), cdata(x), cdata(y), &loss, nullptr);
    if (rc != 0) gvc_host::die("gvc_mse_host", rc);
    return loss;
}

void gnn::MSE_grad(const matrix &x, const matrix &y, matrix &grad_out) {
    grad_out.resize(x.get_height(), x.get_width());
    if (!(x.get_height() * x.get_width())) return;
    const int rc = gvc_mse_host(gvc_host::context(), x.get_height(), (int)x.get_width(), cdata(x), cdata(y), nullptr, mdata(grad_out));
    if (rc != 0) gvc_host::die("gvc_mse_host", rc);
}

void gnn::SGD_step(model_training &m, size_t batch_size, float lr, float momentum, float weight_decay) {
    gvc_ctx *ctx = gvc_host::context();
    for (auto &c : m.layers)
        if (auto *l = std::get_if<linear_layer_training>(&c)) {
            size_state(*l);
            int rc = gvc_sgd_host(ctx, l->l.W.get_height() * l->l.W.get_width(), mdata(l->l.W), mdata(l->grad_W), mdata(l->vel_W),
                                  batch_size, lr, momentum, weight_decay);
            if (rc == 0)
                rc = gvc_sgd_host(ctx, l->l.bias.get_width(), mdata(l->l.bias), mdata(l->grad_bias), mdata(l->vel_bias), batch_size, lr,
                                  momentum, weight_decay);
            if (rc != 0) gvc_host::die("gvc_sgd_host", rc);
        }
}

void gnn::zero_grad(model_training &m) {
    for (auto &c : m.layers)
        if (auto *l = std::get_if<linear_layer_training>(&c)) {
            std::fill(l->grad_W.raw().begin(), l->grad_W.raw().end(), 0.0f);
            std::fill(l->grad_bias.raw().begin(), l->grad_bias.raw().end(), 0.0f);
        }
}
