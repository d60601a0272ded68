// gvc_train_api.cuh -- C ABI of the training path (included at the end of
// gvc_api.cu, whose context, generic forward kernels and OpenBLAS-ordered dot it uses).
//
// Reference: old_files/src/lib/gnn_training.cpp (model_training::predict :81-96, ::backprop :98-129, the
// four layer backwards :17-65, MSE_loss / MSE_grad :175-190, SGD_step :192-224, zero_grad :226-235).
//
//   gvc_trainer       a model_training on the device: parameters, gradients, velocities, and -- between a
//                     predict and its backprop -- every layer's input (what the reference keeps in in_copy).
//                     One predict = 21 generic layer kernels (the same ones predict() of a non-fused model
//                     runs: same bits), one backprop = one kernel per layer kind; nothing but x / y / the
//                     loss crosses PCIe.  The graph is the context's current graph.
//   *_backward_host   the single-layer entry points the drop-in's layer structs call (host buffers in and
//                     out), mirroring gvc_linear_host / gvc_graph_layer_host of the forward.
#pragma once
#include "gvc_train.cuh"

struct gvc_trainer {
    struct Layer {
        int kind = 0, rows = 0, cols = 0;
        float *W = nullptr, *b = nullptr, *gW = nullptr, *gb = nullptr, *vW = nullptr, *vb = nullptr;   // into `params`
        float *saved = nullptr;              // the layer's input of the last predict (linear, ReLU, sigmoid)
        size_t saved_cap = 0;
        int in_w = 0;                        // its width
    };
    gvc_ctx *c = nullptr;
    std::vector<Layer> L;
    float *params = nullptr;                 // [W b | gW gb | vW vb] of all linear layers, one allocation
    size_t n_params = 0;
    float *act[2] = {nullptr, nullptr}, *grad[2] = {nullptr, nullptr};
    size_t act_cap = 0, grad_cap = 0;
    float *partial = nullptr;                // per-CTA partial sums of grad_W / grad_bias
    double *loss_partial = nullptr;
    float *d_out = nullptr;                  // output of the last predict (one of act[])
    uint32_t n_fwd = 0;                      // vertices of the last predict (0: nothing to back-propagate)
    int in_w = 1, out_w = 1, max_w = 1;
};

namespace {

constexpr int kTrainParts = 592;             // CTAs of the grad_W reduction (4 per SM)

int train_grow(float **p, size_t *cap, size_t want) {
    if (want <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    const size_t w = want + want / 8 + 64;
    cudaError_t e = cudaMalloc(p, w * sizeof(float));
    if (e != cudaSuccess) return fail(GVC_ERR_ALLOC, "cudaMalloc(%zu bytes): %s", w * sizeof(float), cudaGetErrorString(e));
    *cap = w;
    return 0;
}

int trainer_check(const gvc_trainer *t) {
    if (!t || !t->c) return fail(GVC_ERR_ARG, "null trainer");
    return 0;
}

// one layer backward on the device.  z: the layer's saved input; g: gradient of its output; out: gradient of
// its input (n x in_w); linear layers also accumulate into gW / gb.
int linear_backward_device(gvc_ctx *c, uint64_t n, int K, int Nout, const float *in, const float *g, const float *W,
                           float *gW, float *gb, float *out, float *partial, int mode) {
    if (mode == GVC_MODE_EXACT) {
        // dot(in_copy, grad_in, grad_W, true, false, 1.0f) :19, in OpenBLAS' order (m = K, n = Nout, k = rows)
        generic_sgemm_kernel<<<blocks_for((uint64_t)K * Nout, 64), 64, 0, c->stream>>>(1, 0, (uint64_t)K, (uint64_t)Nout, n, in, (uint64_t)K,
                                                                                   g, (uint64_t)Nout, 1.0f, gW, (uint64_t)Nout);
        train_bias_grad_exact_kernel<<<blocks_for(Nout, 32), 32, 0, c->stream>>>(g, Nout, n, gb);
        // dot(grad_in, l.W, grad_out, false, true, 0.0f) :25 (m = rows, n = K, k = Nout)
        if (out)
            generic_sgemm_kernel<<<blocks_for(n * K, 256), 256, 0, c->stream>>>(0, 1, n, (uint64_t)K, (uint64_t)Nout, g, (uint64_t)Nout, W,
                                                                              (uint64_t)Nout, 0.0f, out, (uint64_t)K);
        c->launches += out ? 3 : 2;
    } else {
        const int parts = (int)std::min<uint64_t>(kTrainParts, (n + kTrainSlab - 1) / kTrainSlab);
        const size_t smem = (size_t)kTrainSlab * (K + 1 + Nout) * sizeof(float);
        train_linear_dw_partial_kernel<<<parts, 256, smem, c->stream>>>(in, g, K, Nout, n, partial);
        train_linear_dw_final_kernel<<<blocks_for((uint64_t)(K + 1) * Nout, 128), 128, 0, c->stream>>>(partial, parts, K, Nout, gW, gb);
        if (out)
            train_linear_dx_kernel<<<blocks_for(n * K, 256), 256, (size_t)K * Nout * sizeof(float), c->stream>>>(g, K, Nout, W, out, n);
        c->launches += out ? 3 : 2;
    }
    GVC_CUDA(cudaGetLastError());
    return 0;
}

int graph_ready(const gvc_ctx *c) {
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (c->v_begin != 0 || c->v_end != c->n_global) return fail(GVC_ERR_STATE, "whole-graph context required");
    return 0;
}

}  // namespace

extern "C" {

int gvc_trainer_create(gvc_ctx *c, int n_layers, const int *kinds, const int *rows, const int *cols,
                       const float *const *W, const float *const *bias, gvc_trainer **out) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!out) return fail(GVC_ERR_ARG, "out is null");
    *out = nullptr;
    if (n_layers <= 0 || !kinds) return fail(GVC_ERR_ARG, "empty model");
    // width of the model's input: what the first linear layer expects, walked back through the graph layers
    // before it (each maps w to 2 w + 3); a model without linear layers takes width 1
    int in_w = 1;
    for (int i = 0; i < n_layers; ++i) {
        if (kinds[i] < GVC_LINEAR || kinds[i] > GVC_SIGMOID) return fail(GVC_ERR_ARG, "layer %d: unknown kind %d", i, kinds[i]);
        if (kinds[i] != GVC_LINEAR) continue;
        if (!rows || !cols || !W || !bias || !W[i] || !bias[i] || rows[i] <= 0 || cols[i] <= 0)
            return fail(GVC_ERR_ARG, "layer %d: linear layer needs rows, cols, W and bias", i);
        in_w = rows[i];
        for (int j = i - 1; j >= 0; --j)
            if (kinds[j] == GVC_GRAPH) {
                if (in_w < 5 || (in_w - 3) % 2) return fail(GVC_ERR_ARG, "layer %d: width %d cannot come out of a graph layer", i, rows[i]);
                in_w = (in_w - 3) / 2;
            }
        break;
    }
    int w = in_w, maxw = in_w;
    size_t floats = 0;
    for (int i = 0; i < n_layers; ++i) {
        if (kinds[i] < GVC_LINEAR || kinds[i] > GVC_SIGMOID) return fail(GVC_ERR_ARG, "layer %d: unknown kind %d", i, kinds[i]);
        if (kinds[i] == GVC_LINEAR) {
            if (!rows || !cols || !W || !bias || !W[i] || !bias[i] || rows[i] <= 0 || cols[i] <= 0)
                return fail(GVC_ERR_ARG, "layer %d: linear layer needs rows, cols, W and bias", i);
            if (rows[i] > 35 || cols[i] > 32)
                return fail(GVC_ERR_UNSUPPORTED, "layer %d: %d x %d; the training kernels take up to 35 x 32", i, rows[i], cols[i]);
            if (rows[i] != w) return fail(GVC_ERR_ARG, "layer %d expects width %d, got %d", i, rows[i], w);
            w = cols[i];
            floats += (size_t)(rows[i] + 1) * cols[i];
        } else if (kinds[i] == GVC_GRAPH) {
            w = 2 * w + 3;
        }
        maxw = std::max(maxw, w);
    }
    if ((rc = use_device(c))) return rc;
    gvc_trainer *t = new (std::nothrow) gvc_trainer();
    if (!t) return fail(GVC_ERR_ALLOC, "out of host memory");
    t->c = c;
    t->n_params = floats;
    t->out_w = w;
    t->max_w = maxw;
    cudaError_t e = cudaMalloc(&t->params, std::max<size_t>(1, 3 * floats) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&t->partial, (size_t)kTrainParts * 36 * 32 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&t->loss_partial, 1024 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(t->params, 0, std::max<size_t>(1, 3 * floats) * sizeof(float), c->stream);   // grad = vel = 0, :8
    if (e != cudaSuccess) {
        if (t->params) cudaFree(t->params);
        if (t->partial) cudaFree(t->partial);
        delete t;
        return fail(GVC_ERR_ALLOC, "trainer buffers: %s", cudaGetErrorString(e));
    }
    t->L.resize(n_layers);
    size_t at = 0;
    for (int i = 0; i < n_layers; ++i) {
        gvc_trainer::Layer &l = t->L[i];
        l.kind = kinds[i];
        if (l.kind != GVC_LINEAR) continue;
        l.rows = rows[i]; l.cols = cols[i];
        const size_t wn = (size_t)l.rows * l.cols;
        l.W = t->params + at; l.b = l.W + wn;
        l.gW = l.W + floats; l.gb = l.b + floats;
        l.vW = l.W + 2 * floats; l.vb = l.b + 2 * floats;
        at += wn + l.cols;
        cudaMemcpyAsync(l.W, W[i], wn * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(l.b, bias[i], (size_t)l.cols * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    }
    t->in_w = in_w;
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { gvc_trainer_destroy(t); return fail(1000 + (int)e, "trainer upload: %s", cudaGetErrorString(e)); }
    *out = t;
    return 0;
}

void gvc_trainer_destroy(gvc_trainer *t) {
    if (!t) return;
    if (t->c) { cudaSetDevice(t->c->device); cudaStreamSynchronize(t->c->stream); }
    for (auto &l : t->L) if (l.saved) cudaFree(l.saved);
    for (int k = 0; k < 2; ++k) { if (t->act[k]) cudaFree(t->act[k]); if (t->grad[k]) cudaFree(t->grad[k]); }
    if (t->params) cudaFree(t->params);
    if (t->partial) cudaFree(t->partial);
    if (t->loss_partial) cudaFree(t->loss_partial);
    delete t;
}

int gvc_trainer_input_width(const gvc_trainer *t) { return t ? t->in_w : 0; }
int gvc_trainer_output_width(const gvc_trainer *t) { return t ? t->out_w : 0; }

// model_training::predict :81-96.  x: n x in_w (host), out: n x out_w (host, may be null: the output stays on
// the device for gvc_trainer_mse_backprop).  scales: WEIGHT_SCALE of every graph layer (n_scales == 1: all
// the same).  Every linear / ReLU / sigmoid layer keeps its input (the reference's in_copy).
int gvc_trainer_predict(gvc_trainer *t, const float *x, const float *scales, int n_scales, float *out, int mode) {
    int rc;
    if ((rc = trainer_check(t))) return rc;
    gvc_ctx *c = t->c;
    if ((rc = graph_ready(c))) return rc;
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (n_scales < 1 || !scales) return fail(GVC_ERR_ARG, "no weight scale");
    const uint32_t n = c->n_global;
    t->n_fwd = 0;
    if (!n) return 0;
    if (!x) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    size_t cap0 = t->act_cap, cap1 = t->act_cap;
    if ((rc = train_grow(&t->act[0], &cap0, (size_t)n * t->max_w))) return rc;
    if ((rc = train_grow(&t->act[1], &cap1, (size_t)n * t->max_w))) { t->act_cap = 0; return rc; }
    t->act_cap = std::min(cap0, cap1);
    GVC_CUDA(cudaMemcpyAsync(t->act[0], x, (size_t)n * t->in_w * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const float *cur = t->act[0];
    int which = 1, w = t->in_w, gi = 0;
    const int T = 256;
    for (auto &l : t->L) {
        float *dst = t->act[which];
        int wo = w;
        if (l.kind != GVC_GRAPH) {                        // in_copy :13, :45, :56
            if ((rc = train_grow(&l.saved, &l.saved_cap, (size_t)n * w))) return rc;
            GVC_CUDA(cudaMemcpyAsync(l.saved, cur, (size_t)n * w * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
            l.in_w = w;
        }
        switch (l.kind) {
        case GVC_LINEAR:
            wo = l.cols;
            if (mode == GVC_MODE_EXACT)
                generic_linear_kernel<true><<<blocks_for((uint64_t)n * wo, T), T, 0, c->stream>>>(cur, l.rows, l.cols, l.W, l.b, dst, n);
            else
                generic_linear_kernel<false><<<blocks_for((uint64_t)n * wo, T), T, 0, c->stream>>>(cur, l.rows, l.cols, l.W, l.b, dst, n);
            break;
        case GVC_GRAPH: {
            wo = 2 * w + 3;
            const float s = scales[std::min(gi, n_scales - 1)];
            ++gi;
            l.in_w = w;
            generic_graph_kernel<<<blocks_for((uint64_t)n * wo, T), T, 0, c->stream>>>(c->row_ptr, c->col, c->Wv, c->NWv, cur, w, dst, n, 0, s);
            break;
        }
        case GVC_RELU:
            generic_relu_kernel<<<blocks_for((uint64_t)n * w, T), T, 0, c->stream>>>(cur, dst, (uint64_t)n * w);
            break;
        default:
            if (mode == GVC_MODE_EXACT)
                generic_sigmoid_kernel<true><<<blocks_for((uint64_t)n * w, T), T, 0, c->stream>>>(cur, dst, (uint64_t)n * w);
            else
                generic_sigmoid_kernel<false><<<blocks_for((uint64_t)n * w, T), T, 0, c->stream>>>(cur, dst, (uint64_t)n * w);
            break;
        }
        GVC_CUDA(cudaGetLastError());
        c->launches++;
        cur = dst;
        which ^= 1;
        w = wo;
    }
    t->d_out = const_cast<float *>(cur);
    if (out) GVC_CUDA(cudaMemcpyAsync(out, cur, (size_t)n * w * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    t->n_fwd = n;
    return 0;
}

namespace {
// model_training::backprop :98-129 with the gradient of the output already in t->grad[0]
int backprop_device(gvc_trainer *t, float *grad_out_host, int mode) {
    gvc_ctx *c = t->c;
    const uint32_t n = t->n_fwd;
    const int T = 256;
    int which = 0;
    for (int i = (int)t->L.size() - 1; i >= 0; --i) {
        gvc_trainer::Layer &l = t->L[i];
        const float *g = t->grad[which];
        float *dst = t->grad[which ^ 1];
        const bool first = i == 0;
        int rc;
        switch (l.kind) {
        case GVC_LINEAR:
            // the gradient of the model's input is only computed when somebody asked for it
            if ((rc = linear_backward_device(c, n, l.rows, l.cols, l.saved, g, l.W, l.gW, l.gb, (first && !grad_out_host) ? nullptr : dst,
                                             t->partial, mode)))
                return rc;
            break;
        case GVC_GRAPH:
            if (first && !grad_out_host) break;
            train_graph_backward_kernel<<<blocks_for((uint64_t)n * l.in_w, T), T, 0, c->stream>>>(c->row_ptr, c->col, g, l.in_w, dst, n);
            c->launches++;
            break;
        case GVC_RELU:
            train_relu_backward_kernel<<<blocks_for((uint64_t)n * l.in_w, T), T, 0, c->stream>>>(l.saved, g, dst, (uint64_t)n * l.in_w);
            c->launches++;
            break;
        default:
            if (mode == GVC_MODE_EXACT)
                train_sigmoid_backward_kernel<true><<<blocks_for((uint64_t)n * l.in_w, T), T, 0, c->stream>>>(l.saved, g, dst, (uint64_t)n * l.in_w);
            else
                train_sigmoid_backward_kernel<false><<<blocks_for((uint64_t)n * l.in_w, T), T, 0, c->stream>>>(l.saved, g, dst, (uint64_t)n * l.in_w);
            c->launches++;
            break;
        }
        GVC_CUDA(cudaGetLastError());
        which ^= 1;
    }
    if (grad_out_host)
        GVC_CUDA(cudaMemcpyAsync(grad_out_host, t->grad[which], (size_t)n * t->in_w * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int backprop_prepare(gvc_trainer *t, int mode) {
    int rc;
    if ((rc = trainer_check(t))) return rc;
    if ((rc = graph_ready(t->c))) return rc;
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (!t->n_fwd || t->n_fwd != t->c->n_global) return fail(GVC_ERR_STATE, "backprop without a predict on the current graph");
    if ((rc = use_device(t->c))) return rc;
    size_t cap0 = t->grad_cap, cap1 = t->grad_cap;
    if ((rc = train_grow(&t->grad[0], &cap0, (size_t)t->n_fwd * t->max_w))) return rc;
    if ((rc = train_grow(&t->grad[1], &cap1, (size_t)t->n_fwd * t->max_w))) { t->grad_cap = 0; return rc; }
    t->grad_cap = std::min(cap0, cap1);
    return 0;
}
}  // namespace

// grad_in: gradient of the loss w.r.t. the last predict's output (n x out_w); grad_out (may be null): w.r.t.
// its input (n x in_w).  Accumulates into the gradients of every linear layer.
int gvc_trainer_backprop(gvc_trainer *t, const float *grad_in, float *grad_out, int mode) {
    int rc;
    if (t && t->c && t->c->have_graph && t->c->n_global == 0) return 0;
    if ((rc = backprop_prepare(t, mode))) return rc;
    if (!grad_in) return fail(GVC_ERR_ARG, "null buffer");
    GVC_CUDA(cudaMemcpyAsync(t->grad[0], grad_in, (size_t)t->n_fwd * t->out_w * sizeof(float), cudaMemcpyHostToDevice, t->c->stream));
    return backprop_device(t, grad_out, mode);
}

// MSE_loss(out, y) + MSE_grad(out, y, grad) + backprop(grad) on the device: only y goes up and the loss comes
// back (what gnn_train.cpp's run_model does per graph, old_files/src/apps/gnn_train.cpp:85-99)
int gvc_trainer_mse_backprop(gvc_trainer *t, const float *y, float *loss, int mode) {
    int rc;
    if (t && t->c && t->c->have_graph && t->c->n_global == 0) { if (loss) *loss = 0.0f; return 0; }
    if ((rc = backprop_prepare(t, mode))) return rc;
    if (!y) return fail(GVC_ERR_ARG, "null buffer");
    gvc_ctx *c = t->c;
    const uint64_t n = t->n_fwd, count = n * t->out_w;
    float *dy = t->grad[1];
    GVC_CUDA(cudaMemcpyAsync(dy, y, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const int parts = (int)std::min<uint64_t>(1024, (n + 255) / 256);
    train_mse_loss_kernel<<<parts, 256, 0, c->stream>>>(t->d_out, dy, n, t->out_w, t->loss_partial);
    train_mse_grad_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(t->d_out, dy, t->grad[0], count, (float)t->out_w);
    GVC_CUDA(cudaGetLastError());
    c->launches += 2;
    std::vector<double> part(parts);
    GVC_CUDA(cudaMemcpyAsync(part.data(), t->loss_partial, parts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if ((rc = backprop_device(t, nullptr, mode))) return rc;          // synchronises
    double s = 0.0;
    for (double v : part) s += v;
    if (loss) *loss = (float)(s / (double)n);
    return 0;
}

int gvc_trainer_sgd_step(gvc_trainer *t, uint64_t batch_size, float lr, float momentum, float weight_decay) {
    int rc;
    if ((rc = trainer_check(t))) return rc;
    if (!batch_size) return fail(GVC_ERR_ARG, "batch size 0");
    if ((rc = use_device(t->c))) return rc;
    gvc_ctx *c = t->c;
    // parameters, gradients and velocities are three equally laid out blocks: one launch for the whole model
    if (t->n_params) {
        train_sgd_kernel<<<blocks_for(t->n_params, 256), 256, 0, c->stream>>>(t->params, t->params + t->n_params, t->params + 2 * t->n_params,
                                                                             (int)t->n_params, (float)batch_size, lr, momentum, weight_decay);
        GVC_CUDA(cudaGetLastError());
        c->launches++;
    }
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int gvc_trainer_zero_grad(gvc_trainer *t) {
    int rc;
    if ((rc = trainer_check(t))) return rc;
    if ((rc = use_device(t->c))) return rc;
    if (t->n_params) GVC_CUDA(cudaMemsetAsync(t->params + t->n_params, 0, t->n_params * sizeof(float), t->c->stream));
    GVC_CUDA(cudaStreamSynchronize(t->c->stream));
    return 0;
}

// what: 0 parameters, 1 gradients, 2 velocities of linear layer `layer` (index among ALL layers)
int gvc_trainer_read(gvc_trainer *t, int what, int layer, float *W, float *bias) {
    int rc;
    if ((rc = trainer_check(t))) return rc;
    if (what < 0 || what > 2 || layer < 0 || layer >= (int)t->L.size() || t->L[layer].kind != GVC_LINEAR)
        return fail(GVC_ERR_ARG, "no such parameter block");
    if ((rc = use_device(t->c))) return rc;
    const gvc_trainer::Layer &l = t->L[layer];
    const size_t off = (size_t)what * t->n_params;
    if (W) GVC_CUDA(cudaMemcpyAsync(W, l.W + off, (size_t)l.rows * l.cols * sizeof(float), cudaMemcpyDeviceToHost, t->c->stream));
    if (bias) GVC_CUDA(cudaMemcpyAsync(bias, l.b + off, (size_t)l.cols * sizeof(float), cudaMemcpyDeviceToHost, t->c->stream));
    GVC_CUDA(cudaStreamSynchronize(t->c->stream));
    return 0;
}

int gvc_trainer_write(gvc_trainer *t, int what, int layer, const float *W, const float *bias) {
    int rc;
    if ((rc = trainer_check(t))) return rc;
    if (what < 0 || what > 2 || layer < 0 || layer >= (int)t->L.size() || t->L[layer].kind != GVC_LINEAR)
        return fail(GVC_ERR_ARG, "no such parameter block");
    if ((rc = use_device(t->c))) return rc;
    const gvc_trainer::Layer &l = t->L[layer];
    const size_t off = (size_t)what * t->n_params;
    if (W) GVC_CUDA(cudaMemcpyAsync(l.W + off, W, (size_t)l.rows * l.cols * sizeof(float), cudaMemcpyHostToDevice, t->c->stream));
    if (bias) GVC_CUDA(cudaMemcpyAsync(l.b + off, bias, (size_t)l.cols * sizeof(float), cudaMemcpyHostToDevice, t->c->stream));
    GVC_CUDA(cudaStreamSynchronize(t->c->stream));
    return 0;
}

// ---- single layers with host buffers (what the drop-in's *_training structs call) ---------------------------

// linear_layer_training::backward :17-26.  in: n x K (the layer's in_copy), grad_in: n x Nout, W: K x Nout;
// grad_W (K x Nout) and grad_bias (Nout) are read, accumulated into and written back; grad_out: n x K.
int gvc_linear_backward_host(gvc_ctx *c, uint64_t n, int K, int Nout, const float *in, const float *grad_in, const float *W,
                             float *grad_W, float *grad_bias, float *grad_out, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (K <= 0 || Nout <= 0 || K > 35 || Nout > 32) return fail(GVC_ERR_UNSUPPORTED, "layer shape %d x %d (up to 35 x 32)", K, Nout);
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (!W || !grad_W || !grad_bias) return fail(GVC_ERR_ARG, "null buffer");
    if (!n) return 0;
    if (!in || !grad_in || !grad_out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    const size_t wn = (size_t)K * Nout, small = 2 * wn + 2 * (size_t)Nout, part = (size_t)kTrainParts * (K + 1) * Nout;
    if ((rc = c->d_ping.reserve(n * K + n * Nout + small + part))) return rc;
    if ((rc = c->d_pong.reserve(n * K))) return rc;
    float *d_in = c->d_ping.p, *d_g = d_in + n * K, *dW = d_g + n * Nout, *dgW = dW + wn, *dgb = dgW + wn, *dpart = dgb + 2 * (size_t)Nout;
    GVC_CUDA(cudaMemcpyAsync(d_in, in, n * K * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(d_g, grad_in, n * Nout * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dW, W, wn * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dgW, grad_W, wn * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dgb, grad_bias, (size_t)Nout * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if ((rc = linear_backward_device(c, n, K, Nout, d_in, d_g, dW, dgW, dgb, c->d_pong.p, dpart, mode))) return rc;
    GVC_CUDA(cudaMemcpyAsync(grad_W, dgW, wn * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaMemcpyAsync(grad_bias, dgb, (size_t)Nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaMemcpyAsync(grad_out, c->d_pong.p, n * K * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// graph_layer_training::backward :32-42 on the context's graph.  grad_in: n x (2 width + 3), grad_out: n x width.
int gvc_graph_backward_host(gvc_ctx *c, const float *grad_in, int width, float *grad_out) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = graph_ready(c))) return rc;
    if (width <= 0) return fail(GVC_ERR_ARG, "width must be positive");
    const size_t n = c->n_global;
    if (!n) return 0;
    if (!grad_in || !grad_out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    Scratch2 s;
    if ((rc = host_io_begin(c, n * (2 * (size_t)width + 3), n * width, grad_in, &s))) return rc;
    train_graph_backward_kernel<<<blocks_for(n * width, 256), 256, 0, c->stream>>>(c->row_ptr, c->col, s.a, width, s.b, (uint32_t)n);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    return host_io_end(c, n * width, grad_out, s);
}

namespace {
int activation_backward_host(gvc_ctx *c, int kind, uint64_t count, const float *z, const float *g, float *out, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (!count) return 0;
    if (!z || !g || !out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    if ((rc = c->d_ping.reserve(2 * count))) return rc;
    if ((rc = c->d_pong.reserve(count))) return rc;
    float *dz = c->d_ping.p, *dg = dz + count;
    GVC_CUDA(cudaMemcpyAsync(dz, z, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dg, g, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (kind == GVC_RELU) train_relu_backward_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(dz, dg, c->d_pong.p, count);
    else if (mode == GVC_MODE_EXACT) train_sigmoid_backward_kernel<true><<<blocks_for(count, 256), 256, 0, c->stream>>>(dz, dg, c->d_pong.p, count);
    else train_sigmoid_backward_kernel<false><<<blocks_for(count, 256), 256, 0, c->stream>>>(dz, dg, c->d_pong.p, count);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    GVC_CUDA(cudaMemcpyAsync(out, c->d_pong.p, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
}  // namespace

int gvc_relu_backward_host(gvc_ctx *c, uint64_t count, const float *z, const float *grad_in, float *grad_out) {
    return activation_backward_host(c, GVC_RELU, count, z, grad_in, grad_out, GVC_MODE_EXACT);
}
int gvc_sigmoid_backward_host(gvc_ctx *c, uint64_t count, const float *z, const float *grad_in, float *grad_out, int mode) {
    return activation_backward_host(c, GVC_SIGMOID, count, z, grad_in, grad_out, mode);
}

// MSE_loss :175-182 and MSE_grad :184-190 for host matrices (n x w); either of loss / grad may be null
int gvc_mse_host(gvc_ctx *c, uint64_t n, int w, const float *x, const float *y, float *loss, float *grad) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (w <= 0) return fail(GVC_ERR_ARG, "width must be positive");
    if (!n) { if (loss) *loss = NAN; return 0; }                  // the reference divides by the height: 0 / 0
    if (!x || !y) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    const uint64_t count = n * (uint64_t)w;
    if ((rc = c->d_ping.reserve(2 * count + 2048))) return rc;
    if ((rc = c->d_pong.reserve(count))) return rc;
    float *dx = c->d_ping.p, *dy = dx + count;
    double *dpart = reinterpret_cast<double *>(c->d_ping.p + ((2 * count + 1) & ~(uint64_t)1));
    GVC_CUDA(cudaMemcpyAsync(dx, x, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dy, y, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const int parts = (int)std::min<uint64_t>(1000, (n + 255) / 256);
    std::vector<double> part(parts);
    if (loss) {
        train_mse_loss_kernel<<<parts, 256, 0, c->stream>>>(dx, dy, n, w, dpart);
        GVC_CUDA(cudaMemcpyAsync(part.data(), dpart, parts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    if (grad) {
        train_mse_grad_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(dx, dy, c->d_pong.p, count, (float)w);
        GVC_CUDA(cudaMemcpyAsync(grad, c->d_pong.p, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    GVC_CUDA(cudaGetLastError());
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    if (loss) {
        double s = 0.0;
        for (double v : part) s += v;
        *loss = (float)(s / (double)n);
    }
    return 0;
}

// SGD_step :192-224 for one host parameter array (param, grad and vel are updated in place)
int gvc_sgd_host(gvc_ctx *c, uint64_t count, float *param, float *grad, float *vel, uint64_t batch_size, float lr,
                 float momentum, float weight_decay) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!count) return 0;
    if (!param || !grad || !vel || !batch_size) return fail(GVC_ERR_ARG, "null buffer or batch size 0");
    if (count > (1u << 30)) return fail(GVC_ERR_UNSUPPORTED, "parameter array too large");
    if ((rc = use_device(c))) return rc;
    if ((rc = c->d_ping.reserve(3 * count))) return rc;
    float *dp = c->d_ping.p, *dg = dp + count, *dv = dg + count;
    GVC_CUDA(cudaMemcpyAsync(dp, param, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dg, grad, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(dv, vel, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    train_sgd_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(dp, dg, dv, (int)count, (float)batch_size, lr, momentum, weight_decay);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    GVC_CUDA(cudaMemcpyAsync(param, dp, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaMemcpyAsync(grad, dg, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaMemcpyAsync(vel, dv, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
