// gvc_train.cuh -- device kernels of the training path: the backward pass of the
// four layer kinds, the loss and the optimiser step of the reference's
// old_files/src/lib/gnn_training.cpp (included by gvc_api.cu; uses its generic forward kernels and its
// OpenBLAS-ordered dot).
//
//   linear   grad_W += in^T . grad_in, grad_bias += column sums, grad_out = grad_in . W^T      (:17-26)
//   graph    grad_out[u] = sum over v in N(u) of grad_in[v][0..w)  +  grad_in[u][w..2w)          (:32-42)
//   ReLU     z >= 0 ? g : 0  (z = the layer's INPUT, zero included)                             (:50-53)
//   sigmoid  f(z) (1 - f(z)) g                                                                   (:61-65)
//   MSE      loss = mean over rows of (sum (x - y)^2 / width); grad = 2 (x - y) / width          (:175-190)
//   SGD      grad += 2 wd w; vel = momentum vel + grad / batch; w -= lr vel   (fused, see below)   (:192-224)
//
// Two modes, as everywhere in libgvc.  EXACT keeps the reference's fp32 operation order wherever it is
// defined by the source (every elementwise op; the sequential neighbour sum of the graph backward) and
// runs the three dot() calls through blas_dot_element, i.e. in the order of the OpenBLAS kernel the
// parity tests pin -- bit-identical gradients, one thread per output element (slow: the sum over the N
// rows of grad_W is sequential).  FAST reduces grad_W / grad_bias over the rows in parallel (per-CTA
// partial sums in a fixed order, double accumulation in the final pass: deterministic, within 1e-5 of
// the exact result) and uses FFMA.
#pragma once

namespace gvc {

// graph_layer_training::backward :32-42 -- one thread per element of grad_out, neighbours in adjacency
// order, then the vertex's own "self" columns (the forward's column quirk is NOT undone there: the
// reference passes the gradient of columns w..2w straight through, and so do we)
__global__ void train_graph_backward_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                            const float *__restrict__ grad_in, int w, float *__restrict__ grad_out,
                                            uint32_t n) {
    const int iw = 2 * w + 3;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * w) return;
    const uint32_t u = (uint32_t)(idx / w);
    const int c = (int)(idx % w);
    float acc = 0.0f;
    for (uint32_t e = row_ptr[u], end = row_ptr[u + 1]; e < end; ++e)
        acc = __fadd_rn(acc, grad_in[(size_t)col[e] * iw + c]);
    grad_out[idx] = __fadd_rn(acc, grad_in[(size_t)u * iw + w + c]);
}

__global__ void train_relu_backward_kernel(const float *__restrict__ z, const float *__restrict__ g,
                                           float *__restrict__ out, uint64_t count) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) out[idx] = z[idx] >= 0.0f ? g[idx] : 0.0f;
}

template <bool EXACT>
__global__ void train_sigmoid_backward_kernel(const float *__restrict__ z, const float *__restrict__ g,
                                              float *__restrict__ out, uint64_t count) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const float f = sigmoid_ref<EXACT>(z[idx]);          // the reference evaluates f(z) twice; same value
    out[idx] = __fmul_rn(__fmul_rn(f, __fsub_rn(1.0f, f)), g[idx]);
}

// grad_out = grad_in . W^T (dot(grad_in, l.W, grad_out, false, true, 0.0f), :25), fast mode
__global__ void train_linear_dx_kernel(const float *__restrict__ g, int K, int Nout, const float *__restrict__ Wm,
                                       float *__restrict__ out, uint64_t n) {
    extern __shared__ float wsm[];                       // W, K x Nout
    for (int i = threadIdx.x; i < K * Nout; i += blockDim.x) wsm[i] = Wm[i];
    __syncthreads();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * (uint64_t)K) return;
    const uint64_t i = idx / K;
    const int k = (int)(idx % K);
    const float *gr = g + i * Nout;
    float acc = 0.0f;
    for (int j = 0; j < Nout; ++j) acc = fmaf(gr[j], wsm[k * Nout + j], acc);
    out[idx] = acc;
}

// grad_W / grad_bias, fast mode, pass 1: CTA b sums rows b, b + gridDim.x, ... in slabs of 32 staged
// through shared memory; thread t owns the output elements t, t + 256, ... of the (K + 1) x Nout block
// (row K = the bias: an input column of ones).  partial[b][(K + 1) * Nout].
constexpr int kTrainSlab = 32;
__global__ void __launch_bounds__(256) train_linear_dw_partial_kernel(const float *__restrict__ in, const float *__restrict__ g,
                                                                      int K, int Nout, uint64_t n, float *__restrict__ partial) {
    extern __shared__ float sm[];
    float *a = sm;                                   // kTrainSlab x (K + 1)
    float *b = sm + kTrainSlab * (K + 1);            // kTrainSlab x Nout
    const int outs = (K + 1) * Nout;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};        // (35 + 1) * 32 / 256 = 4.5 -> at most 5 per thread
    const uint64_t n_slabs = (n + kTrainSlab - 1) / kTrainSlab;
    for (uint64_t s = blockIdx.x; s < n_slabs; s += gridDim.x) {
        const uint64_t r0 = s * kTrainSlab;
        const int rows = (int)min((uint64_t)kTrainSlab, n - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * K; i += blockDim.x) a[(i / K) * (K + 1) + i % K] = in[r0 * K + i];
        for (int i = threadIdx.x; i < rows; i += blockDim.x) a[i * (K + 1) + K] = 1.0f;
        for (int i = threadIdx.x; i < rows * Nout; i += blockDim.x) b[i] = g[r0 * Nout + i];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int o = threadIdx.x + 256 * q;
            if (o < outs) {
                const int k = o / Nout, j = o % Nout;
                float v = acc[q];
                for (int r = 0; r < rows; ++r) v = fmaf(a[r * (K + 1) + k], b[r * Nout + j], v);
                acc[q] = v;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int o = threadIdx.x + 256 * q;
        if (o < outs) partial[(size_t)blockIdx.x * outs + o] = acc[q];
    }
}
// pass 2: the partial sums in CTA order, in double; added to grad_W (K x Nout) and grad_bias (Nout)
__global__ void train_linear_dw_final_kernel(const float *__restrict__ partial, int n_parts, int K, int Nout,
                                             float *__restrict__ grad_W, float *__restrict__ grad_b) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x, outs = (K + 1) * Nout;
    if (o >= outs) return;
    double s = 0.0;
    for (int p = 0; p < n_parts; ++p) s += (double)partial[(size_t)p * outs + o];
    float *dst = o < K * Nout ? grad_W + o : grad_b + (o - K * Nout);
    *dst = (float)((double)*dst + s);
}

// grad_bias in the reference's order (:21-22): rows ascending, one sequential fp32 sum per column (exact mode)
__global__ void train_bias_grad_exact_kernel(const float *__restrict__ g, int Nout, uint64_t n, float *__restrict__ grad_b) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Nout) return;
    float acc = grad_b[j];
    for (uint64_t i = 0; i < n; ++i) acc = __fadd_rn(acc, g[i * Nout + j]);
    grad_b[j] = acc;
}

// MSE_grad :184-190 (elementwise, exact in both modes): (2 (x - y)) / width
__global__ void train_mse_grad_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ grad,
                                      uint64_t count, float width) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) grad[idx] = __fdiv_rn(__fmul_rn(2.0f, __fsub_rn(x[idx], y[idx])), width);
}
// MSE_loss :175-182: per row the squared errors summed in column order and divided by the width (fp32, as the
// reference); the sum over the rows is a sequential fp32 chain there -- here per-CTA sums in double, combined in
// CTA order by the caller (closer to the true mean than the reference's own chain; compared with a tolerance)
__global__ void __launch_bounds__(256) train_mse_loss_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                             uint64_t n, int w, double *__restrict__ partial) {
    __shared__ double red[256];
    double s = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float se = 0.0f;
        for (int j = 0; j < w; ++j) {
            const float d = __fsub_rn(x[i * w + j], y[i * w + j]);
            se = __fadd_rn(se, __fmul_rn(d, d));
        }
        s += (double)__fdiv_rn(se, (float)w);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// SGD_step :192-224 for one parameter array (weights or bias of one layer).  The three a * b + c
// expressions of the source are evaluated as fused multiply-adds: that is what the reference IS once
// compiled with its own flags (-O3 -march=native, GCC's default -ffp-contract=fast) on any CPU with FMA.
__global__ void train_sgd_kernel(float *__restrict__ param, float *__restrict__ grad, float *__restrict__ vel, int count,
                                 float batch, float lr, float momentum, float weight_decay) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float g = grad[i];
    if (weight_decay > 0.0f) {                                              // :194-206 (the regularised gradient is stored back)
        g = __fmaf_rn(__fmul_rn(2.0f, weight_decay), param[i], g);
        grad[i] = g;
    }
    const float v = __fmaf_rn(momentum, vel[i], __fdiv_rn(g, batch));     // :213-214
    vel[i] = v;
    param[i] = __fmaf_rn(-lr, v, param[i]);                                // :217-218
}

}  // namespace gvc
