// gvc_api.cu -- the C ABI of include/gvc.h on top of the sm_100a kernels.
//
// Host-side counterpart of gnn::model::predict (reference
// src/gnn_inference.cpp:67-81): owns the device copies of the model and of the
// graph's CSR view, picks the fused three-stage path for the GNN_VC
// architecture or the per-layer path for any other layer
// sequence, and enqueues everything on one stream.  No CPU fallback exists: if
// CUDA is unusable every entry point reports an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gvc.h"
#include "gvc_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define GVC_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(1000 + (int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                \
    } while (0)

struct HostLayer {
    int kind = 0, rows = 0, cols = 0;
    std::vector<float> W, b;
    float *dW = nullptr, *db = nullptr;
};

// Device memory of a context comes out of a few big chunks (one cudaMalloc per chunk, bump
// allocation, nothing is returned before the context goes): a first predict() then pays for ONE
// allocation sized for its graph instead of twenty (cudaMalloc costs up to a millisecond each, and
// GNN_VC's first graph is its largest).  Buffers that outgrow their place move to a new one; the
// old place stays unused until the context is destroyed (graphs shrink in a solver run).
struct Arena {
    struct Chunk { char *p; size_t size, used; };
    std::vector<Chunk> chunks;
    size_t next_chunk = 0;                 // hint: size of the next chunk (set before a burst of reservations)
    void *alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!chunks.empty()) {
            Chunk &c = chunks.back();
            if (c.size - c.used >= bytes) { void *r = c.p + c.used; c.used += bytes; return r; }
        }
        size_t want = std::max(std::max(bytes, next_chunk), (size_t)8 << 20);    // small requests share a chunk
        next_chunk = 0;
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && want > bytes) { cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) { fail(GVC_ERR_ALLOC, "cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e)); return nullptr; }
        chunks.push_back({static_cast<char *>(p), want, bytes});
        return p;
    }
    void release_all() {
        for (auto &c : chunks) cudaFree(c.p);
        chunks.clear();
    }
};
thread_local Arena *tl_arena = nullptr;    // the arena of the context an entry point works on (use_device)

template <typename T>
struct DevBuf {   // grow-only device buffer
    T *p = nullptr;
    size_t cap = 0;
    bool own = false;                      // allocated with cudaMalloc (no arena at hand), freed on release
    int reserve(size_t n) {
        if (n <= cap) return 0;
        release();
        size_t want = n + n / 8 + 64;
        if (tl_arena) {
            p = static_cast<T *>(tl_arena->alloc(want * sizeof(T)));
            if (!p) return GVC_ERR_ALLOC;
        } else {
            cudaError_t e = cudaMalloc(&p, want * sizeof(T));
            if (e != cudaSuccess) { p = nullptr; return fail(GVC_ERR_ALLOC, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); }
            own = true;
        }
        cap = want;
        return 0;
    }
    void release() { if (p && own) cudaFree(p); p = nullptr; cap = 0; own = false; }
};

template <typename T>
struct PinBuf {   // grow-only pinned host buffer
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
        if (e != cudaSuccess) return fail(GVC_ERR_ALLOC, "cudaMallocHost(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// GVC_TRACE=1: wall time of the steps of an upload on stderr (synchronises at every step)
struct Tracer {
    bool on;
    std::chrono::steady_clock::time_point t;
    Tracer() : on(std::getenv("GVC_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void tick(const char *what) {
        if (!on) return;
        cudaDeviceSynchronize();
        auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "gvc trace: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

}  // namespace

struct gvc_ctx {
    Arena arena;                         // all DevBufs below live in it
    int device = 0;
    cudaStream_t stream = nullptr;       // where work is enqueued (own_stream unless gvc_set_stream)
    cudaStream_t own_stream = nullptr;
    uint64_t launches = 0;

    // model
    std::vector<HostLayer> layers;
    bool fused = false;
    float *d_stage_params[3] = {nullptr, nullptr, nullptr};
    float *d_stage_params_fast[3] = {nullptr, nullptr, nullptr};     // fragment-ordered TF32 hi/lo weights (fast mode on the tensor cores)
    std::vector<float> layer_scales;     // gvc_model_weight_scales: WEIGHT_SCALE of every graph layer (empty: the call's scalar)

    // graph shard
    bool have_graph = false;
    uint32_t n_global = 0, v_begin = 0, v_end = 0;
    int tail_override = -1;          // -1: default rule, 0: no tail vertex here, 1: tail_local
    uint32_t tail_local = 0;
    uint64_t nnz = 0;
    const uint32_t *row_ptr = nullptr, *col = nullptr, *Wv = nullptr, *NWv = nullptr;   // device views (the caller's vertex ids)
    // what the FUSED stage kernels and their schedule read: the same arrays, or -- large whole-graph contexts --
    // a copy of the graph with its vertices renumbered in the order of the degree schedule (relabel_rows)
    const uint32_t *f_row_ptr = nullptr, *f_col = nullptr, *f_W = nullptr, *f_NW = nullptr;
    bool relabelled = false;
    size_t l2_window = (size_t)-1;       // persisting L2 window of the stage launches, bytes (-1: not decided yet)
    uint32_t forwards_on_graph = 0;      // fused forwards since the last graph upload (relabel_rows pays off from the second on)
    DevBuf<uint32_t> d_row_order, d_pos_of, r_row_ptr, r_col, r_W, r_NW;     // internal row r holds vertex d_row_order[r]
    DevBuf<float> d_xp, d_sp, d_keys_p;                                       // x, scores, keys in row order
    DevBuf<uint8_t> d_side_p;
    DevBuf<uint32_t> own_row_ptr, own_col, own_W, own_NW;
    DevBuf<uint64_t> d_row_ptr64;        // the ABI's 64-bit offsets, narrowed and checked on the device
    DevBuf<uint32_t> d_flag;             // upload validation result (kBadRowPtr | kBadCol)
    PinBuf<uint64_t> stg_row_ptr;        // gvc_graph_staging: pinned host buffers the caller fills
    PinBuf<uint32_t> stg_col, stg_W, stg_NW;
    // gvc_graph_upload_stream: the graph as the host holds it (one edge array with holes + a range per
    // vertex) goes through a small ring of pinned slots, filled by worker threads through the caller's
    // callbacks while earlier slots are in flight; the packed CSR is built on the device
    PinBuf<unsigned char> ring;          // n_slots x slot_bytes
    size_t slot_bytes = 0;
    int n_slots = 0;
    std::vector<cudaEvent_t> slot_ev;    // slot's last copy has left the host buffer
    DevBuf<uint32_t> d_span, d_rb, d_re; // raw edge span, per-vertex [begin, end) into it
    DevBuf<uint64_t> d_blk;              // block sums of the degree scan (+ the total at the end)
    cudaStream_t copy_stream = nullptr;  // the adjacency travels here while the schedule is built on `stream`
    cudaEvent_t ev_begin = nullptr, ev_copy = nullptr, ev_vert = nullptr;
    // schedule (gvc_kernels.cuh): vertices counting-sorted by degree bin + tile classes
    DevBuf<uint32_t> d_order, d_bins, d_sync;
    DevBuf<uint4> d_vrec;
    DevBuf<float> d_feat;
    // fast-mode hub chunks (gvc::HubSplit): [0] the ring vertices (width 16), [1] the stage-0 giants
    // [2]: the chunks of the exact-mode "parallel exact" vertices (gvc_px.cuh), with their scratch rows and batch records
    DevBuf<uint4> d_hub_chunk[3];
    DevBuf<uint2> d_hub_info[3];
    DevBuf<float> d_hub_partial[2];
    DevBuf<uint32_t> d_hub_count;
    DevBuf<float> d_px_scratch, d_px_S, d_px_T, d_px_rec;
    DevBuf<uint32_t> d_px_flag;
    uint32_t px_ctr_off = 0;             // where the PX counters start inside d_sync
    gvc::Schedule sched{};
    gvc::PeerOut peers[2] = {};          // stage 0 / stage 1 outputs mirrored into the other ranks' buffers
    gvc::PartMap parts{};                // gvc_peer_owners: who owns which vertex range (n_parts == 0: not told)
    DevBuf<uint8_t> d_peer_mask;         // per local vertex: the peers that read its row
    bool have_peer_mask = false;
    uint32_t n_live = 0;                 // positions of `order` with a non-empty adjacency list
    int num_sms = 148;
    int ctas_per_sm[3] = {1, 1, 1};      // resident CTAs per SM of each stage kernel (occupancy query)

    // activations
    DevBuf<float> d_x, d_h1, d_h2, d_scores, d_ping, d_pong;
    DevBuf<float> d_keys;                // selection keys of the last forward (gvc_forward_keys)
    DevBuf<uint8_t> d_side;
    uint32_t x_resident_n = 0;           // gvc_graph_upload_stream_x left the input of this many vertices in d_x (gvc_forward with x = NULL)
    uint32_t keys_valid_n = 0;           // d_keys/d_side hold the keys of the last gvc_forward for this many vertices
    float *keys_out = nullptr;           // where stage 2 of the NEXT launch writes them (null: not wanted)
    uint8_t *side_out = nullptr;
    PinBuf<float> pin_x, pin_scores;

    uint32_t n_local() const { return v_end - v_begin; }
};

namespace {

using namespace gvc;

const int kFusedKinds[21] = {GVC_GRAPH, GVC_LINEAR, GVC_RELU, GVC_LINEAR, GVC_RELU, GVC_LINEAR, GVC_RELU,
                             GVC_GRAPH, GVC_LINEAR, GVC_RELU, GVC_LINEAR, GVC_RELU, GVC_LINEAR, GVC_RELU,
                             GVC_GRAPH, GVC_LINEAR, GVC_RELU, GVC_LINEAR, GVC_RELU, GVC_LINEAR, GVC_SIGMOID};
const int kFusedDims[9][2] = {{5, 32}, {32, 32}, {32, 16}, {35, 32}, {32, 32}, {32, 16}, {35, 32}, {32, 16}, {16, 1}};

bool detect_fused(const std::vector<HostLayer> &L) {
    if (L.size() != 21) return false;
    int li = 0;
    for (int i = 0; i < 21; ++i) {
        if (L[i].kind != kFusedKinds[i]) return false;
        if (L[i].kind == GVC_LINEAR) {
            if (L[i].rows != kFusedDims[li][0] || L[i].cols != kFusedDims[li][1]) return false;
            for (float v : L[i].W) if (!std::isfinite(v)) return false;   // dropping rows 32..34 needs finite weights
            for (float v : L[i].b) if (!std::isfinite(v)) return false;
            ++li;
        }
    }
    return true;
}

// Pack one stage's three dense layers as gvc_kernels.cuh's StageDims describes.
std::vector<float> pack_stage(const std::vector<HostLayer> &L, int stage) {
    std::vector<const HostLayer *> lin;
    for (auto &l : L) if (l.kind == GVC_LINEAR) lin.push_back(&l);
    const StageDims D = stage_dims(stage);
    const int K[3] = {D.Ka, D.Kb, D.Kc}, N[3] = {D.Na, D.Nb, D.Nc};
    std::vector<float> out;
    for (int j = 0; j < 3; ++j) {
        const HostLayer *l = lin[stage * 3 + j];
        out.insert(out.end(), l->W.begin(), l->W.begin() + (size_t)K[j] * N[j]);   // first K rows (drops 32..34 of 35)
        out.insert(out.end(), l->b.begin(), l->b.begin() + N[j]);
    }
    return out;
}

// The fast mode's parameter block of one stage: weights in mma fragment order, split into TF32 hi / lo
// (gvc_kernels.cuh stage_mma_floats).
float tf32_rna(float x) {                       // cvt.rna.tf32.f32: 10 mantissa bits, ties away from zero
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}
std::vector<float> pack_stage_mma(const std::vector<HostLayer> &L, int stage) {
    std::vector<const HostLayer *> lin;
    for (auto &l : L) if (l.kind == GVC_LINEAR) lin.push_back(&l);
    const StageDims D = stage_dims(stage);
    const int K[3] = {D.Ka, D.Kb, D.Kc}, N[3] = {D.Na, D.Nb, D.Nc};
    std::vector<float> out;
    for (int j = 0; j < 3; ++j) {
        const HostLayer *l = lin[stage * 3 + j];
        auto w = [&](int k, int n) { return k < K[j] ? l->W[(size_t)k * N[j] + n] : 0.0f; };   // rows 32..34 of the 35-row matrices are dead
        if (N[j] >= 8) {
            const int KB = (K[j] + 7) / 8, NB = N[j] / 8;
            for (int kb = 0; kb < KB; ++kb)
                for (int nb = 0; nb < NB; ++nb) {
                    float blk[128];
                    for (int lane = 0; lane < 32; ++lane) {
                        const int kk = 8 * kb + lane % 4, nn = 8 * nb + lane / 4;
                        const float b0 = w(kk, nn), b1 = w(kk + 4, nn);
                        const float h0 = tf32_rna(b0), h1 = tf32_rna(b1);
                        blk[lane] = h0; blk[32 + lane] = h1;
                        blk[64 + lane] = tf32_rna(b0 - h0); blk[96 + lane] = tf32_rna(b1 - h1);
                    }
                    out.insert(out.end(), blk, blk + 128);
                }
        } else {
            out.insert(out.end(), l->W.begin(), l->W.begin() + (size_t)K[j] * N[j]);
        }
        out.insert(out.end(), l->b.begin(), l->b.begin() + N[j]);
    }
    return out;
}

// ---- generic per-layer kernels (any layer sequence) ------------------------------

// graph_layer::forward :27-42, one thread per output element
__global__ void generic_graph_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                     const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                     const float *__restrict__ in, int w, float *__restrict__ out,
                                     uint32_t n_local, uint32_t v_begin, float scale) {
    const int ow = 2 * w + 3;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_local * ow) return;
    const uint32_t u = (uint32_t)(idx / ow);
    const int c = (int)(idx % ow);
    const uint32_t beg = row_ptr[u], end = row_ptr[u + 1];
    float v = 0.0f;
    if (c < w) {
        for (uint32_t e = beg; e < end; ++e) v = __fadd_rn(v, in[(size_t)col[e] * w + c]);
    } else if (c < 2 * w) {
        v = in[(size_t)(v_begin + u) * w + (c - w)];
    }
    if (c == w + 1) v = __uint2float_rn(end - beg);
    if (c == w + 2) v = __fdiv_rn(__uint2float_rn(Wv[u]), scale);
    if (c == w + 3) v = __fdiv_rn(__uint2float_rn(NWv[u]), scale);
    out[idx] = v;
}

// ---- dot(), src/matrix.cpp:106-122: the cblas_sgemm call in OpenBLAS' operation order ------------
// OpenBLAS 0.3.15 "Prescott" (8x4 SSE3 micro-kernel, no FMA), one thread, in row-major terms
// (oracle/gnn_oracle.c holds the derivation): rows go 4 at a time, then 2, then 1; columns 8, 4, 2, 1;
// the (row class, column class) pair fixes how the k sum of an element is split over accumulators; k is
// cut into blocks of 128 (a rest of 129..255 is halved), every block's sum is added to C in turn.
enum { kSumSeq = 0, kSumTwo = 1, kSumFour = 2, kSumEight = 3 };

__device__ __forceinline__ int blas_row_class(uint64_t i, uint64_t m) {
    const uint64_t m4 = m & ~3ull;
    if (i < m4) return 0;
    if ((m & 2) && i < m4 + 2) return 1;
    return 2;
}
__device__ __forceinline__ int blas_col_class(uint64_t j, uint64_t n) {
    uint64_t p = n & ~7ull;
    if (j < p) return 0;
    if (n & 4) { if (j < p + 4) return 1; p += 4; }
    if (n & 2) { if (j < p + 2) return 2; p += 2; }
    return 3;
}
__device__ __forceinline__ int blas_sum_scheme(int rc, int cc) {
    // rows: 4 | 2 | 1 ; cols: 8 4 2 1
    const int t = rc == 0 ? 0x0011 : rc == 1 ? 0x0111 : 0x1132;      // nibbles, column class 0 (8 cols) is the top one
    return (t >> (4 * (3 - cc))) & 0xF;
}

// one k block [lo, hi) of one element: a[k * sa] * b[k * sb]
__device__ __forceinline__ float blas_block_sum(const float *__restrict__ a, uint64_t sa, const float *__restrict__ b,
                                                uint64_t sb, uint64_t lo, uint64_t hi, int scheme) {
    const uint64_t K = hi - lo;
    a += lo * sa; b += lo * sb;
    uint64_t k = 0;
    if (scheme == kSumSeq) {
        float acc = 0.0f;
        for (; k < K; ++k) acc = __fadd_rn(acc, __fmul_rn(a[k * sa], b[k * sb]));
        return acc;
    }
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f, c5 = 0.f, c6 = 0.f, c7 = 0.f;
    auto mac = [&](float &c, uint64_t kk) { c = __fadd_rn(c, __fmul_rn(a[kk * sa], b[kk * sb])); };
    if (scheme == kSumTwo) {
        const uint64_t body = K / 8 * 8;
        for (; k < body; k += 2) { mac(c0, k); mac(c1, k + 1); }
        for (; k < K; ++k) mac(c0, k);
        return __fadd_rn(c0, c1);
    }
    if (scheme == kSumFour) {
        const uint64_t body = K / 8 * 8;
        for (; k < body; k += 4) { mac(c0, k); mac(c1, k + 1); mac(c2, k + 2); mac(c3, k + 3); }
        for (; k < K; ++k) mac(c0, k);
        return __fadd_rn(__fadd_rn(c0, c1), __fadd_rn(c2, c3));
    }
    const uint64_t body = K / 16 * 16;
    for (; k < body; k += 8) {
        mac(c0, k); mac(c1, k + 1); mac(c2, k + 2); mac(c3, k + 3);
        mac(c4, k + 4); mac(c5, k + 5); mac(c6, k + 6); mac(c7, k + 7);
    }
    for (; k < K; ++k) mac(c0, k);
    return __fadd_rn(__fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c4, c6)), __fadd_rn(__fadd_rn(c1, c3), __fadd_rn(c5, c7)));
}

// element (i, j) of op(A) op(B) + c_in over all k blocks; c_in = beta * C (0 for beta == 0)
__device__ __forceinline__ float blas_dot_element(const float *__restrict__ a, uint64_t sa, const float *__restrict__ b,
                                                  uint64_t sb, uint64_t k, int scheme, float c_in) {
    float c = c_in;
    uint64_t lo = 0;
    while (lo < k) {
        uint64_t len = k - lo;
        if (len >= 256) len = 128;
        else if (len > 128) len = (len / 2 + 7) / 8 * 8;
        c = __fadd_rn(c, blas_block_sum(a, sa, b, sb, lo, lo + len, scheme));
        lo += len;
    }
    return c;
}

// linear_layer::forward :20-25: dot(in, W, out) with beta = 0, then the row-wise bias add
template <bool EXACT>
__global__ void generic_linear_kernel(const float *__restrict__ in, int K, int Nout,
                                      const float *__restrict__ Wm, const float *__restrict__ bias,
                                      float *__restrict__ out, uint64_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * (uint64_t)Nout) return;
    const uint64_t i = idx / Nout;
    const int j = (int)(idx % Nout);
    const float *a = in + i * K;
    float r;
    if (!EXACT) {
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) acc = fmaf(a[k], Wm[(size_t)k * Nout + j], acc);
        r = acc;
    } else {
        r = blas_dot_element(a, 1, Wm + j, (uint64_t)Nout, (uint64_t)K,
                             blas_sum_scheme(blas_row_class(i, n), blas_col_class((uint64_t)j, (uint64_t)Nout)), 0.0f);
    }
    out[idx] = __fadd_rn(r, bias[j]);
}

__global__ void generic_relu_kernel(const float *__restrict__ in, float *__restrict__ out, uint64_t count) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) out[idx] = relu_ref(in[idx]);
}

template <bool EXACT>
__global__ void generic_sigmoid_kernel(const float *__restrict__ in, float *__restrict__ out, uint64_t count) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) out[idx] = sigmoid_ref<EXACT>(in[idx]);
}

// dot(), src/matrix.cpp:106-122: C = op(A) op(B) + beta C, one thread per element of C, in the
// reference's (OpenBLAS') summation order for every shape, transpose and beta
__global__ void generic_sgemm_kernel(int ta, int tb, uint64_t m, uint64_t n, uint64_t k,
                                     const float *__restrict__ A, uint64_t lda, const float *__restrict__ B,
                                     uint64_t ldb, float beta, float *__restrict__ Cm, uint64_t ldc) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * n) return;
    const uint64_t i = idx / n, j = idx % n;
    const float *a = ta ? A + i : A + i * lda;
    const float *b = tb ? B + j * ldb : B + j;
    float *c = Cm + i * ldc + j;
    const float c_in = beta == 0.0f ? 0.0f : __fmul_rn(beta, *c);
    *c = blas_dot_element(a, ta ? lda : 1, b, tb ? 1 : ldb, k, blas_sum_scheme(blas_row_class(i, m), blas_col_class(j, n)), c_in);
}

inline unsigned blocks_for(uint64_t work, int threads) { return (unsigned)((work + threads - 1) / threads); }

// size of the persisting L2 window of the stage launches (see launch_stage): 48 MB, the device's limits permitting
size_t l2_window_bytes(gvc_ctx *c) {
    if (c->l2_window != (size_t)-1) return c->l2_window;
    const char *e = std::getenv("GVC_L2_WINDOW_MB");
    size_t want = e ? (size_t)std::strtoull(e, nullptr, 10) << 20 : (size_t)48 << 20;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
    want = std::min(want, std::min((size_t)std::max(max_persist, 0), (size_t)std::max(max_window, 0)));
    if (want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); want = 0; }
    c->l2_window = want;
    return want;
}

template <int STAGE>
int launch_stage(gvc_ctx *c, const float *d_in, float *d_out, float scale, int mode) {
    const uint32_t nl = c->n_local();
    if (nl == 0) return 0;
    const Schedule &sc = c->sched;
    const bool exact = mode == GVC_MODE_EXACT;
    const uint32_t n_tasks = STAGE == 0 ? (exact ? sc.n_giant1 : sc.n_chunks1) + (nl - sc.n_giant1 + kTileVerts - 1) / kTileVerts + sc.n_feat_tiles
                                        : (exact ? sc.n_px : 0u) + (sc.n_mid + 7) / 8 + sc.n_tiles + sc.n_feat_tiles;
    const unsigned want = std::max<unsigned>(STAGE == 0 ? (exact ? sc.n_chunks_px : 0u)
                                                        : (exact ? std::max(sc.n_ring - sc.n_px, sc.n_chunks_px) : sc.n_chunks16),
                                             (n_tasks + kWarpsPerCta - 1) / kWarpsPerCta);
    // persistent kernel: never more CTAs than can be resident at once (warps wait on each other's
    // feature vectors; a CTA that is not running could never deliver its ring tasks)
    const unsigned grid = std::max(1u, std::min<unsigned>(c->ctas_per_sm[STAGE] * c->num_sms, want));
    Schedule sc_launch = sc;
    sc_launch.n_ring_ctas = std::min<unsigned>(grid, (unsigned)c->num_sms);   // one ring CTA per SM at most
    const size_t smem = exact ? stage_smem_bytes<STAGE, true>() : stage_smem_bytes<STAGE, false>();
    // task counter + per-feature-tile completion counters start at zero
    PeerOut peers{};
    if (STAGE < 2) peers = c->peers[STAGE];
    peers.n_live = c->n_live;
    peers.keys = STAGE == 2 ? (c->relabelled && c->keys_out ? c->d_keys_p.p : c->keys_out) : nullptr;
    peers.side = STAGE == 2 ? (c->relabelled && c->keys_out ? c->d_side_p.p : c->side_out) : nullptr;
    peers.mask = c->have_peer_mask ? c->d_peer_mask.p - c->v_begin : nullptr;
    const int hk = STAGE == 0 ? 1 : 0;
    const uint32_t n_split = STAGE == 0 ? sc.n_giant1 : sc.n_ring;
    HubSplit hub{c->d_hub_chunk[hk].p, c->d_hub_info[hk].p, c->d_hub_partial[hk].p,
                 c->d_sync.p + kSyncCounters + sc.n_feat_tiles};
    // counters start at zero: task claims, feature tiles, then (fast) chunks done per split vertex or (exact) the
    // claim and completion counters of the parallel exact sums
    const size_t n_counters = exact ? (sc.n_px ? (size_t)c->px_ctr_off + 8 + 3 * (size_t)sc.n_px : kSyncCounters + (size_t)sc.n_feat_tiles)
                                    : kSyncCounters + (size_t)sc.n_feat_tiles + n_split;
    GVC_CUDA(cudaMemsetAsync(c->d_sync.p, 0, n_counters * sizeof(uint32_t), c->stream));
    PxArgs px{c->d_hub_chunk[2].p, c->d_hub_info[2].p, c->d_px_scratch.p, c->d_px_S.p, c->d_px_T.p, c->d_px_rec.p, c->d_px_flag.p,
              c->d_sync.p + c->px_ctr_off, sc.n_chunks_px, sc.n_px};
    // Cooperative launch: warps of this persistent kernel wait for feature vectors that other CTAs
    // produce, so every CTA of the grid must be resident at once.  The grid is sized for that (above);
    // the launch attribute makes the driver guarantee it even when the stream shares the GPU with other
    // work (a framework stream, MPS, concurrent kernels) instead of relying on dispatch order.
    // Rows in schedule order put the rows of the high-degree vertices at the FRONT of h: on graphs whose rows do
    // not fit L2 that front is pinned there for the two wide stages (persisting access window on the launching
    // stream), so that the id stream and the cold rows do not evict it.  Measured on R-MAT scale 23 (same box,
    // ms per stage): no window 1.77 / 4.24 / 4.00; 48 MB 1.78 / 4.14 / 3.90; 64 MB 1.84 / 4.10 / 3.90; 80 MB 1.92 /
    // 4.09 / 3.96 -- the set-aside is taken from everybody else's L2, stage 0 included, so more is not better.
    // GVC_L2_WINDOW_MB overrides the size (0 = off).
    static const size_t window_min_rows_bytes = [] {                    // GVC_L2_WINDOW_MIN_MB: rows of at least this size get the window
        const char *e = std::getenv("GVC_L2_WINDOW_MIN_MB");
        return e ? (size_t)std::strtoull(e, nullptr, 10) << 20 : (size_t)96 << 20;
    }();
    const bool wants_window = STAGE >= 1 && c->relabelled && (size_t)c->n_global * 64 > window_min_rows_bytes;
    const size_t l2_window = wants_window ? l2_window_bytes(c) : 0;     // the device limit is only touched for such graphs
    const bool windowed = l2_window != 0;
    if (windowed) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = const_cast<float *>(d_in);
        av.accessPolicyWindow.num_bytes = std::min(l2_window, (size_t)c->n_global * 64);
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();   // a hint only
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kCtaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute coop{};
    coop.id = cudaLaunchAttributeCooperative;
    coop.val.cooperative = 1;
    cfg.attrs = &coop;
    cfg.numAttrs = 1;
    const uint32_t *a_rp = c->f_row_ptr, *a_col = c->f_col, *a_w = c->f_W, *a_nw = c->f_NW, *a_order = c->d_order.p;
    const uint4 *a_vrec = c->d_vrec.p;
    float *a_feat = c->d_feat.p;
    uint32_t *a_sync = c->d_sync.p;
    const float *a_params = (exact || !GVC_FAST_MMA) ? c->d_stage_params[STAGE] : c->d_stage_params_fast[STAGE];
    const uint32_t a_vb = c->v_begin;
    if (mode == GVC_MODE_EXACT) {
        GVC_CUDA(cudaLaunchKernelEx(&cfg, stage_kernel<STAGE, true>, a_rp, a_col, a_w, a_nw, a_order, a_vrec, sc_launch, hub, px, peers,
                                    a_feat, a_sync, d_in, d_out, a_params, a_vb, scale));
    } else {
        GVC_CUDA(cudaLaunchKernelEx(&cfg, stage_kernel<STAGE, false>, a_rp, a_col, a_w, a_nw, a_order, a_vrec, sc_launch, hub, px, peers,
                                    a_feat, a_sync, d_in, d_out, a_params, a_vb, scale));
    }
    c->launches++;
    if (windowed) {                                  // later work on this stream is none of the window's business
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
    // OpenBLAS' 1-row remainder kernel: last vertex of an odd-sized graph (exact mode only)
    const bool default_tail = (c->n_global & 1u) && c->v_end == c->n_global;
    const bool has_tail = c->tail_override < 0 ? default_tail : c->tail_override == 1;
    if (mode == GVC_MODE_EXACT && has_tail) {
        const uint32_t tail = c->tail_override == 1 ? c->tail_local : nl - 1;
        stage_tail_kernel<STAGE><<<1, 32, 0, c->stream>>>(c->f_row_ptr, c->f_col, c->f_W, c->f_NW, d_in, d_out,
                                                          c->d_stage_params[STAGE], tail, c->v_begin, scale, peers);
        GVC_CUDA(cudaGetLastError());
        c->launches++;
    }
    return 0;
}

// ---- upload-time checks, on the device (the arrays are there anyway) -----------------------
constexpr uint32_t kBadRowPtr = 1u, kBadCol = 2u;

// 64-bit ABI offsets -> the 32-bit layout the kernels read; flags a non-monotone array
__global__ void narrow_row_ptr_kernel(const uint64_t *__restrict__ rp64, uint32_t n_local, uint64_t nnz,
                                      uint32_t *__restrict__ rp32, uint32_t *__restrict__ flag) {
    bool bad = false;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u <= n_local; u += gridDim.x * blockDim.x) {
        const uint64_t v = rp64[u];
        rp32[u] = (uint32_t)v;
        if (v > nnz || (u < n_local && rp64[u + 1] < v)) bad = true;
    }
    if (bad) atomicOr(flag, kBadRowPtr);
}

// every neighbour id must name a vertex of the global graph (the gather would read outside h)
__global__ void check_col_kernel(const uint32_t *__restrict__ col, uint64_t nnz, uint32_t n_global,
                                 uint32_t *__restrict__ flag) {
    bool bad = false;
    const uint64_t n4 = nnz / 4;
    const uint4 *c4 = reinterpret_cast<const uint4 *>(col);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = c4[i];
        bad |= (v.x >= n_global) | (v.y >= n_global) | (v.z >= n_global) | (v.w >= n_global);
    }
    if (blockIdx.x == 0 && threadIdx.x < (nnz & 3)) bad |= col[n4 * 4 + threadIdx.x] >= n_global;
    if (bad) atomicOr(flag, kBadCol);
}

// Per-vertex list of the peers that read its row (gvc_peer_owners), rebuilt with every graph.
int build_peer_mask(gvc_ctx *c) {
    c->have_peer_mask = false;
    const uint32_t nl = c->n_local();
    const PartMap &pm = c->parts;
    if (!pm.n_parts || !nl || pm.bounds[pm.n_parts] != c->n_global) return 0;
    int rc;
    if ((rc = c->d_peer_mask.reserve(nl))) return rc;
    uint32_t all = 0;
    for (int k = 0; k < pm.n_parts; ++k)
        if (pm.peer_of_part[k] >= 0) all |= 1u << pm.peer_of_part[k];
    peer_mask_kernel<<<std::min<unsigned>(1184, (nl + 255) / 256), 256, 0, c->stream>>>(c->row_ptr, c->col, nl, pm, all, c->d_peer_mask.p);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    c->have_peer_mask = true;
    return 0;
}

// Degree schedule of the shard's vertices (see gvc_kernels.cuh).  Part of the graph upload:
// one histogram kernel, a 132-entry scan on the host, one scatter kernel.
int build_schedule(gvc_ctx *c) {
    const uint32_t nl = c->n_local();
    c->sched = Schedule{};
    if (nl == 0) return 0;
    int rc;
    if ((rc = c->d_order.reserve(nl))) return rc;
    if ((rc = c->d_vrec.reserve(nl))) return rc;
    if ((rc = c->d_bins.reserve(kNumDegBins))) return rc;
    GVC_CUDA(cudaMemsetAsync(c->d_bins.p, 0, kNumDegBins * sizeof(uint32_t), c->stream));
    const unsigned grid = std::min<unsigned>(1184, (nl + 255) / 256);
    degree_hist_kernel<<<grid, 256, 0, c->stream>>>(c->f_row_ptr, nl, c->d_bins.p);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    uint32_t hist[kNumDegBins], start[kNumDegBins];
    GVC_CUDA(cudaMemcpyAsync(hist, c->d_bins.p, sizeof(hist), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    uint32_t pos = 0, n_ring = 0, n_pre = 0, n_giant1 = 0, n_px = 0;
    // Which vertices get their sequential sums emulated in parallel (gvc_px.cuh)?  Only those whose
    // chain would otherwise be the critical path of a stage: the emulation gathers a hub's rows twice
    // and keeps a scratch copy, so a 20 000-neighbour vertex of a 258 M-entry shard (chain 0.07 ms in a
    // 4.6 ms stage) is better left to one ring CTA, while the same vertex in a 30 M-entry shard is not.
    // Chain ~3.4 ns per neighbour against ~0.016 ns per adjacency entry of the shard for everything else.
    const uint64_t px_deg = std::max<uint64_t>(kPxMinDeg, c->nnz / kPxNnzPerNeighbour);
    auto bin_min_degree = [](int b) -> uint64_t {                // inverse of degree_bin: smallest degree of bin b
        if (b < 4) return (uint64_t)b;
        const int lg = (b + 4) / 4, frac = (b + 4) % 4;
        return (uint64_t)(4 + frac) << (lg - 2);
    };
    for (int b = kNumDegBins - 1; b >= 0; --b) {          // descending degree
        if (b == degree_bin(kRingMinDeg) - 1) n_ring = pos;
        if (b == degree_bin(kGiant1MinDeg) - 1) n_giant1 = pos;
        if (bin_min_degree(b) >= px_deg) n_px = pos + hist[b];       // bins are walked by descending degree
        if (b == degree_bin(kMidMinDeg) - 1) n_pre = pos;
        start[b] = pos;
        pos += hist[b];
    }
    c->n_live = nl - hist[0];                            // bin 0 = degree 0, the end of `order`
    Schedule &sc = c->sched;
    sc.n_local = nl;
    sc.n_ring = n_ring;
    sc.n_giant1 = n_giant1;
    sc.n_px = std::min(n_px, std::min(n_ring, n_giant1));
    sc.n_mid = n_pre - n_ring;
    sc.n_tiles = (nl - n_pre + kTileVerts - 1) / kTileVerts;
    sc.n_feat_tiles = (n_pre + kTileVerts - 1) / kTileVerts;
    if ((rc = c->d_feat.reserve((size_t)n_pre * 32))) return rc;
    // counters: task claims, feature tiles, then one "chunks done" counter per split vertex
    c->px_ctr_off = kSyncCounters + sc.n_feat_tiles + std::max(n_ring, n_giant1);
    if ((rc = c->d_sync.reserve((size_t)c->px_ctr_off + 8 + 3 * (size_t)sc.n_px))) return rc;
    GVC_CUDA(cudaMemcpyAsync(c->d_bins.p, start, sizeof(start), cudaMemcpyHostToDevice, c->stream));
    degree_scatter_kernel<<<grid, 256, 0, c->stream>>>(c->f_row_ptr, c->f_W, nl, c->d_bins.p, c->d_order.p, c->d_vrec.p);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    // fast-mode chunk lists of the two hub classes
    const uint32_t n_class[3] = {n_ring, n_giant1, sc.n_px}, chunk_len[3] = {kChunk16, kChunk1, kPxChunk};
    uint32_t n_chunks[3] = {0, 0, 0};
    if ((rc = c->d_hub_count.reserve(3))) return rc;
    GVC_CUDA(cudaMemsetAsync(c->d_hub_count.p, 0, 3 * sizeof(uint32_t), c->stream));
    for (int k = 0; k < 3; ++k) {
        if (!n_class[k]) continue;
        const size_t bound = (size_t)(c->nnz / chunk_len[k]) + n_class[k];
        if ((rc = c->d_hub_chunk[k].reserve(bound))) return rc;
        if ((rc = c->d_hub_info[k].reserve(n_class[k]))) return rc;
        if (k < 2 && (rc = c->d_hub_partial[k].reserve(bound * (k == 0 ? 16 : 1)))) return rc;
        hub_chunks_kernel<<<std::min<unsigned>(1184, (n_class[k] + 255) / 256), 256, 0, c->stream>>>(
            c->d_vrec.p, n_class[k], chunk_len[k], c->d_hub_count.p + k, c->d_hub_info[k].p, c->d_hub_chunk[k].p,
            (uint32_t)std::min<size_t>(bound, 0xFFFFFFFFu));
        GVC_CUDA(cudaGetLastError());
        c->launches++;
    }
    GVC_CUDA(cudaMemcpyAsync(n_chunks, c->d_hub_count.p, sizeof(n_chunks), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));   // `start` and `n_chunks` live on this stack frame
    sc.n_chunks16 = (uint32_t)std::min<uint64_t>(n_chunks[0], c->nnz / kChunk16 + n_ring);
    sc.n_chunks1 = (uint32_t)std::min<uint64_t>(n_chunks[1], c->nnz / kChunk1 + n_giant1);
    sc.n_chunks_px = (uint32_t)std::min<uint64_t>(n_chunks[2], c->nnz / kPxChunk + sc.n_px);
    if (sc.n_chunks_px) {   // scratch copy of the gathered rows and the per-batch records of the parallel exact sums
        if ((rc = c->d_px_scratch.reserve((size_t)sc.n_chunks_px * kPxChunk * 16))) return rc;
        if ((rc = c->d_px_S.reserve((size_t)sc.n_chunks_px * 64 * 16))) return rc;
        if ((rc = c->d_px_T.reserve((size_t)sc.n_chunks_px * 16))) return rc;
        if ((rc = c->d_px_rec.reserve((size_t)sc.n_chunks_px * 64 * 32))) return rc;
        if ((rc = c->d_px_flag.reserve((size_t)sc.n_chunks_px * 64))) return rc;
    }
    // (the per-vertex peer lists read the adjacency itself: the callers build them once it has landed)
    c->have_peer_mask = false;
    return 0;
}

// ---- packed CSR from ranges into a raw edge span (gvc_graph_upload_stream) --------------------------
constexpr uint32_t kBadRange = 4u;
constexpr int kScanItems = 4096;          // vertices per CTA of the degree scan (256 threads x 16)

// degree of vertex u, and the range check: begin <= end <= span_len
__device__ __forceinline__ uint32_t range_degree(const uint32_t *__restrict__ rb, const uint32_t *__restrict__ re,
                                                 uint32_t u, uint64_t span_len, bool &bad) {
    const uint32_t b = rb[u], e = re[u];
    if (b > e || e > span_len) { bad = true; return 0u; }
    return e - b;
}

__global__ void __launch_bounds__(256)
range_block_sums_kernel(const uint32_t *__restrict__ rb, const uint32_t *__restrict__ re, uint32_t n, uint64_t span_len,
                        uint64_t *__restrict__ blk, uint32_t *__restrict__ flag) {
    __shared__ uint64_t part[8];
    const uint32_t base = blockIdx.x * kScanItems;
    bool bad = false;
    uint64_t sum = 0;
    for (int t = 0; t < 16; ++t) {
        const uint32_t u = base + t * 256 + threadIdx.x;
        if (u < n) sum += range_degree(rb, re, u, span_len, bad);
    }
    for (int m = 16; m; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t s = 0;
        for (int w = 0; w < 8; ++w) s += part[w];
        blk[blockIdx.x] = s;
    }
    if (bad) atomicOr(flag, kBadRange);
}

// exclusive scan of the block sums in place, one CTA; blk[nb] receives the total
__global__ void __launch_bounds__(1024)
range_scan_blocks_kernel(uint64_t *__restrict__ blk, uint32_t nb) {
    __shared__ uint64_t part[1024];
    const uint32_t per = (nb + 1023) / 1024;
    const uint32_t lo = min(nb, threadIdx.x * per), hi = min(nb, lo + per);
    uint64_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += blk[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int t = 0; t < 1024; ++t) { const uint64_t v = part[t]; part[t] = run; run += v; }
        blk[nb] = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) { const uint64_t v = blk[i]; blk[i] = run; run += v; }
}

// row_ptr[u] = exclusive prefix of the degrees (32-bit: the caller has checked the total)
__global__ void __launch_bounds__(256)
range_row_ptr_kernel(const uint32_t *__restrict__ rb, const uint32_t *__restrict__ re, uint32_t n, uint64_t span_len,
                     const uint64_t *__restrict__ blk, uint32_t *__restrict__ row_ptr) {
    __shared__ uint32_t warp_sum[8];
    const uint32_t base = blockIdx.x * kScanItems + threadIdx.x * 16;      // 16 consecutive vertices per thread
    bool bad = false;
    uint32_t d[16], s = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) { d[t] = (base + t < n) ? range_degree(rb, re, base + t, span_len, bad) : 0u; s += d[t]; }
    uint32_t inc = s;                                                      // inclusive scan over the CTA's threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int m = 1; m < 32; m <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, m); if (lane >= m) inc += v; }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t off = (uint32_t)blk[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += warp_sum[w];
    uint32_t run = off + inc - s;
#pragma unroll
    for (int t = 0; t < 16; ++t)
        if (base + t < n) { row_ptr[base + t] = run; run += d[t]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) row_ptr[n] = (uint32_t)blk[gridDim.x];
}

// col[packed offset + i] = span[begin[u] + i], by degree class of the schedule (a list of 64 452 entries
// and a list of 3 must not cost the same): warp tasks, grid-strided --
//   ring vertices (deg >= 2048)  their 4096-entry chunks (the fast-mode chunk list), 512 entries per warp task
//   mid vertices (64 <= deg)     one warp per vertex
//   the rest                     32 vertices per warp, their short lists copied one after the other
// vrec[pos] = {id, packed begin, packed end, W}.  Every id is checked against n_global on the way.
__global__ void __launch_bounds__(256)
range_compact_kernel(const uint32_t *__restrict__ span, const uint32_t *__restrict__ rb, const uint4 *__restrict__ vrec,
                     const uint4 *__restrict__ ring_chunk, uint32_t n_ring_chunks, uint32_t n_ring, uint32_t n_mid,
                     uint32_t n_local, uint32_t n_global, uint32_t *__restrict__ col, uint32_t *__restrict__ flag,
                     const uint32_t *__restrict__ remap /* null, or new id of every vertex (relabel_rows) */) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t t_ring = n_ring_chunks * 8u, t_mid = n_mid, n_low = n_local - n_ring - n_mid;
    const uint32_t n_tasks = t_ring + t_mid + (n_low + 31) / 32;
    bool bad = false;
    auto copy = [&](uint32_t src, uint32_t dst, uint32_t len) {
        for (uint32_t i = lane; i < len; i += 32) {
            const uint32_t id = __ldcs(span + src + i);
            bad |= id >= n_global;
            col[dst + i] = (remap && id < n_global) ? __ldg(remap + id) : id;
        }
    };
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tasks; t += warps) {
        if (t < t_ring) {
            const uint4 ck = __ldg(ring_chunk + (t >> 3));                 // {pos, packed first, packed end, chunk}
            const uint4 r = __ldg(vrec + ck.x);
            const uint32_t lo = ck.y + (t & 7u) * 512u, hi = min(ck.z, lo + 512u);
            if (lo < hi) copy(__ldg(rb + r.x) + (lo - r.y), lo, hi - lo);
        } else if (t < t_ring + t_mid) {
            const uint4 r = __ldg(vrec + n_ring + (t - t_ring));
            copy(__ldg(rb + r.x), r.y, r.z - r.y);
        } else {
            const uint32_t pos = n_ring + n_mid + (t - t_ring - t_mid) * 32u + lane;
            uint4 r = make_uint4(0u, 0u, 0u, 0u);
            uint32_t src = 0;
            if (pos < n_local) { r = __ldg(vrec + pos); src = __ldg(rb + r.x); }
            for (int v = 0; v < 32; ++v) {
                const uint32_t s_v = __shfl_sync(0xffffffffu, src, v), d_v = __shfl_sync(0xffffffffu, r.y, v),
                               l_v = __shfl_sync(0xffffffffu, r.z - r.y, v);
                copy(s_v, d_v, l_v);
            }
        }
    }
    if (bad) atomicOr(flag, kBadCol);
}

// ---- rows in schedule order (large whole-graph contexts) ---------------------------------------------------
// A vertex's 64-byte row shares its 128-byte L2 line with the row of the next vertex id.  On a graph whose rows
// do not fit L2 (R-MAT scale 23: 537 MB against 126 MB) what L2 holds are the rows of the high-degree
// vertices -- each dragging a cold neighbour's row along.  With the vertices renumbered in the order of the degree
// schedule the hot rows lie next to each other and L2 holds twice as many of them: measured on that graph
// (tools/relabel_probe.py) the three stages take 1.74 / 4.23 / 3.92 ms instead of 1.78 / 4.46 / 4.18.
// So for whole-graph contexts of GVC_ROW_ORDER_MIN_VERTICES (default 500 000) vertices and more the fused path
// works on an internal copy of the graph in that numbering (built on the device with the streamed upload's
// kernels: the original adjacency is the "span", the new rows are ranges into it, ids are translated on the
// way); x is permuted on the way in, scores and selection keys on the way out.  Per-vertex arithmetic and the
// order of every neighbour sum are unchanged -- same bits.  The generic per-layer kernels and the training path
// keep reading the original arrays.
__global__ void relabel_inverse_kernel(const uint32_t *__restrict__ order, uint32_t n, uint32_t *__restrict__ pos_of) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) pos_of[order[p]] = p;
}
__global__ void relabel_vertices_kernel(const uint32_t *__restrict__ order, const uint32_t *__restrict__ row_ptr,
                                        const uint32_t *__restrict__ W, const uint32_t *__restrict__ NW, uint32_t n,
                                        uint32_t *__restrict__ rb, uint32_t *__restrict__ re, uint32_t *__restrict__ W2,
                                        uint32_t *__restrict__ NW2) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t v = order[p];
        rb[p] = row_ptr[v]; re[p] = row_ptr[v + 1]; W2[p] = W[v]; NW2[p] = NW[v];
    }
}
template <typename T>
__global__ void rows_gather_kernel(const T *__restrict__ src, const uint32_t *__restrict__ order, uint32_t n, T *__restrict__ dst) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) dst[p] = src[order[p]];
}
template <typename T>
__global__ void rows_scatter_kernel(const T *__restrict__ src, const uint32_t *__restrict__ order, uint32_t n, T *__restrict__ dst) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) dst[order[p]] = src[p];
}

// WHEN: building the copy costs about as much as fifteen forwards save (R-MAT scale 23: 6.6 ms against 0.4 ms per
// forward), and a solver's predict() runs ONE forward per uploaded graph -- so the rows are reordered when a graph
// is forwarded for the GVC_ROW_ORDER_AFTER-th time after its upload (default 1: at the start of the second
// forward; 0: at upload already).  Callers that keep a graph resident and run many forwards on it (the
// device-timed benchmark loop, batched scoring of one graph with several inputs) get it, predict() does not.
// Called by every upload path once the original CSR is complete, validated and scheduled (at_upload), and at
// the start of every fused forward.
int relabel_rows(gvc_ctx *c, bool at_upload) {
    if (c->relabelled) return 0;
    const char *env = std::getenv("GVC_ROW_ORDER_MIN_VERTICES");          // read per call: tests switch it
    const uint64_t min_n = env ? std::strtoull(env, nullptr, 10) : 500000ull;
    const char *env_after = std::getenv("GVC_ROW_ORDER_AFTER");
    const uint64_t after = env_after ? std::strtoull(env_after, nullptr, 10) : 1ull;
    if (at_upload ? after != 0 : c->forwards_on_graph < after) return 0;
    const uint32_t n = c->n_global;
    if (c->v_begin != 0 || c->v_end != n || n < min_n || n < 2 || c->nnz == 0) return 0;
    int rc;
    const unsigned grid = std::min<unsigned>(1184, (n + 255) / 256);
    if ((rc = c->d_row_order.reserve(n))) return rc;
    if ((rc = c->d_pos_of.reserve(n))) return rc;
    if ((rc = c->d_rb.reserve(n))) return rc;
    if ((rc = c->d_re.reserve(n))) return rc;
    if ((rc = c->r_W.reserve(n))) return rc;
    if ((rc = c->r_NW.reserve(n))) return rc;
    if ((rc = c->r_row_ptr.reserve((size_t)n + 1))) return rc;
    if ((rc = c->r_col.reserve(c->nnz + 4))) return rc;
    if ((rc = c->d_xp.reserve(n))) return rc;
    if ((rc = c->d_sp.reserve(n))) return rc;
    if ((rc = c->d_keys_p.reserve(n))) return rc;
    if ((rc = c->d_side_p.reserve(n))) return rc;
    if ((rc = c->d_flag.reserve(1))) return rc;
    const uint32_t n_scan = (n + kScanItems - 1) / kScanItems;
    if ((rc = c->d_blk.reserve((size_t)n_scan + 1))) return rc;
    // row r of the internal numbering = position r of the schedule just built on the original graph
    GVC_CUDA(cudaMemcpyAsync(c->d_row_order.p, c->d_order.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    relabel_inverse_kernel<<<grid, 256, 0, c->stream>>>(c->d_row_order.p, n, c->d_pos_of.p);
    relabel_vertices_kernel<<<grid, 256, 0, c->stream>>>(c->d_row_order.p, c->row_ptr, c->Wv, c->NWv, n, c->d_rb.p, c->d_re.p,
                                                          c->r_W.p, c->r_NW.p);
    GVC_CUDA(cudaMemsetAsync(c->d_flag.p, 0, sizeof(uint32_t), c->stream));
    range_block_sums_kernel<<<n_scan, 256, 0, c->stream>>>(c->d_rb.p, c->d_re.p, n, c->nnz, c->d_blk.p, c->d_flag.p);
    range_scan_blocks_kernel<<<1, 1024, 0, c->stream>>>(c->d_blk.p, n_scan);
    range_row_ptr_kernel<<<n_scan, 256, 0, c->stream>>>(c->d_rb.p, c->d_re.p, n, c->nnz, c->d_blk.p, c->r_row_ptr.p);
    GVC_CUDA(cudaGetLastError());
    c->launches += 5;
    // the schedule of the renumbered graph (its `order` is the identity up to ties inside a degree bin)
    const uint32_t *span = c->col;
    c->f_row_ptr = c->r_row_ptr.p; c->f_col = c->r_col.p; c->f_W = c->r_W.p; c->f_NW = c->r_NW.p;
    if ((rc = build_schedule(c))) { c->have_graph = false; return rc; }
    const Schedule &sc = c->sched;
    const uint64_t warp_tasks = (uint64_t)sc.n_chunks16 * 8 + sc.n_mid + (n - sc.n_ring - sc.n_mid + 31) / 32;
    range_compact_kernel<<<(unsigned)std::min<uint64_t>(148 * 8, (warp_tasks + 7) / 8), 256, 0, c->stream>>>(
        span, c->d_rb.p, c->d_vrec.p, c->d_hub_chunk[0].p, sc.n_chunks16, sc.n_ring, sc.n_mid, n, n, c->r_col.p, c->d_flag.p,
        c->d_pos_of.p);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    // OpenBLAS' 1-row remainder kernel belongs to the LAST vertex of the caller's numbering (odd counts only)
    uint32_t tail_row = 0, flag = 0;
    GVC_CUDA(cudaMemcpyAsync(&tail_row, c->d_pos_of.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    if (flag) { c->have_graph = false; return fail(GVC_ERR_STATE, "internal renumbering failed its own checks (%u)", flag); }
    c->tail_override = (n & 1u) ? 1 : 0;
    c->tail_local = tail_row;
    c->relabelled = true;
    return 0;
}

// x of the caller's numbering into row order; scores and selection keys back (relabelled contexts only)
void rows_in(gvc_ctx *c, const float *d_x) {
    const uint32_t n = c->n_global;
    rows_gather_kernel<float><<<std::min<unsigned>(1184, (n + 255) / 256), 256, 0, c->stream>>>(d_x, c->d_row_order.p, n, c->d_xp.p);
    c->launches++;
}
int rows_out(gvc_ctx *c, float *d_scores) {
    const uint32_t n = c->n_global;
    const unsigned grid = std::min<unsigned>(1184, (n + 255) / 256);
    rows_scatter_kernel<float><<<grid, 256, 0, c->stream>>>(c->d_sp.p, c->d_row_order.p, n, d_scores);
    c->launches++;
    if (c->keys_out) {
        rows_scatter_kernel<float><<<grid, 256, 0, c->stream>>>(c->d_keys_p.p, c->d_row_order.p, n, c->keys_out);
        rows_scatter_kernel<uint8_t><<<grid, 256, 0, c->stream>>>(c->d_side_p.p, c->d_row_order.p, n, c->side_out);
        c->launches += 2;
    }
    GVC_CUDA(cudaGetLastError());
    return 0;
}

// The ring of pinned slots (grow-only; sized once for the first, largest graph of a GNN_VC run).
int ensure_ring(gvc_ctx *c, int n_slots, size_t slot_bytes) {
    if (c->n_slots >= n_slots && c->slot_bytes >= slot_bytes) return 0;
    for (auto &e : c->slot_ev) cudaEventDestroy(e);
    c->slot_ev.clear();
    c->ring.release();
    c->n_slots = 0; c->slot_bytes = 0;
    int rc;
    if ((rc = c->ring.reserve((size_t)n_slots * slot_bytes))) return rc;
    c->slot_ev.resize(n_slots);
    for (auto &e : c->slot_ev) GVC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->n_slots = n_slots; c->slot_bytes = slot_bytes;
    return 0;
}

template <int STAGE>
int set_stage_attrs(gvc_ctx *c) {
    const int smem = (int)stage_smem_bytes<STAGE, true>(), smem_fast = (int)stage_smem_bytes<STAGE, false>();
    GVC_CUDA(cudaFuncSetAttribute(stage_kernel<STAGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GVC_CUDA(cudaFuncSetAttribute(stage_kernel<STAGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fast));
    int a = 0, b = 0;
    GVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, stage_kernel<STAGE, true>, kCtaThreads, smem));
    GVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, stage_kernel<STAGE, false>, kCtaThreads, smem_fast));
    c->ctas_per_sm[STAGE] = std::max(1, std::min(kCtasPerSm, std::min(a, b)));
    return 0;
}

int check_ctx(const gvc_ctx *c) {
    if (!c) return fail(GVC_ERR_ARG, "null context");
    return 0;
}

int use_device(const gvc_ctx *c) {
    GVC_CUDA(cudaSetDevice(c->device));
    tl_arena = &const_cast<gvc_ctx *>(c)->arena;
    return 0;
}

// Before the first big burst of reservations: one chunk for everything a graph of n vertices and
// `entries` adjacency entries needs (offsets, ids twice while a streamed upload compacts them, weights,
// schedule, x, h1, h2, scores, feature vectors of the high-degree vertices).
void arena_hint(gvc_ctx *c, uint64_t n, uint64_t entries) {
    if (!c->arena.chunks.empty()) return;
    c->arena.next_chunk = (size_t)(n * 216 + entries * 9 + (4u << 20));
}

int ensure_activations(gvc_ctx *c) {
    int rc;
    if ((rc = c->d_x.reserve(c->n_global))) return rc;
    if ((rc = c->d_h1.reserve((size_t)c->n_global * 16))) return rc;
    if ((rc = c->d_h2.reserve((size_t)c->n_global * 16))) return rc;
    if ((rc = c->d_scores.reserve(c->n_local()))) return rc;
    return 0;
}

// generic path: any layer sequence, single shard
int forward_generic(gvc_ctx *c, const float *d_x, float scale, float *d_scores, int mode) {
    const uint32_t n = c->n_global;
    int w = 1, maxw = 1;
    for (auto &l : c->layers) {
        if (l.kind == GVC_LINEAR) w = l.cols;
        else if (l.kind == GVC_GRAPH) w = 2 * w + 3;
        if (w > maxw) maxw = w;
    }
    if (w != 1) return fail(GVC_ERR_UNSUPPORTED, "model output width %d != 1", w);
    int rc;
    if ((rc = c->d_ping.reserve((size_t)n * maxw))) return rc;
    if ((rc = c->d_pong.reserve((size_t)n * maxw))) return rc;
    const float *cur = d_x;
    float *bufs[2] = {c->d_ping.p, c->d_pong.p};
    int which = 0;
    w = 1;
    const int T = 256;
    size_t gi = 0;                       // index among the graph layers
    for (size_t li = 0; li < c->layers.size(); ++li) {
        auto &l = c->layers[li];
        const bool last = li + 1 == c->layers.size();
        float *dst = last ? d_scores : bufs[which];
        int wo = w;
        switch (l.kind) {
        case GVC_LINEAR:
            if (l.rows != w) return fail(GVC_ERR_ARG, "layer %zu expects width %d, got %d", li, l.rows, w);
            wo = l.cols;
            if (mode == GVC_MODE_EXACT)
                generic_linear_kernel<true><<<blocks_for((uint64_t)n * wo, T), T, 0, c->stream>>>(cur, l.rows, l.cols, l.dW, l.db, dst, n);
            else
                generic_linear_kernel<false><<<blocks_for((uint64_t)n * wo, T), T, 0, c->stream>>>(cur, l.rows, l.cols, l.dW, l.db, dst, n);
            break;
        case GVC_GRAPH: {
            wo = 2 * w + 3;
            const float ls = gi < c->layer_scales.size() ? c->layer_scales[gi] : scale;   // this layer's own WEIGHT_SCALE
            ++gi;
            generic_graph_kernel<<<blocks_for((uint64_t)n * wo, T), T, 0, c->stream>>>(c->row_ptr, c->col, c->Wv, c->NWv, cur, w, dst, n, 0, ls);
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
    return 0;
}

int set_graph_views(gvc_ctx *c, uint32_t n_global, uint32_t v_begin, uint32_t v_end, const uint32_t *rp,
                    const uint32_t *col, const uint32_t *W, const uint32_t *NW, uint64_t nnz) {
    c->n_global = n_global; c->v_begin = v_begin; c->v_end = v_end; c->nnz = nnz;
    c->row_ptr = rp; c->col = col; c->Wv = W; c->NWv = NW;
    c->f_row_ptr = rp; c->f_col = col; c->f_W = W; c->f_NW = NW;
    c->relabelled = false;
    c->forwards_on_graph = 0;
    c->have_graph = true;
    c->keys_valid_n = 0;
    c->x_resident_n = 0;
    c->tail_override = -1;
    int rc;
    if ((rc = ensure_activations(c))) return rc;
    return build_schedule(c);
}

}  // namespace

// ================================ C ABI ==========================================
extern "C" {

const char *gvc_last_error(void) { return g_err.c_str(); }
int gvc_internal_fail(int code, const char *msg) { return fail(code, "%s", msg); }   // for the other translation units of libgvc
int gvc_abi_version(void) { return 4; }   // 4: upload_stream_x, ctx_warm, trainer + backward entry points (additive)

int gvc_ctx_create(gvc_ctx **out, int device) {
    if (!out) return fail(GVC_ERR_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(GVC_ERR_NO_DEVICE, "no CUDA device: %s (libgvc has no CPU path)", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(GVC_ERR_ARG, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    GVC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GVC_ERR_NO_DEVICE, "device %d is sm_%d%d; libgvc is built for sm_100a only", device, prop.major, prop.minor);
    gvc_ctx *c = new (std::nothrow) gvc_ctx();
    if (!c) return fail(GVC_ERR_ALLOC, "out of host memory");
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    GVC_CUDA(cudaSetDevice(device));
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(1000 + (int)e, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    c->stream = c->own_stream;
    int rc = 0;
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_begin, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_vert, cudaEventDisableTiming)) != cudaSuccess)
        rc = fail(1000 + (int)e, "copy stream/events: %s", cudaGetErrorString(e));
    if (rc || (rc = set_stage_attrs<0>(c)) || (rc = set_stage_attrs<1>(c)) || (rc = set_stage_attrs<2>(c))) {
        if (c->ev_begin) cudaEventDestroy(c->ev_begin);
        if (c->ev_copy) cudaEventDestroy(c->ev_copy);
        if (c->ev_vert) cudaEventDestroy(c->ev_vert);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        cudaStreamDestroy(c->own_stream);
        delete c;
        return rc;
    }
    *out = c;
    return 0;
}

void gvc_ctx_destroy(gvc_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    for (auto &l : c->layers) { if (l.dW) cudaFree(l.dW); if (l.db) cudaFree(l.db); }
    for (auto &p : c->d_stage_params) if (p) cudaFree(p);
    for (auto &p : c->d_stage_params_fast) if (p) cudaFree(p);
    c->own_row_ptr.release(); c->own_col.release(); c->own_W.release(); c->own_NW.release();
    c->d_row_ptr64.release(); c->d_flag.release();
    c->d_span.release(); c->d_rb.release(); c->d_re.release(); c->d_blk.release();
    for (auto &e : c->slot_ev) cudaEventDestroy(e);
    c->ring.release();
    c->stg_row_ptr.release(); c->stg_col.release(); c->stg_W.release(); c->stg_NW.release();
    c->d_order.release(); c->d_vrec.release(); c->d_bins.release(); c->d_sync.release(); c->d_feat.release();
    c->d_peer_mask.release();
    for (int k = 0; k < 3; ++k) { c->d_hub_chunk[k].release(); c->d_hub_info[k].release(); }
    for (int k = 0; k < 2; ++k) c->d_hub_partial[k].release();
    c->d_px_scratch.release(); c->d_px_S.release(); c->d_px_T.release(); c->d_px_rec.release(); c->d_px_flag.release();
    c->d_hub_count.release();
    c->d_x.release(); c->d_h1.release(); c->d_h2.release(); c->d_scores.release();
    c->d_ping.release(); c->d_pong.release(); c->d_keys.release(); c->d_side.release();
    c->pin_x.release(); c->pin_scores.release();
    c->arena.release_all();
    tl_arena = nullptr;
    cudaEventDestroy(c->ev_begin);
    cudaEventDestroy(c->ev_copy);
    cudaEventDestroy(c->ev_vert);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int gvc_model_upload(gvc_ctx *c, int n_layers, const int *kinds, const int *rows, const int *cols,
                     const float *const *W, const float *const *bias) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (n_layers <= 0 || !kinds) return fail(GVC_ERR_ARG, "empty model");   // assert(!layers.empty()), :68
    // validate everything before the context is touched: a rejected model leaves the old one in place
    for (int i = 0; i < n_layers; ++i) {
        if (kinds[i] < GVC_LINEAR || kinds[i] > GVC_SIGMOID) return fail(GVC_ERR_ARG, "layer %d: unknown kind %d", i, kinds[i]);
        if (kinds[i] != GVC_LINEAR) continue;
        if (!rows || !cols || !W || !bias || !W[i] || !bias[i] || rows[i] <= 0 || cols[i] <= 0)
            return fail(GVC_ERR_ARG, "layer %d: linear layer needs rows, cols, W and bias", i);
    }
    if ((rc = use_device(c))) return rc;
    GVC_CUDA(cudaStreamSynchronize(c->stream));      // no forward may still read the old parameters
    // From here on a failure (CUDA allocation or copy) leaves the context WITHOUT a model -- never
    // with a half-built one that the stage kernels could be launched on.
    auto drop_model = [c] {
        for (auto &l : c->layers) { if (l.dW) cudaFree(l.dW); if (l.db) cudaFree(l.db); }
        c->layers.clear();
        for (auto &p : c->d_stage_params) { if (p) cudaFree(p); p = nullptr; }
        for (auto &p : c->d_stage_params_fast) { if (p) cudaFree(p); p = nullptr; }
        c->fused = false;
        c->layer_scales.clear();
    };
    drop_model();
    auto upload = [&]() -> int {
        c->layers.resize(n_layers);
        for (int i = 0; i < n_layers; ++i) {
            HostLayer &l = c->layers[i];
            l.kind = kinds[i];
            if (l.kind != GVC_LINEAR) continue;
            l.rows = rows[i]; l.cols = cols[i];
            l.W.assign(W[i], W[i] + (size_t)l.rows * l.cols);
            l.b.assign(bias[i], bias[i] + l.cols);
            GVC_CUDA(cudaMalloc(&l.dW, l.W.size() * sizeof(float)));
            GVC_CUDA(cudaMalloc(&l.db, l.b.size() * sizeof(float)));
            GVC_CUDA(cudaMemcpyAsync(l.dW, l.W.data(), l.W.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            GVC_CUDA(cudaMemcpyAsync(l.db, l.b.data(), l.b.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        if (detect_fused(c->layers)) {
            for (int s = 0; s < 3; ++s) {
                std::vector<float> p = pack_stage(c->layers, s);
                GVC_CUDA(cudaMalloc(&c->d_stage_params[s], p.size() * sizeof(float)));
                GVC_CUDA(cudaMemcpy(c->d_stage_params[s], p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
                std::vector<float> pf = pack_stage_mma(c->layers, s);
                GVC_CUDA(cudaMalloc(&c->d_stage_params_fast[s], pf.size() * sizeof(float)));
                GVC_CUDA(cudaMemcpy(c->d_stage_params_fast[s], pf.data(), pf.size() * sizeof(float), cudaMemcpyHostToDevice));
            }
        }
        GVC_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    if ((rc = upload())) { drop_model(); return rc; }
    c->fused = detect_fused(c->layers);              // only a completely uploaded model is ever "fused"
    return 0;
}

int gvc_model_weight_scales(gvc_ctx *c, int n_graph_layers, const float *scales) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (n_graph_layers == 0) { c->layer_scales.clear(); return 0; }
    int have = 0;
    for (auto &l : c->layers) have += l.kind == GVC_GRAPH;
    if (n_graph_layers != have) return fail(GVC_ERR_ARG, "%d scales for a model with %d graph layers", n_graph_layers, have);
    if (!scales) return fail(GVC_ERR_ARG, "null scales");
    c->layer_scales.assign(scales, scales + n_graph_layers);
    return 0;
}

int gvc_model_is_fused(const gvc_ctx *c) { return c && c->fused ? 1 : 0; }

int gvc_graph_upload_shard(gvc_ctx *c, uint32_t n_global, uint32_t v_begin, uint32_t v_end,
                           const uint64_t *row_ptr, const uint32_t *col, const uint32_t *W,
                           const uint32_t *NW) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (v_begin > v_end || v_end > n_global) return fail(GVC_ERR_ARG, "bad shard [%u,%u) of %u", v_begin, v_end, n_global);
    const uint32_t nl = v_end - v_begin;
    if (nl && (!row_ptr || !W || !NW)) return fail(GVC_ERR_ARG, "null graph arrays");
    if ((rc = use_device(c))) return rc;
    const uint64_t nnz = nl ? row_ptr[nl] : 0;
    if (nl && row_ptr[0] != 0) return fail(GVC_ERR_ARG, "row_ptr[0] must be 0");
    if (nnz >= (1ull << 32)) return fail(GVC_ERR_UNSUPPORTED, "shard has %llu adjacency entries; 2^32 is the limit", (unsigned long long)nnz);
    if (nnz && !col) return fail(GVC_ERR_ARG, "null col");
    Tracer tr;
    arena_hint(c, std::max<uint64_t>(nl, n_global / 2), nnz);
    if ((rc = c->own_row_ptr.reserve((size_t)nl + 1))) return rc;
    if ((rc = c->own_col.reserve(nnz + 4))) return rc;      // readable up to the next multiple of four ids
    if ((rc = c->own_W.reserve(nl))) return rc;
    if ((rc = c->own_NW.reserve(nl))) return rc;
    if ((rc = c->d_row_ptr64.reserve((size_t)nl + 1))) return rc;
    if ((rc = c->d_flag.reserve(1))) return rc;
    c->have_graph = false;
    tr.tick("upload: device buffers");
    // The adjacency (nearly all of the bytes) travels on the copy stream while the offsets are
    // narrowed and the degree schedule is built on the compute stream.  From the pinned buffers
    // of gvc_graph_staging both run as DMA; from pageable memory the copies stage through the
    // driver and the calls serialise, with the same result.
    GVC_CUDA(cudaMemsetAsync(c->d_flag.p, 0, sizeof(uint32_t), c->stream));
    GVC_CUDA(cudaEventRecord(c->ev_begin, c->stream));                 // earlier forwards are done with own_col
    GVC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_begin, 0));
    if (nl) {
        GVC_CUDA(cudaMemcpyAsync(c->d_row_ptr64.p, row_ptr, ((size_t)nl + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        GVC_CUDA(cudaMemcpyAsync(c->own_W.p, W, (size_t)nl * 4, cudaMemcpyHostToDevice, c->stream));
        GVC_CUDA(cudaMemcpyAsync(c->own_NW.p, NW, (size_t)nl * 4, cudaMemcpyHostToDevice, c->stream));
    }
    if (nnz) {
        GVC_CUDA(cudaMemcpyAsync(c->own_col.p, col, nnz * 4, cudaMemcpyHostToDevice, c->copy_stream));
        check_col_kernel<<<(unsigned)std::min<uint64_t>(1184, (nnz / 4 + 255) / 256 + 1), 256, 0, c->copy_stream>>>(c->own_col.p, nnz, n_global, c->d_flag.p);
        GVC_CUDA(cudaGetLastError());
        c->launches++;
    }
    GVC_CUDA(cudaEventRecord(c->ev_copy, c->copy_stream));
    tr.tick("upload: copies + id check");
    if (nl) {
        narrow_row_ptr_kernel<<<std::min<unsigned>(1184, nl / 256 + 1), 256, 0, c->stream>>>(c->d_row_ptr64.p, nl, nnz, c->own_row_ptr.p, c->d_flag.p);
        GVC_CUDA(cudaGetLastError());
        c->launches++;
    }
    // the schedule is built from the offsets: they must be known good first (a negative "degree"
    // would send the chunk builder past its buffers); the id check may still be running
    uint32_t flag = 0;
    GVC_CUDA(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    tr.tick("upload: narrow offsets");
    if (flag & kBadRowPtr) {
        cudaStreamSynchronize(c->copy_stream);
        return fail(GVC_ERR_ARG, "row_ptr is not monotone or exceeds row_ptr[n]");
    }
    if ((rc = set_graph_views(c, n_global, v_begin, v_end, c->own_row_ptr.p, c->own_col.p, c->own_W.p, c->own_NW.p, nnz))) {
        cudaStreamSynchronize(c->copy_stream);
        c->have_graph = false;
        return rc;
    }
    tr.tick("upload: schedule");
    GVC_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));           // forwards start after the adjacency landed
    // the peer lists are read off the adjacency: only now is it complete on this stream (building them
    // inside the schedule raced with the copy when gvc_peer_owners had been called before the upload)
    if ((rc = build_peer_mask(c))) {
        cudaStreamSynchronize(c->copy_stream);
        c->have_graph = false;
        return rc;
    }
    GVC_CUDA(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));                        // also: the caller's buffers are free again
    if (flag) {
        c->have_graph = false;
        return fail(GVC_ERR_ARG, "a neighbour id is >= %u vertices", n_global);
    }
    return relabel_rows(c, true);
}

int gvc_graph_upload_stream_x(gvc_ctx *c, uint32_t n, uint64_t span_len, gvc_fill_vertices_fn fill_vertices,
                              gvc_fill_span_fn fill_span, void *user, int n_threads, const float *x) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (n && !fill_vertices) return fail(GVC_ERR_ARG, "null vertex callback");
    if (span_len && !fill_span) return fail(GVC_ERR_ARG, "null span callback");
    if (span_len >= (1ull << 32)) return fail(GVC_ERR_UNSUPPORTED, "edge span of %llu entries; 2^32 is the limit", (unsigned long long)span_len);
    if ((rc = use_device(c))) return rc;
    Tracer tr;
    c->have_graph = false;
    c->x_resident_n = 0;
    // Work items: vertex chunks (begin, end, W, NW and -- when x travels along -- x: 5 arrays of v_chunk
    // words in one slot) and span chunks (one slot each).  Every item costs two or more driver calls
    // (copy + event), and those serialise over all threads: measured at 256 KB per slot on the 146 MB
    // benchmark graph, 544 items kept the copy engine at half of the PCIe rate.  Large graphs therefore
    // use 1 MB slots, small ones 256 KB (more pieces to overlap filling with copying).
    // Item size: as large as possible (fewer driver calls) while every worker still gets four or more items to
    // overlap its filling with its copies -- total / (4 x workers), between 64 KB and 1 MB.  (Tying the item size
    // to the ring's slot size, which never shrinks, left the later, smaller graphs of a solver run with 2-6
    // workers: ER 1M / 5M, second predict, 22 items of 1 MB on 6 threads.)  GVC_SLOT_KB overrides.
    static const size_t env_slot_kb = [] { const char *e = std::getenv("GVC_SLOT_KB"); return e ? (size_t)std::strtoull(e, nullptr, 10) : 0; }();
    static const int env_threads = [] { const char *e = std::getenv("GVC_UPLOAD_THREADS"); return e ? std::atoi(e) : 0; }();
    // Small graphs (up to GVC_UPLOAD_SINGLE_KB, default 4 MB, of vertex arrays + span) are filled and sent by the
    // calling thread alone, in four or five pieces: starting helper threads and their driver calls contending
    // with each other cost more than they overlap there.  Measured, `predict` on ER graphs with helper threads /
    // alone: 20 000 vertices (1.2 MB) 0.53 / 0.40 ms; 50 000 (3 MB) 0.69 / 0.60 ms; 100 000 (6 MB) 0.93 / 0.94 ms.
    static const uint64_t single_bytes = [] { const char *e = std::getenv("GVC_UPLOAD_SINGLE_KB"); return e ? std::strtoull(e, nullptr, 10) << 10 : 4ull << 20; }();
    const uint64_t total_bytes = 20ull * n + 4ull * span_len;
    const bool single = n_threads <= 0 && env_threads <= 0 && total_bytes <= single_bytes;     // an explicit thread count is honoured
    const int want_workers = single ? 1 : n_threads > 0 ? n_threads : env_threads > 0 ? env_threads
                                           : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    size_t slot_bytes = env_slot_kb ? (env_slot_kb << 10) : (size_t)std::min<uint64_t>(1u << 20, total_bytes / (4ull * want_workers));
    slot_bytes = std::max<size_t>(64u << 10, (slot_bytes + 65535) & ~(size_t)65535);
    const uint32_t v_chunk = (uint32_t)(slot_bytes / 20 / 1024 * 1024);           // 5 arrays per item
    const uint64_t s_chunk = slot_bytes / sizeof(uint32_t);
    const uint64_t n_vchunks = ((uint64_t)n + v_chunk - 1) / v_chunk, n_schunks = (span_len + s_chunk - 1) / s_chunk;
    const uint64_t n_items = n_vchunks + n_schunks;
    int workers = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)want_workers, (n_items + 3) / 4));   // >= 4 items per thread
    if ((rc = ensure_ring(c, std::max(2 * workers, 16) / workers * workers, slot_bytes))) return rc;
    arena_hint(c, n, span_len);
    if ((rc = c->d_rb.reserve(n))) return rc;
    if ((rc = c->d_re.reserve(n))) return rc;
    if ((rc = c->own_W.reserve(n))) return rc;
    if ((rc = c->own_NW.reserve(n))) return rc;
    if (x && (rc = c->d_x.reserve(n))) return rc;
    if ((rc = c->own_row_ptr.reserve((size_t)n + 1))) return rc;
    if ((rc = c->d_span.reserve(span_len + 4))) return rc;
    if ((rc = c->own_col.reserve(span_len + 4))) return rc;      // nnz <= span_len once the ranges are known good
    if ((rc = c->d_flag.reserve(1))) return rc;
    const uint32_t n_scan = (n + kScanItems - 1) / kScanItems;
    if ((rc = c->d_blk.reserve((size_t)n_scan + 1))) return rc;
    tr.tick("stream: buffers");
    GVC_CUDA(cudaMemsetAsync(c->d_flag.p, 0, sizeof(uint32_t), c->stream));
    GVC_CUDA(cudaEventRecord(c->ev_begin, c->stream));           // earlier forwards are done with the old arrays
    GVC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_begin, 0));

    // ---- host side: workers fill slots through the callbacks and send them off ------------------
    std::atomic<uint64_t> next{0}, v_issued{0};
    std::atomic<int> err{0};
    std::atomic<uint64_t> ns_wait{0}, ns_fill{0}, ns_issue{0};        // GVC_TRACE: where the workers' time goes
    auto now_ns = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const int slots_per_worker = c->n_slots / workers;
    auto work = [&](int w) {
        if (cudaSetDevice(c->device) != cudaSuccess) { err = 1; return; }
        int turn = 0;
        for (;;) {
            const uint64_t it = next.fetch_add(1);
            if (it >= n_items || err.load()) break;
            const int slot = slots_per_worker * w + (turn++ % slots_per_worker);
            unsigned char *host = c->ring.p + (size_t)slot * c->slot_bytes;
            const uint64_t t0 = tr.on ? now_ns() : 0;
            if (cudaEventSynchronize(c->slot_ev[slot]) != cudaSuccess) { err = 1; break; }    // its last copy has left
            const uint64_t t1 = tr.on ? now_ns() : 0;
            uint64_t t2 = 0;
            cudaError_t e = cudaSuccess;
            if (it < n_vchunks) {
                const uint32_t first = (uint32_t)(it * v_chunk), cnt = std::min<uint32_t>(v_chunk, n - first);
                uint32_t *b = reinterpret_cast<uint32_t *>(host), *en = b + v_chunk, *w_ = en + v_chunk, *nw = w_ + v_chunk;
                float *xs = reinterpret_cast<float *>(nw + v_chunk);
                fill_vertices(user, first, cnt, b, en, w_, nw);
                if (x) std::memcpy(xs, x + first, (size_t)cnt * sizeof(float));
                t2 = tr.on ? now_ns() : 0;
                e = cudaMemcpyAsync(c->d_rb.p + first, b, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->copy_stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_re.p + first, en, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->copy_stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(c->own_W.p + first, w_, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->copy_stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(c->own_NW.p + first, nw, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->copy_stream);
                if (e == cudaSuccess && x) e = cudaMemcpyAsync(c->d_x.p + first, xs, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->copy_stream);
            } else {
                const uint64_t off = (it - n_vchunks) * s_chunk, cnt = std::min<uint64_t>(s_chunk, span_len - off);
                fill_span(user, off, cnt, reinterpret_cast<uint32_t *>(host));
                t2 = tr.on ? now_ns() : 0;
                e = cudaMemcpyAsync(c->d_span.p + off, host, (size_t)cnt * 4, cudaMemcpyHostToDevice, c->copy_stream);
            }
            if (e == cudaSuccess) e = cudaEventRecord(c->slot_ev[slot], c->copy_stream);
            if (e != cudaSuccess) { err = 1000 + (int)e; break; }
            if (it < n_vchunks) v_issued.fetch_add(1);
            if (tr.on) { ns_wait += t1 - t0; ns_fill += t2 - t1; ns_issue += now_ns() - t2; }
        }
    };
    // ---- device side, part 1 (needs the per-vertex arrays only): degrees -> offsets, degree schedule.
    // With helper threads this runs on the calling thread WHILE the span chunks are still being filled
    // and copied: the vertex chunks are the first items, the event below follows their copies.
    uint64_t nnz = 0;
    auto offsets_and_schedule = [&]() -> int {
        if (n) {
            range_block_sums_kernel<<<n_scan, 256, 0, c->stream>>>(c->d_rb.p, c->d_re.p, n, span_len, c->d_blk.p, c->d_flag.p);
            range_scan_blocks_kernel<<<1, 1024, 0, c->stream>>>(c->d_blk.p, n_scan);
            GVC_CUDA(cudaGetLastError());
            c->launches += 2;
            uint32_t flag = 0;
            GVC_CUDA(cudaMemcpyAsync(&nnz, c->d_blk.p + n_scan, sizeof(nnz), cudaMemcpyDeviceToHost, c->stream));
            GVC_CUDA(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, c->stream));
            GVC_CUDA(cudaStreamSynchronize(c->stream));
            if (flag & kBadRange) return fail(GVC_ERR_ARG, "a vertex range is reversed or ends past the edge span");
            if (nnz >= (1ull << 32)) return fail(GVC_ERR_UNSUPPORTED, "graph has %llu adjacency entries; 2^32 is the limit", (unsigned long long)nnz);
            range_row_ptr_kernel<<<n_scan, 256, 0, c->stream>>>(c->d_rb.p, c->d_re.p, n, span_len, c->d_blk.p, c->own_row_ptr.p);
            GVC_CUDA(cudaGetLastError());
            c->launches++;
        }
        // the degree schedule needs offsets and weights only; the lists are moved class by class afterwards
        int r = set_graph_views(c, n, 0, n, c->own_row_ptr.p, c->own_col.p, c->own_W.p, c->own_NW.p, nnz);
        c->have_graph = false;                                   // not before the adjacency is in place and checked
        return r;
    };
    int rc_sched = 0;
    if (workers == 1 && (single || n_items <= 4)) {
        work(0);                                                  // small graph: no helper thread
        if (err.load()) { cudaStreamSynchronize(c->copy_stream); return fail(err.load(), "streamed upload: a copy failed"); }
        GVC_CUDA(cudaEventRecord(c->ev_copy, c->copy_stream));
        tr.tick("stream: fill + copies issued");
        GVC_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
        rc_sched = offsets_and_schedule();
    } else {
        std::vector<std::thread> th;
        for (int w = 0; w < workers; ++w) th.emplace_back(work, w);
        while (v_issued.load() < n_vchunks && !err.load()) std::this_thread::yield();
        cudaError_t e = cudaSuccess;
        if (!err.load()) {
            e = cudaEventRecord(c->ev_vert, c->copy_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_vert, 0);
            if (e == cudaSuccess) rc_sched = offsets_and_schedule();
        }
        for (auto &t : th) t.join();
        if (e != cudaSuccess) { cudaStreamSynchronize(c->copy_stream); return fail(1000 + (int)e, "streamed upload: %s", cudaGetErrorString(e)); }
        if (err.load()) { cudaStreamSynchronize(c->copy_stream); return fail(err.load(), "streamed upload: a copy failed"); }
        GVC_CUDA(cudaEventRecord(c->ev_copy, c->copy_stream));
        tr.tick("stream: fill + copies, offsets + schedule meanwhile");
        GVC_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    }
    if (tr.on)
        std::fprintf(stderr, "gvc trace: stream: %d workers, %llu items of %zu KB; per worker: wait for a slot %.3f ms, fill %.3f ms, issue %.3f ms\n",
                     workers, (unsigned long long)n_items, slot_bytes >> 10, ns_wait.load() * 1e-6 / workers, ns_fill.load() * 1e-6 / workers,
                     ns_issue.load() * 1e-6 / workers);
    if (rc_sched) { cudaStreamSynchronize(c->copy_stream); return rc_sched; }
    tr.tick("stream: offsets + schedule");

    // ---- device side, part 2: lists -> packed CSR, id checks -------------------------------------
    if (nnz) {
        const Schedule &sc = c->sched;
        const uint64_t warp_tasks = (uint64_t)sc.n_chunks16 * 8 + sc.n_mid + (n - sc.n_ring - sc.n_mid + 31) / 32;
        range_compact_kernel<<<(unsigned)std::min<uint64_t>(148 * 8, (warp_tasks + 7) / 8), 256, 0, c->stream>>>(
            c->d_span.p, c->d_rb.p, c->d_vrec.p, c->d_hub_chunk[0].p, sc.n_chunks16, sc.n_ring, sc.n_mid, n, n,
            c->own_col.p, c->d_flag.p, nullptr);
        GVC_CUDA(cudaGetLastError());
        c->launches++;
    }
    c->have_graph = true;
    if ((rc = build_peer_mask(c))) { c->have_graph = false; return rc; }
    uint32_t flag = 0;
    GVC_CUDA(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    tr.tick("stream: compaction");
    if (flag) {
        c->have_graph = false;
        return fail(GVC_ERR_ARG, "a neighbour id is >= %u vertices", n);
    }
    if ((rc = relabel_rows(c, true))) return rc;
    if (x) c->x_resident_n = n;
    return 0;
}

int gvc_ctx_warm(gvc_ctx *c, uint64_t graph_bytes_hint) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = use_device(c))) return rc;
    // what the first streamed upload would otherwise pay for inside predict(): pinning the ring (about
    // 0.5 ms per MB) and the first device allocation
    const bool large = graph_bytes_hint == 0 || graph_bytes_hint >= (32ull << 20);
    if ((rc = ensure_ring(c, 16, large ? ((size_t)1 << 20) : ((size_t)256 << 10)))) return rc;
    if (graph_bytes_hint) {
        c->arena.next_chunk = (size_t)std::min<uint64_t>(graph_bytes_hint, 8ull << 30);
        void *p = c->arena.alloc(256);
        if (!p) return GVC_ERR_ALLOC;
    }
    return 0;
}

int gvc_graph_upload_stream(gvc_ctx *c, uint32_t n, uint64_t span_len, gvc_fill_vertices_fn fill_vertices,
                            gvc_fill_span_fn fill_span, void *user, int n_threads) {
    return gvc_graph_upload_stream_x(c, n, span_len, fill_vertices, fill_span, user, n_threads, nullptr);
}

int gvc_graph_staging(gvc_ctx *c, uint32_t n_local, uint64_t nnz, uint64_t **row_ptr, uint32_t **col,
                      uint32_t **W, uint32_t **NW) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!row_ptr || !col || !W || !NW) return fail(GVC_ERR_ARG, "null out pointer");
    if ((rc = use_device(c))) return rc;
    Tracer tr;
    if ((rc = c->stg_row_ptr.reserve((size_t)n_local + 1))) return rc;
    if ((rc = c->stg_col.reserve(nnz + 4))) return rc;
    if ((rc = c->stg_W.reserve((size_t)n_local + 1))) return rc;
    if ((rc = c->stg_NW.reserve((size_t)n_local + 1))) return rc;
    tr.tick("staging: pinned buffers");
    *row_ptr = c->stg_row_ptr.p; *col = c->stg_col.p; *W = c->stg_W.p; *NW = c->stg_NW.p;
    return 0;
}

int gvc_graph_upload(gvc_ctx *c, uint32_t n, const uint64_t *row_ptr, const uint32_t *col,
                     const uint32_t *W, const uint32_t *NW) {
    return gvc_graph_upload_shard(c, n, 0, n, row_ptr, col, W, NW);
}

int gvc_graph_adopt_device(gvc_ctx *c, uint32_t n_global, uint32_t v_begin, uint32_t v_end,
                           const uint32_t *d_row_ptr, const uint32_t *d_col, const uint32_t *d_W,
                           const uint32_t *d_NW) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (v_begin > v_end || v_end > n_global) return fail(GVC_ERR_ARG, "bad shard [%u,%u) of %u", v_begin, v_end, n_global);
    if (v_end > v_begin && (!d_row_ptr || !d_W || !d_NW)) return fail(GVC_ERR_ARG, "null graph arrays");
    if ((rc = use_device(c))) return rc;
    // The gather reads neighbour ids as aligned groups of four (gvc_kernels.cuh ld_id4): the id
    // array must start on a 16-byte boundary and be readable up to the next multiple of four
    // entries.  A caller's buffer that does not guarantee that (a slice of a bigger array, a
    // length that is not a multiple of four) is copied once into a buffer that does.
    const uint32_t nl = v_end - v_begin;
    uint32_t nnz = 0;
    if (nl) {
        GVC_CUDA(cudaMemcpyAsync(&nnz, d_row_ptr + nl, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        GVC_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (nnz && !d_col) return fail(GVC_ERR_ARG, "null col");
    const uint32_t *col = d_col;
    if (nnz && ((reinterpret_cast<uintptr_t>(d_col) & 15u) || (nnz & 3u))) {
        if ((rc = c->own_col.reserve((size_t)nnz + 4))) return rc;
        GVC_CUDA(cudaMemcpyAsync(c->own_col.p, d_col, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, c->stream));
        col = c->own_col.p;
    }
    if ((rc = set_graph_views(c, n_global, v_begin, v_end, d_row_ptr, col, d_W, d_NW, nnz))) return rc;
    if ((rc = build_peer_mask(c))) return rc;          // same stream as the copy above: ordered
    return relabel_rows(c, true);
}

int gvc_graph_set_tail(gvc_ctx *c, int has_tail, uint32_t local_index) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (has_tail && local_index >= c->n_local()) return fail(GVC_ERR_ARG, "tail vertex %u outside the shard", local_index);
    c->tail_override = has_tail ? 1 : 0;
    c->tail_local = local_index;
    return 0;
}

int gvc_peer_alloc(gvc_ctx *c, uint64_t bytes, void **d_ptr, unsigned char *handle) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!d_ptr || !handle || !bytes) return fail(GVC_ERR_ARG, "null argument");
    if ((rc = use_device(c))) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == GVC_PEER_HANDLE_BYTES, "handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(GVC_ERR_ALLOC, "cudaMalloc(%llu bytes): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(1000 + (int)e, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    GVC_CUDA(cudaMemsetAsync(p, 0, bytes, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    std::memcpy(handle, &h, sizeof(h));
    *d_ptr = p;
    return 0;
}

int gvc_peer_open(gvc_ctx *c, const unsigned char *handle, void **d_ptr) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!d_ptr || !handle) return fail(GVC_ERR_ARG, "null argument");
    if ((rc = use_device(c))) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(1000 + (int)e, "cudaIpcOpenMemHandle: %s (peer access between the two GPUs is required)", cudaGetErrorString(e));
    *d_ptr = p;
    return 0;
}

int gvc_peer_close(gvc_ctx *c, void *d_ptr) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = use_device(c))) return rc;
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    for (auto &po : c->peers)
        for (int q = 0; q < po.n; ++q)
            if (po.p[q] == d_ptr) po.n = 0;                  // never keep a pointer that is about to vanish
    GVC_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int gvc_peer_free(gvc_ctx *c, void *d_ptr) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = use_device(c))) return rc;
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    GVC_CUDA(cudaFree(d_ptr));
    return 0;
}

int gvc_stage_peers(gvc_ctx *c, int stage, int n_peers, float *const *d_out_peers) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (stage != 0 && stage != 1) return fail(GVC_ERR_ARG, "only stages 0 and 1 write rows that other shards read");
    if (n_peers < 0 || n_peers > kMaxPeers) return fail(GVC_ERR_UNSUPPORTED, "%d peers; at most %d (8 GPUs)", n_peers, kMaxPeers);
    if (n_peers && !d_out_peers) return fail(GVC_ERR_ARG, "null peer table");
    PeerOut po{};
    for (int q = 0; q < n_peers; ++q) {
        if (!d_out_peers[q]) return fail(GVC_ERR_ARG, "peer %d is null", q);
        po.p[q] = d_out_peers[q];
    }
    po.n = n_peers;
    c->peers[stage] = po;
    return 0;
}

int gvc_peer_owners(gvc_ctx *c, int n_parts, const uint32_t *bounds, const int *peer_of_part) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (n_parts == 0) { c->parts = PartMap{}; c->have_peer_mask = false; return 0; }
    if (n_parts < 0 || n_parts > kMaxPeers + 1) return fail(GVC_ERR_UNSUPPORTED, "%d parts; at most %d", n_parts, kMaxPeers + 1);
    if (!bounds || !peer_of_part) return fail(GVC_ERR_ARG, "null argument");
    PartMap pm{};
    pm.n_parts = n_parts;
    for (int k = 0; k <= n_parts; ++k) {
        if (k && bounds[k] < bounds[k - 1]) return fail(GVC_ERR_ARG, "bounds not ascending");
        pm.bounds[k] = bounds[k];
    }
    for (int k = 0; k < n_parts; ++k) {
        if (peer_of_part[k] < -1 || peer_of_part[k] >= kMaxPeers) return fail(GVC_ERR_ARG, "peer index %d out of range", peer_of_part[k]);
        pm.peer_of_part[k] = peer_of_part[k];
    }
    c->parts = pm;
    if (!c->have_graph) return 0;
    if ((rc = use_device(c))) return rc;
    return build_peer_mask(c);
}

int gvc_stage_device(gvc_ctx *c, int stage, const float *d_in, float *d_out, float scale, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->fused) return fail(GVC_ERR_STATE, "stage kernels need the GNN_VC architecture (model not uploaded or not fused)");
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (c->n_local() && (!d_in || !d_out)) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    if (stage == 0) {                                // a forward begins: is this graph worth reordering by now?
        if ((rc = relabel_rows(c, false))) return rc;
        c->forwards_on_graph++;
    }
    switch (stage) {
    case 0:
        if (c->relabelled) {                         // x in the caller's numbering -> row order
            rows_in(c, d_in);
            return launch_stage<0>(c, c->d_xp.p, d_out, scale, mode);
        }
        return launch_stage<0>(c, d_in, d_out, scale, mode);
    case 1: return launch_stage<1>(c, d_in, d_out, scale, mode);
    case 2:
        if (c->relabelled) {                         // scores (and keys) in row order -> the caller's numbering
            if ((rc = launch_stage<2>(c, d_in, c->d_sp.p, scale, mode))) return rc;
            return rows_out(c, d_out);
        }
        return launch_stage<2>(c, d_in, d_out, scale, mode);
    default: return fail(GVC_ERR_ARG, "stage %d out of range", stage);
    }
}

int gvc_forward_device(gvc_ctx *c, const float *d_x, float scale, float *d_scores, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (c->layers.empty()) return fail(GVC_ERR_STATE, "no model uploaded");
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (c->v_begin != 0 || c->v_end != c->n_global)
        return fail(GVC_ERR_STATE, "gvc_forward_device needs a whole-graph context; shards go through gvc_stage_device");
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (c->n_global == 0) return 0;   // predict on an empty graph is a no-op
    if (!d_x || !d_scores) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    if (!c->fused) return forward_generic(c, d_x, scale, d_scores, mode);
    // graph_layer::WEIGHT_SCALE is per layer (src/gnn_inference.cpp:38-40): stage s divides by its own
    const float s0 = c->layer_scales.size() == 3 ? c->layer_scales[0] : scale;
    const float s1 = c->layer_scales.size() == 3 ? c->layer_scales[1] : scale;
    const float s2 = c->layer_scales.size() == 3 ? c->layer_scales[2] : scale;
    if ((rc = relabel_rows(c, false))) return rc;     // a graph that is forwarded again gets its rows reordered
    c->forwards_on_graph++;
    if (c->relabelled) {
        rows_in(c, d_x);
        if ((rc = launch_stage<0>(c, c->d_xp.p, c->d_h1.p, s0, mode))) return rc;
        if ((rc = launch_stage<1>(c, c->d_h1.p, c->d_h2.p, s1, mode))) return rc;
        if ((rc = launch_stage<2>(c, c->d_h2.p, c->d_sp.p, s2, mode))) return rc;
        return rows_out(c, d_scores);
    }
    if ((rc = launch_stage<0>(c, d_x, c->d_h1.p, s0, mode))) return rc;
    if ((rc = launch_stage<1>(c, c->d_h1.p, c->d_h2.p, s1, mode))) return rc;
    return launch_stage<2>(c, c->d_h2.p, d_scores, s2, mode);
}

int gvc_forward(gvc_ctx *c, const float *x, float scale, float *scores, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    const uint32_t n = c->n_global;
    if (n == 0) return 0;
    if (!scores) return fail(GVC_ERR_ARG, "null buffer");
    if (!x && c->x_resident_n != n) return fail(GVC_ERR_ARG, "x is null and no input came with the graph (gvc_graph_upload_stream_x)");
    if ((rc = use_device(c))) return rc;
    Tracer tr;
    // x goes up and the scores come back through the ring of pinned slots (the same one the streamed
    // graph upload uses): nothing of the graph's size is pinned per call.  x = NULL: it came with the graph.
    if ((rc = ensure_ring(c, 16, 256u << 10))) return rc;
    const size_t per = c->slot_bytes / sizeof(float);
    const size_t chunks = ((size_t)n + per - 1) / per;
    for (size_t k = 0; x && k < chunks; ++k) {
        const int slot = (int)(k % c->n_slots);
        const size_t off = k * per, cnt = std::min(per, (size_t)n - off);
        float *host = reinterpret_cast<float *>(c->ring.p + (size_t)slot * c->slot_bytes);
        GVC_CUDA(cudaEventSynchronize(c->slot_ev[slot]));
        std::memcpy(host, x + off, cnt * sizeof(float));
        GVC_CUDA(cudaMemcpyAsync(c->d_x.p + off, host, cnt * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        GVC_CUDA(cudaEventRecord(c->slot_ev[slot], c->stream));
    }
    if (x) c->x_resident_n = 0;
    tr.tick("forward: x up");
    // the selection keys are a by-product of stage 2 (5 bytes per vertex): kept on the device for gvc_last_keys
    const bool own_keys = c->fused && !c->keys_out;
    if (own_keys) {
        if ((rc = c->d_keys.reserve(n))) return rc;
        if ((rc = c->d_side.reserve(n))) return rc;
        c->keys_out = c->d_keys.p;
        c->side_out = c->d_side.p;
    }
    rc = gvc_forward_device(c, c->d_x.p, scale, c->d_scores.p, mode);
    if (own_keys) { c->keys_out = nullptr; c->side_out = nullptr; c->keys_valid_n = rc ? 0 : n; }
    if (rc) return rc;
    for (size_t k0 = 0; k0 < chunks; k0 += c->n_slots) {           // as many chunks as there are slots per round
        const size_t k1 = std::min(chunks, k0 + (size_t)c->n_slots);
        for (size_t k = k0; k < k1; ++k) {
            const int slot = (int)(k % c->n_slots);
            const size_t off = k * per, cnt = std::min(per, (size_t)n - off);
            GVC_CUDA(cudaMemcpyAsync(c->ring.p + (size_t)slot * c->slot_bytes, c->d_scores.p + off, cnt * sizeof(float),
                                     cudaMemcpyDeviceToHost, c->stream));
            GVC_CUDA(cudaEventRecord(c->slot_ev[slot], c->stream));
        }
        for (size_t k = k0; k < k1; ++k) {
            const int slot = (int)(k % c->n_slots);
            const size_t off = k * per, cnt = std::min(per, (size_t)n - off);
            GVC_CUDA(cudaEventSynchronize(c->slot_ev[slot]));
            std::memcpy(scores + off, c->ring.p + (size_t)slot * c->slot_bytes, cnt * sizeof(float));
        }
    }
    tr.tick("forward: kernels + scores down");
    return 0;
}

int gvc_forward_device_keys(gvc_ctx *c, const float *d_x, float scale, float *d_scores, float *d_keys,
                            unsigned char *d_side, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->fused) return fail(GVC_ERR_UNSUPPORTED, "selection keys come out of the fused stage-2 kernel (GNN_VC architecture only)");
    if (c->n_global && (!d_keys || !d_side)) return fail(GVC_ERR_ARG, "null buffer");
    c->keys_out = d_keys;
    c->side_out = d_side;
    rc = gvc_forward_device(c, d_x, scale, d_scores, mode);
    c->keys_out = nullptr;
    c->side_out = nullptr;
    return rc;
}

int gvc_forward_keys(gvc_ctx *c, const float *x, float scale, float *scores, float *keys, unsigned char *side, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->fused) return fail(GVC_ERR_UNSUPPORTED, "selection keys come out of the fused stage-2 kernel (GNN_VC architecture only)");
    if ((rc = gvc_forward(c, x, scale, scores, mode))) return rc;
    return gvc_last_keys(c, keys, side);
}

int gvc_last_keys(gvc_ctx *c, float *keys, unsigned char *side) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->have_graph || c->keys_valid_n != c->n_global) return fail(GVC_ERR_STATE, "no forward of the current graph has left selection keys");
    if (!c->n_global) return 0;
    if (!keys || !side) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    GVC_CUDA(cudaMemcpyAsync(keys, c->d_keys.p, (size_t)c->n_global * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaMemcpyAsync(side, c->d_side.p, (size_t)c->n_global, cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int gvc_graph_layer_device(gvc_ctx *c, const float *d_in, int width, float *d_out, float scale) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (width <= 0) return fail(GVC_ERR_ARG, "width must be positive");
    if ((rc = use_device(c))) return rc;
    const uint32_t nl = c->n_local();
    if (!nl) return 0;
    const uint64_t work = (uint64_t)nl * (2 * width + 3);
    generic_graph_kernel<<<blocks_for(work, 256), 256, 0, c->stream>>>(c->row_ptr, c->col, c->Wv, c->NWv, d_in, width,
                                                                       d_out, nl, c->v_begin, scale);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int gvc_linear_layer_device(gvc_ctx *c, int layer_index, uint64_t n, const float *d_in, float *d_out) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (layer_index < 0 || layer_index >= (int)c->layers.size() || c->layers[layer_index].kind != GVC_LINEAR)
        return fail(GVC_ERR_ARG, "layer %d is not a linear layer", layer_index);
    if ((rc = use_device(c))) return rc;
    if (!n) return 0;
    auto &l = c->layers[layer_index];
    generic_linear_kernel<true><<<blocks_for(n * l.cols, 256), 256, 0, c->stream>>>(d_in, l.rows, l.cols, l.dW, l.db, d_out, n);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int gvc_relu_device(gvc_ctx *c, uint64_t count, const float *d_in, float *d_out) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = use_device(c))) return rc;
    if (!count) return 0;
    generic_relu_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(d_in, d_out, count);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int gvc_sigmoid_device(gvc_ctx *c, uint64_t count, const float *d_in, float *d_out, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = use_device(c))) return rc;
    if (!count) return 0;
    if (mode == GVC_MODE_EXACT)
        generic_sigmoid_kernel<true><<<blocks_for(count, 256), 256, 0, c->stream>>>(d_in, d_out, count);
    else
        generic_sigmoid_kernel<false><<<blocks_for(count, 256), 256, 0, c->stream>>>(d_in, d_out, count);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

// ---- host-buffer single layers ------------------------------------------------------
namespace {
struct Scratch2 {   // two device scratch areas for the host-buffer layer calls
    float *a = nullptr, *b = nullptr;
};
int host_io_begin(gvc_ctx *c, size_t in_floats, size_t out_floats, const float *in, Scratch2 *s) {
    int rc;
    if ((rc = c->d_ping.reserve(in_floats))) return rc;
    if ((rc = c->d_pong.reserve(out_floats))) return rc;
    s->a = c->d_ping.p; s->b = c->d_pong.p;
    if (in_floats) GVC_CUDA(cudaMemcpyAsync(s->a, in, in_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return 0;
}
int host_io_end(gvc_ctx *c, size_t out_floats, float *out, const Scratch2 &s) {
    if (out_floats) GVC_CUDA(cudaMemcpyAsync(out, s.b, out_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
}  // namespace

int gvc_graph_layer_host(gvc_ctx *c, const float *in, int width, float *out, float scale) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!c->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (c->v_begin != 0 || c->v_end != c->n_global) return fail(GVC_ERR_STATE, "whole-graph context required");
    if (width <= 0) return fail(GVC_ERR_ARG, "width must be positive");
    const size_t n = c->n_global;
    if (!n) return 0;
    if (!in || !out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    Scratch2 s;
    if ((rc = host_io_begin(c, n * width, n * (2 * (size_t)width + 3), in, &s))) return rc;
    if ((rc = gvc_graph_layer_device(c, s.a, width, s.b, scale))) return rc;
    return host_io_end(c, n * (2 * (size_t)width + 3), out, s);
}

int gvc_linear_host(gvc_ctx *c, uint64_t n, int K, int Nout, const float *in, const float *W,
                    const float *bias, float *out, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (K <= 0 || Nout <= 0) return fail(GVC_ERR_ARG, "bad layer shape %d x %d", K, Nout);
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (!n) return 0;
    if (!in || !W || !bias || !out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    Scratch2 s;
    const size_t wf = (size_t)K * Nout + Nout;
    if ((rc = host_io_begin(c, n * K + wf, n * Nout, in, &s))) return rc;
    float *dW = s.a + n * K, *db = dW + (size_t)K * Nout;
    GVC_CUDA(cudaMemcpyAsync(dW, W, (size_t)K * Nout * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    GVC_CUDA(cudaMemcpyAsync(db, bias, (size_t)Nout * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (mode == GVC_MODE_EXACT)
        generic_linear_kernel<true><<<blocks_for(n * Nout, 256), 256, 0, c->stream>>>(s.a, K, Nout, dW, db, s.b, n);
    else
        generic_linear_kernel<false><<<blocks_for(n * Nout, 256), 256, 0, c->stream>>>(s.a, K, Nout, dW, db, s.b, n);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    return host_io_end(c, n * Nout, out, s);
}

int gvc_relu_host(gvc_ctx *c, uint64_t count, const float *in, float *out) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!count) return 0;
    if (!in || !out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    Scratch2 s;
    if ((rc = host_io_begin(c, count, count, in, &s))) return rc;
    if ((rc = gvc_relu_device(c, count, s.a, s.b))) return rc;
    return host_io_end(c, count, out, s);
}

int gvc_sigmoid_host(gvc_ctx *c, uint64_t count, const float *in, float *out, int mode) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    if (!count) return 0;
    if (!in || !out) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    Scratch2 s;
    if ((rc = host_io_begin(c, count, count, in, &s))) return rc;
    if ((rc = gvc_sigmoid_device(c, count, s.a, s.b, mode))) return rc;
    return host_io_end(c, count, out, s);
}

int gvc_sgemm_host(gvc_ctx *c, int ta, int tb, uint64_t m, uint64_t n, uint64_t k, const float *A,
                   uint64_t lda, const float *B, uint64_t ldb, float beta, float *Cm, uint64_t ldc) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!m || !n) return 0;
    if ((k && (!A || !B)) || !Cm) return fail(GVC_ERR_ARG, "null buffer");
    if ((rc = use_device(c))) return rc;
    const size_t a_floats = (ta ? k : m) * lda, b_floats = (tb ? n : k) * ldb, c_floats = m * ldc;
    if ((rc = c->d_ping.reserve(a_floats + b_floats))) return rc;
    if ((rc = c->d_pong.reserve(c_floats))) return rc;
    float *dA = c->d_ping.p, *dB = dA + a_floats, *dC = c->d_pong.p;
    if (a_floats) GVC_CUDA(cudaMemcpyAsync(dA, A, a_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (b_floats) GVC_CUDA(cudaMemcpyAsync(dB, B, b_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (beta != 0.0f) GVC_CUDA(cudaMemcpyAsync(dC, Cm, c_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    generic_sgemm_kernel<<<blocks_for(m * n, 256), 256, 0, c->stream>>>(ta, tb, m, n, k, dA, lda, dB, ldb, beta, dC, ldc);
    GVC_CUDA(cudaGetLastError());
    c->launches++;
    GVC_CUDA(cudaMemcpyAsync(Cm, dC, c_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ================================ several GPUs, one process =======================================
// SURVEY.md 8(e) behind one call: vertex-range shards over the devices of a group, every shard a
// gvc_ctx of its own.  The row exchange is the stage kernels' own (each finished row is stored into
// the h buffers of the devices that own a neighbour, over NVLink peer access), and what remains
// between two stages -- "every device has finished its stage" -- is expressed with CUDA events that
// the streams wait on; the host thread never blocks inside a forward and no collective library is
// involved.
struct gvc_group {
    std::vector<gvc_ctx *> ctx;
    std::vector<uint32_t> bounds;            // vertex ranges of the shards: [bounds[d], bounds[d + 1])
    std::vector<cudaEvent_t> done;           // per device: "my stage kernel (and its peer stores) has finished"
    std::vector<uint64_t> rp_tmp;            // host scratch for a shard's rebased offsets
    uint32_t n = 0;
    bool have_graph = false;
};

namespace {
int group_check(const gvc_group *g) {
    if (!g || g->ctx.empty()) return fail(GVC_ERR_ARG, "null group");
    return 0;
}
}  // namespace

int gvc_group_create(gvc_group **out, const int *devices, int n_devices) {
    if (!out || !devices) return fail(GVC_ERR_ARG, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > kMaxPeers + 1) return fail(GVC_ERR_UNSUPPORTED, "%d devices; 1..%d", n_devices, kMaxPeers + 1);
    gvc_group *g = new (std::nothrow) gvc_group();
    if (!g) return fail(GVC_ERR_ALLOC, "out of host memory");
    int rc = 0;
    for (int d = 0; d < n_devices && !rc; ++d) {
        gvc_ctx *c = nullptr;
        if ((rc = gvc_ctx_create(&c, devices[d]))) break;
        g->ctx.push_back(c);
    }
    // every device stores rows into every other device's buffers: peer access both ways
    for (int a = 0; a < (int)g->ctx.size() && !rc; ++a)
        for (int b = 0; b < (int)g->ctx.size() && !rc; ++b) {
            const int da = g->ctx[a]->device, db = g->ctx[b]->device;
            if (da == db) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, da, db);
            if (!can) { rc = fail(GVC_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", da, db); break; }
            cudaSetDevice(da);
            const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(1000 + (int)e, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
    for (size_t d = 0; d < g->ctx.size() && !rc; ++d) {
        cudaEvent_t ev = nullptr;
        cudaSetDevice(g->ctx[d]->device);
        const cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) rc = fail(1000 + (int)e, "cudaEventCreate: %s", cudaGetErrorString(e));
        else g->done.push_back(ev);
    }
    if (rc) { gvc_group_destroy(g); return rc; }
    *out = g;
    return 0;
}

void gvc_group_destroy(gvc_group *g) {
    if (!g) return;
    for (auto *c : g->ctx) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    for (size_t d = 0; d < g->done.size(); ++d) { cudaSetDevice(g->ctx[d]->device); cudaEventDestroy(g->done[d]); }
    for (auto *c : g->ctx) gvc_ctx_destroy(c);
    delete g;
}

int gvc_group_size(const gvc_group *g) { return g ? (int)g->ctx.size() : 0; }

int gvc_group_model_upload(gvc_group *g, int n_layers, const int *kinds, const int *rows, const int *cols,
                           const float *const *W, const float *const *bias) {
    int rc;
    if ((rc = group_check(g))) return rc;
    for (auto *c : g->ctx)
        if ((rc = gvc_model_upload(c, n_layers, kinds, rows, cols, W, bias))) return rc;
    if (!g->ctx[0]->fused) return fail(GVC_ERR_UNSUPPORTED, "sharded forwards need the GNN_VC architecture (fused stage kernels)");
    return 0;
}

int gvc_group_model_weight_scales(gvc_group *g, int n_graph_layers, const float *scales) {
    int rc;
    if ((rc = group_check(g))) return rc;
    for (auto *c : g->ctx)
        if ((rc = gvc_model_weight_scales(c, n_graph_layers, scales))) return rc;
    return 0;
}

// Whole graph in, shards out: contiguous vertex ranges with about the same number of adjacency entries
// plus a per-vertex cost (the dense chain), boundaries at multiples of 32.
int gvc_group_graph_upload(gvc_group *g, uint32_t n, const uint64_t *row_ptr, const uint32_t *col, const uint32_t *W,
                           const uint32_t *NW) {
    int rc;
    if ((rc = group_check(g))) return rc;
    if (n && (!row_ptr || !W || !NW)) return fail(GVC_ERR_ARG, "null graph arrays");
    const int P = (int)g->ctx.size();
    g->have_graph = false;
    g->n = n;
    g->bounds.assign(P + 1, n);
    g->bounds[0] = 0;
    const uint64_t nnz = n ? row_ptr[n] : 0;
    const uint64_t per_vertex = 24;                                   // a vertex costs about as much as 24 adjacency entries
    const double total = (double)nnz + (double)per_vertex * n;
    uint32_t u = 0;
    for (int d = 1; d < P; ++d) {
        const double want = total * d / P;
        // first vertex whose prefix cost reaches `want` (binary search on row_ptr[u] + per_vertex * u)
        uint32_t lo = u, hi = n;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if ((double)row_ptr[mid] + (double)per_vertex * mid < want) lo = mid + 1; else hi = mid;
        }
        u = std::min<uint32_t>(n, (lo + 31) / 32 * 32);
        g->bounds[d] = std::max(u, g->bounds[d - 1]);
    }
    // peers: device d mirrors its rows into the buffers of every other device that owns a neighbour
    for (int d = 0; d < P; ++d) {
        gvc_ctx *c = g->ctx[d];
        const uint32_t vb = g->bounds[d], ve = g->bounds[d + 1], nl = ve - vb;
        g->rp_tmp.resize((size_t)nl + 1);
        for (uint32_t i = 0; i <= nl; ++i) g->rp_tmp[i] = row_ptr[vb + i] - row_ptr[vb];
        if ((rc = gvc_graph_upload_shard(c, n, vb, ve, g->rp_tmp.data(), col ? col + row_ptr[vb] : nullptr, W + vb, NW + vb))) return rc;
    }
    for (int d = 0; d < P && P > 1; ++d) {
        gvc_ctx *c = g->ctx[d];
        float *p1[kMaxPeers], *p2[kMaxPeers];
        int peer_of_part[kMaxPeers + 1];
        int q = 0;
        for (int o = 0; o < P; ++o) {
            if (o == d) { peer_of_part[o] = -1; continue; }
            p1[q] = g->ctx[o]->d_h1.p;
            p2[q] = g->ctx[o]->d_h2.p;
            peer_of_part[o] = q++;
        }
        if ((rc = gvc_stage_peers(c, 0, q, p1))) return rc;
        if ((rc = gvc_stage_peers(c, 1, q, p2))) return rc;
        if ((rc = gvc_peer_owners(c, P, g->bounds.data(), peer_of_part))) return rc;
    }
    g->have_graph = true;
    return 0;
}

int gvc_group_bounds(const gvc_group *g, uint32_t *bounds_out) {
    if (!g || !bounds_out) return fail(GVC_ERR_ARG, "null argument");
    for (size_t i = 0; i < g->bounds.size(); ++i) bounds_out[i] = g->bounds[i];
    return 0;
}

// gnn::model::predict over the group: x replicated, three stages with "all devices done" between them,
// every device's score slice copied into scores[bounds[d] ..).
int gvc_group_forward(gvc_group *g, const float *x, float scale, float *scores, int mode) {
    int rc;
    if ((rc = group_check(g))) return rc;
    if (!g->have_graph) return fail(GVC_ERR_STATE, "no graph uploaded");
    if (mode != GVC_MODE_EXACT && mode != GVC_MODE_FAST) return fail(GVC_ERR_ARG, "bad mode %d", mode);
    const uint32_t n = g->n;
    if (!n) return 0;
    if (!x || !scores) return fail(GVC_ERR_ARG, "null buffer");
    const int P = (int)g->ctx.size();
    for (auto *c : g->ctx) {
        if ((rc = use_device(c))) return rc;
        GVC_CUDA(cudaMemcpyAsync(c->d_x.p, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    auto all_done = [&]() -> int {          // every stream waits until every device's stage (and its peer stores) has finished
        for (int d = 0; d < P; ++d) {
            GVC_CUDA(cudaSetDevice(g->ctx[d]->device));
            GVC_CUDA(cudaEventRecord(g->done[d], g->ctx[d]->stream));
        }
        for (int d = 0; d < P; ++d) {
            GVC_CUDA(cudaSetDevice(g->ctx[d]->device));
            for (int o = 0; o < P; ++o)
                if (o != d) GVC_CUDA(cudaStreamWaitEvent(g->ctx[d]->stream, g->done[o], 0));
        }
        return 0;
    };
    // nobody may still be reading h1/h2 of an earlier forward when the first peer stores arrive
    if ((rc = all_done())) return rc;
    for (int stage = 0; stage < 3; ++stage) {
        for (auto *c : g->ctx) {
            const float s = c->layer_scales.size() == 3 ? c->layer_scales[stage] : scale;
            const float *in = stage == 0 ? c->d_x.p : stage == 1 ? c->d_h1.p : c->d_h2.p;
            float *out = stage == 0 ? c->d_h1.p : stage == 1 ? c->d_h2.p : c->d_scores.p;
            if ((rc = gvc_stage_device(c, stage, in, out, s, mode))) return rc;
        }
        if (stage < 2 && (rc = all_done())) return rc;
    }
    for (int d = 0; d < P; ++d) {
        gvc_ctx *c = g->ctx[d];
        const uint32_t vb = g->bounds[d], nl = g->bounds[d + 1] - vb;
        if (!nl) continue;
        GVC_CUDA(cudaSetDevice(c->device));
        GVC_CUDA(cudaMemcpyAsync(scores + vb, c->d_scores.p, (size_t)nl * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    for (auto *c : g->ctx) {
        GVC_CUDA(cudaSetDevice(c->device));
        GVC_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

void *gvc_stream(gvc_ctx *c) { return c ? (void *)c->stream : nullptr; }

int gvc_set_stream(gvc_ctx *c, void *stream) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if ((rc = use_device(c))) return rc;
    GVC_CUDA(cudaStreamSynchronize(c->stream));          // nothing of ours may still be in flight on the old one
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return 0;
}

int gvc_sync(gvc_ctx *c) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

uint64_t gvc_launch_count(const gvc_ctx *c) { return c ? c->launches : 0; }

int gvc_debug_px(gvc_ctx *c, uint32_t *out4) {
    int rc;
    if ((rc = check_ctx(c))) return rc;
    if (!out4) return fail(GVC_ERR_ARG, "null buffer");
    out4[0] = c->sched.n_px; out4[1] = c->sched.n_chunks_px;
    for (int i = 2; i < 8; ++i) out4[i] = 0;
    if (!c->have_graph || !c->sched.n_px) return 0;
    if ((rc = use_device(c))) return rc;
    GVC_CUDA(cudaStreamSynchronize(c->stream));
    GVC_CUDA(cudaMemcpy(out4 + 2, c->d_sync.p + c->px_ctr_off + 2, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

// which vertex every row of the stage outputs belongs to (the identity unless the context keeps its rows in
// schedule order, relabel_rows); returns 1 if the rows are renumbered, 0 if not, < 0 on error
int gvc_debug_row_order(gvc_ctx *c, uint32_t *vertex_of_row) {
    if (!c || !vertex_of_row) return -1;
    if (!c->have_graph) return -1;
    const uint32_t n = c->n_global;
    if (!c->relabelled) {
        for (uint32_t i = 0; i < n; ++i) vertex_of_row[i] = i;
        return 0;
    }
    if (use_device(c)) return -1;
    if (cudaMemcpyAsync(vertex_of_row, c->d_row_order.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    return 1;
}

const float *gvc_debug_h(const gvc_ctx *c, int which) {
    if (!c) return nullptr;
    return which == 0 ? c->d_h1.p : c->d_h2.p;
}

}  // extern "C"

#include "gvc_train_api.cuh"
