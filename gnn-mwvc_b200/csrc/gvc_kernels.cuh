// gvc_kernels.cuh -- sm_100a kernels of the GNN_VC forward (hot path of
// gnn::model::predict, reference src/gnn_inference.cpp:67-81).
//
// One "stage" = one graph layer (src/gnn_inference.cpp:27-42) fused with the
// dense layers, bias adds and activations that follow it (:20-25, :44-52) up to
// the next graph layer.  A warp owns a tile of 32 vertices from gather to
// store; nothing but the 16-float stage output per vertex goes back to HBM.
//
//   phase A  gather-aggregate: neighbour rows are summed in adjacency order
//            (the order the reference adds them, :33-36) with 128-bit loads,
//            4 lanes per 64-byte row; the concatenated feature vector of
//            :37-40 -- including its column quirk, SURVEY.md A.2 -- is written
//            k-major into the warp's shared-memory tile, never to HBM.
//   phase B  dense chain on CUDA cores: each lane keeps a 4-vertex x 8-output
//            (or 4 x 4) block of accumulators in registers, reads activations
//            and weights from shared memory with 128-bit loads, applies bias
//            and ReLU in registers and writes the next activation tile in place.
//   phase C  store: 16 floats per vertex (stages 0/1) or the sigmoid score.
//
// EXACT=true keeps the reference's fp32 operation order (one accumulator per
// output, k ascending, product and sum rounded separately, see
// oracle/gnn_oracle.c) -> bit-identical scores.  EXACT=false contracts to FFMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gvc_expf.h"

namespace gvc {

constexpr int kWarpsPerCta = 8;
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr int kTileVerts = 32;            // vertices per warp tile
constexpr int kTileStride = 36;           // floats per k-row of a tile (32 + 4 pad, keeps float4 alignment)
constexpr int kTileRows = 32;             // widest activation
constexpr int kTileFloats = kTileRows * kTileStride;

// Packed parameter block of one stage, in floats:
//   [W_a (Ka x Na)] [b_a (Na)] [W_b (Kb x Nb)] [b_b (Nb)] [W_c (Kc x Nc)] [b_c (Nc)]
// For the 35-row matrices only rows 0..31 are kept: features 32..34 are +0.0 by
// construction (:29) and an accumulator that starts at +0.0 is never -0.0, so
// adding their +-0 products cannot change any bit (weights must be finite; the
// host checks).
struct StageDims {
    int Ka, Na, Kb, Nb, Kc, Nc;
    __host__ __device__ constexpr int floats() const { return Ka * Na + Na + Kb * Nb + Nb + Kc * Nc + Nc; }
};
__host__ __device__ constexpr StageDims stage_dims(int stage) {
    return stage == 0 ? StageDims{5, 32, 32, 32, 32, 16}
         : stage == 1 ? StageDims{32, 32, 32, 32, 32, 16}
                      : StageDims{32, 32, 32, 16, 16, 1};
}

__device__ __forceinline__ float relu_ref(float v) { return v < 0.0f ? 0.0f : v; }   // std::max(x, 0.0f), :46

template <bool EXACT>
__device__ __forceinline__ float mac(float a, float w, float acc) {
    if constexpr (EXACT) return __fadd_rn(acc, __fmul_rn(a, w));
    else return fmaf(a, w, acc);
}

__device__ __forceinline__ float4 ldg_row4(const float4 *p) { return __ldg(p); }

// ---- phase B: one dense layer + bias + ReLU over the warp's tile, in place -----
// T: k-major tile, T[k * kTileStride + i], i = vertex in tile.  Lane (og, vg)
// owns vertices 4vg..4vg+3 and outputs C*og..C*og+C-1.
template <int K, int NOUT, bool EXACT>
__device__ __forceinline__ void tile_linear_relu(float *__restrict__ T, const float *__restrict__ Wsm,
                                                 const float *__restrict__ bsm, int lane) {
    constexpr int C = NOUT / 4;
    static_assert(C == 8 || C == 4, "NOUT must be 32 or 16");
    const int og = lane >> 3, vg = lane & 7;
    float acc[4][C];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[r][c] = 0.0f;

#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4 *>(T + k * kTileStride + 4 * vg);
        float w[C];
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 ww = *reinterpret_cast<const float4 *>(Wsm + k * NOUT + C * og + 4 * c4);
            w[4 * c4 + 0] = ww.x; w[4 * c4 + 1] = ww.y; w[4 * c4 + 2] = ww.z; w[4 * c4 + 3] = ww.w;
        }
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[r][c] = mac<EXACT>(av[r], w[c], acc[r][c]);
    }
    __syncwarp();   // every lane is done reading the input tile
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float b = bsm[C * og + c];
        float4 o;
        o.x = relu_ref(__fadd_rn(acc[0][c], b));
        o.y = relu_ref(__fadd_rn(acc[1][c], b));
        o.z = relu_ref(__fadd_rn(acc[2][c], b));
        o.w = relu_ref(__fadd_rn(acc[3][c], b));
        *reinterpret_cast<float4 *>(T + (C * og + c) * kTileStride + 4 * vg) = o;
    }
    __syncwarp();
}

// Last dense layer of stages 0/1 (K=32 -> 16) + bias + ReLU, stored straight
// from registers to the stage output rows (64 B per vertex, full sectors).
template <int K, bool EXACT>
__device__ __forceinline__ void tile_linear_relu_store16(const float *__restrict__ T,
                                                         const float *__restrict__ Wsm,
                                                         const float *__restrict__ bsm, int lane,
                                                         float *__restrict__ out_rows /* row of tile vertex 0 */,
                                                         int valid /* vertices of the tile that exist */) {
    const int og = lane >> 3, vg = lane & 7;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4 *>(T + k * kTileStride + 4 * vg);
        const float4 ww = *reinterpret_cast<const float4 *>(Wsm + k * 16 + 4 * og);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float w[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = mac<EXACT>(av[r], w[c], acc[r][c]);
    }
    const float4 b = *reinterpret_cast<const float4 *>(bsm + 4 * og);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = 4 * vg + r;
        if (i < valid) {
            float4 o;
            o.x = relu_ref(__fadd_rn(acc[r][0], b.x));
            o.y = relu_ref(__fadd_rn(acc[r][1], b.y));
            o.z = relu_ref(__fadd_rn(acc[r][2], b.z));
            o.w = relu_ref(__fadd_rn(acc[r][3], b.w));
            *reinterpret_cast<float4 *>(out_rows + (size_t)i * 16 + 4 * og) = o;
        }
    }
}

// sigmoid::forward :49-52
template <bool EXACT>
__device__ __forceinline__ float sigmoid_ref(float v) {
    if constexpr (EXACT) return __fdiv_rn(1.0f, __fadd_rn(1.0f, gvc_expf_glibc(-v)));
    else return 1.0f / (1.0f + expf(-v));
}

// ---- phase A, width 16: 4 lanes per vertex, 8 vertices per pass ---------------
template <bool EXACT>
__device__ __forceinline__ void gather16_tile(float *__restrict__ T, const uint32_t *__restrict__ row_ptr,
                                              const uint32_t *__restrict__ col,
                                              const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                              const float4 *__restrict__ in4, uint32_t tile_base,
                                              uint32_t n_local, uint32_t v_begin, float scale, int lane) {
    const int sv = lane >> 2, q = lane & 3;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        const int i = p * 8 + sv;
        const uint32_t ul = tile_base + i;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 self = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ul < n_local) {
            uint32_t e = row_ptr[ul];
            const uint32_t end = row_ptr[ul + 1];
            const uint32_t deg = end - e;
            for (; e + 4 <= end; e += 4) {
                const uint32_t v0 = __ldg(col + e), v1 = __ldg(col + e + 1), v2 = __ldg(col + e + 2),
                               v3 = __ldg(col + e + 3);
                const float4 r0 = ldg_row4(in4 + (size_t)v0 * 4 + q);
                const float4 r1 = ldg_row4(in4 + (size_t)v1 * 4 + q);
                const float4 r2 = ldg_row4(in4 + (size_t)v2 * 4 + q);
                const float4 r3 = ldg_row4(in4 + (size_t)v3 * 4 + q);
                acc.x = __fadd_rn(acc.x, r0.x); acc.y = __fadd_rn(acc.y, r0.y); acc.z = __fadd_rn(acc.z, r0.z); acc.w = __fadd_rn(acc.w, r0.w);
                acc.x = __fadd_rn(acc.x, r1.x); acc.y = __fadd_rn(acc.y, r1.y); acc.z = __fadd_rn(acc.z, r1.z); acc.w = __fadd_rn(acc.w, r1.w);
                acc.x = __fadd_rn(acc.x, r2.x); acc.y = __fadd_rn(acc.y, r2.y); acc.z = __fadd_rn(acc.z, r2.z); acc.w = __fadd_rn(acc.w, r2.w);
                acc.x = __fadd_rn(acc.x, r3.x); acc.y = __fadd_rn(acc.y, r3.y); acc.z = __fadd_rn(acc.z, r3.z); acc.w = __fadd_rn(acc.w, r3.w);
            }
            for (; e < end; ++e) {
                const uint32_t v = __ldg(col + e);
                const float4 r = ldg_row4(in4 + (size_t)v * 4 + q);
                acc.x = __fadd_rn(acc.x, r.x); acc.y = __fadd_rn(acc.y, r.y); acc.z = __fadd_rn(acc.z, r.z); acc.w = __fadd_rn(acc.w, r.w);
            }
            self = ldg_row4(in4 + (size_t)(v_begin + ul) * 4 + q);                     // :37
            if (q == 0) {   // the quirk: D, W/s, NW/s overwrite self features 1..3 (:38-40)
                self.y = __uint2float_rn(deg);
                self.z = __fdiv_rn(__uint2float_rn(__ldg(Wv + ul)), scale);
                self.w = __fdiv_rn(__uint2float_rn(__ldg(NWv + ul)), scale);
            }
        }
        float *t = T + (4 * q) * kTileStride + i;
        t[0] = acc.x; t[kTileStride] = acc.y; t[2 * kTileStride] = acc.z; t[3 * kTileStride] = acc.w;
        t += 16 * kTileStride;
        t[0] = self.x; t[kTileStride] = self.y; t[2 * kTileStride] = self.z; t[3 * kTileStride] = self.w;
    }
    __syncwarp();
}

// ---- phase A, width 1: one lane per vertex -------------------------------------
__device__ __forceinline__ void gather1_tile(float *__restrict__ T, const uint32_t *__restrict__ row_ptr,
                                             const uint32_t *__restrict__ col,
                                             const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                             const float *__restrict__ x, uint32_t tile_base,
                                             uint32_t n_local, uint32_t v_begin, float scale, int lane) {
    const uint32_t ul = tile_base + lane;
    float agg = 0.0f, xs = 0.0f, fd = 0.0f, fw = 0.0f, fnw = 0.0f;
    if (ul < n_local) {
        uint32_t e = row_ptr[ul];
        const uint32_t end = row_ptr[ul + 1];
        fd = __uint2float_rn(end - e);
        for (; e + 4 <= end; e += 4) {
            const uint32_t v0 = __ldg(col + e), v1 = __ldg(col + e + 1), v2 = __ldg(col + e + 2),
                           v3 = __ldg(col + e + 3);
            const float a0 = __ldg(x + v0), a1 = __ldg(x + v1), a2 = __ldg(x + v2), a3 = __ldg(x + v3);
            agg = __fadd_rn(agg, a0); agg = __fadd_rn(agg, a1); agg = __fadd_rn(agg, a2); agg = __fadd_rn(agg, a3);
        }
        for (; e < end; ++e) agg = __fadd_rn(agg, __ldg(x + __ldg(col + e)));
        xs = __ldg(x + v_begin + ul);
        fw = __fdiv_rn(__uint2float_rn(__ldg(Wv + ul)), scale);
        fnw = __fdiv_rn(__uint2float_rn(__ldg(NWv + ul)), scale);
    }
    T[0 * kTileStride + lane] = agg;    // [agg | x | D | W/s | NW/s], :33-40 with w = 1
    T[1 * kTileStride + lane] = xs;
    T[2 * kTileStride + lane] = fd;
    T[3 * kTileStride + lane] = fw;
    T[4 * kTileStride + lane] = fnw;
    __syncwarp();
}

// ---- the fused stage kernel ------------------------------------------------------
// STAGE 0: in = x [n_global], out = h rows [n_global x 16]
// STAGE 1: in = h [n_global x 16], out = h rows [n_global x 16]
// STAGE 2: in = h [n_global x 16], out = scores [n_local]
template <int STAGE, bool EXACT>
__global__ void __launch_bounds__(kCtaThreads, 3)
stage_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
             const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
             const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ params,
             uint32_t n_local, uint32_t v_begin, float scale) {
    constexpr StageDims D = stage_dims(STAGE);
    extern __shared__ __align__(16) float smem[];
    float *P = smem;                                             // packed parameters
    constexpr int kParamFloats = (D.floats() + 3) / 4 * 4;
    float *tiles = smem + kParamFloats;

    for (int i = threadIdx.x; i < D.floats(); i += kCtaThreads) P[i] = __ldg(params + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile_base = (blockIdx.x * kWarpsPerCta + warp) * kTileVerts;
    if (tile_base >= n_local) return;
    float *T = tiles + warp * kTileFloats;

    const float *Wa = P, *ba = Wa + D.Ka * D.Na;
    const float *Wb = ba + D.Na, *bb = Wb + D.Kb * D.Nb;
    const float *Wc = bb + D.Nb, *bc = Wc + D.Kc * D.Nc;

    if constexpr (STAGE == 0)
        gather1_tile(T, row_ptr, col, Wv, NWv, in, tile_base, n_local, v_begin, scale, lane);
    else
        gather16_tile<EXACT>(T, row_ptr, col, Wv, NWv, reinterpret_cast<const float4 *>(in), tile_base,
                             n_local, v_begin, scale, lane);

    const int valid = (int)min((uint32_t)kTileVerts, n_local - tile_base);
    if constexpr (STAGE < 2) {
        tile_linear_relu<D.Ka, D.Na, EXACT>(T, Wa, ba, lane);
        tile_linear_relu<D.Kb, D.Nb, EXACT>(T, Wb, bb, lane);
        tile_linear_relu_store16<D.Kc, EXACT>(T, Wc, bc, lane,
                                              out + (size_t)(v_begin + tile_base) * 16, valid);
    } else {
        tile_linear_relu<D.Ka, D.Na, EXACT>(T, Wa, ba, lane);
        tile_linear_relu<D.Kb, D.Nb, EXACT>(T, Wb, bb, lane);
        // 16 -> 1: one lane per vertex.  OpenBLAS' 1-column kernel: even/odd
        // accumulators, C = even + odd (oracle/gnn_oracle.c dot_two_acc).
        float ev = 0.0f, od = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            ev = mac<EXACT>(T[k * kTileStride + lane], Wc[k], ev);
            od = mac<EXACT>(T[(k + 1) * kTileStride + lane], Wc[k + 1], od);
        }
        const float s = __fadd_rn(__fadd_rn(ev, od), bc[0]);
        if (lane < valid) out[tile_base + lane] = sigmoid_ref<EXACT>(s);
    }
}

template <int STAGE>
constexpr size_t stage_smem_bytes() {
    constexpr StageDims D = stage_dims(STAGE);
    return ((D.floats() + 3) / 4 * 4 + kWarpsPerCta * kTileFloats) * sizeof(float);
}

// ---- exact-mode tail: the last vertex of an odd-sized graph ---------------------
// OpenBLAS' sgemm takes rows 4 at a time, then 2, then 1; its 1-row kernel sums
// in a different order (two accumulators, see oracle/gnn_oracle.c).  One warp
// redoes the stage for that single vertex in that order.  lane j = output column.
template <int STAGE>
__global__ void __launch_bounds__(32)
stage_tail_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                  const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                  const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ params,
                  uint32_t ul, uint32_t v_begin, float scale) {
    constexpr StageDims D = stage_dims(STAGE);
    __shared__ float f[2][32];
    const int lane = threadIdx.x;
    const uint32_t beg = row_ptr[ul], end = row_ptr[ul + 1];
    f[0][lane] = 0.0f;
    __syncwarp();
    if constexpr (STAGE == 0) {
        if (lane == 0) {
            float agg = 0.0f;
            for (uint32_t e = beg; e < end; ++e) agg = __fadd_rn(agg, in[col[e]]);
            f[0][0] = agg;
            f[0][1] = in[v_begin + ul];
            f[0][2] = __uint2float_rn(end - beg);
            f[0][3] = __fdiv_rn(__uint2float_rn(Wv[ul]), scale);
            f[0][4] = __fdiv_rn(__uint2float_rn(NWv[ul]), scale);
        }
    } else {
        if (lane < 16) {
            float agg = 0.0f;
            for (uint32_t e = beg; e < end; ++e) agg = __fadd_rn(agg, in[(size_t)col[e] * 16 + lane]);
            f[0][lane] = agg;
            float s = in[(size_t)(v_begin + ul) * 16 + lane];
            if (lane == 1) s = __uint2float_rn(end - beg);
            if (lane == 2) s = __fdiv_rn(__uint2float_rn(Wv[ul]), scale);
            if (lane == 3) s = __fdiv_rn(__uint2float_rn(NWv[ul]), scale);
            f[0][16 + lane] = s;
        }
    }
    __syncwarp();
    const float *Wa = params, *ba = Wa + D.Ka * D.Na;
    const float *Wb = ba + D.Na, *bb = Wb + D.Kb * D.Nb;
    const float *Wc = bb + D.Nb, *bc = Wc + D.Kc * D.Nc;
    auto dot2 = [&](const float *a, const float *Wm, int K, int N, int j) {
        float ev = 0.0f, od = 0.0f;
        const int K8 = K / 8 * 8;
        int k = 0;
        for (; k < K8; k += 2) {
            ev = __fadd_rn(ev, __fmul_rn(a[k], Wm[k * N + j]));
            od = __fadd_rn(od, __fmul_rn(a[k + 1], Wm[(k + 1) * N + j]));
        }
        for (; k < K; ++k) ev = __fadd_rn(ev, __fmul_rn(a[k], Wm[k * N + j]));
        return __fadd_rn(ev, od);
    };
    if (lane < D.Na) f[1][lane] = relu_ref(__fadd_rn(dot2(f[0], Wa, D.Ka, D.Na, lane), ba[lane]));
    __syncwarp();
    if (lane < D.Nb) f[0][lane] = relu_ref(__fadd_rn(dot2(f[1], Wb, D.Kb, D.Nb, lane), bb[lane]));
    __syncwarp();
    if constexpr (STAGE < 2) {
        if (lane < 16)
            out[(size_t)(v_begin + ul) * 16 + lane] = relu_ref(__fadd_rn(dot2(f[0], Wc, D.Kc, 16, lane), bc[lane]));
    } else {
        if (lane == 0) {   // 1 row x 1 column kernel: four accumulators, (c0+c1)+(c2+c3)
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < 16; ++k) c[k & 3] = __fadd_rn(c[k & 3], __fmul_rn(f[0][k], Wc[k]));
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(c[0], c[1]), __fadd_rn(c[2], c[3])), bc[0]);
            out[ul] = sigmoid_ref<true>(s);
        }
    }
}

}  // namespace gvc
