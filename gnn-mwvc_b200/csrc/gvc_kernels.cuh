// gvc_kernels.cuh -- sm_100a kernels of the GNN_VC forward (hot path of
// gnn::model::predict, reference src/gnn_inference.cpp:67-81).
//
// One "stage" = one graph layer (src/gnn_inference.cpp:27-42) fused with the
// dense layers, bias adds and activations that follow it (:20-25, :44-52) up to
// the next graph layer.  Nothing but the 16-float stage output per vertex (or the
// score) goes back to HBM.
//
//   gather   neighbour rows are summed in adjacency order -- the order the
//            reference adds them (:33-36) -- with 128-bit loads, 4 lanes per
//            64-byte row; the concatenated feature vector of :37-40 (including
//            its column quirk) is built on chip.
//   dense    a warp owns a tile of 32 vertices: each lane keeps a 4-vertex x
//            8-output (or 4 x 4) block of accumulators in registers, reads
//            activations and weights from shared memory with 128-bit loads,
//            applies bias and ReLU in registers and rewrites the tile in place.
//   store    16 floats per vertex (stages 0/1) or the sigmoid score.
//
// EXACT=true keeps the reference's fp32 operation order (one accumulator per
// output, k ascending, product and sum rounded separately, sequential
// neighbour sums; see oracle/gnn_oracle.c) -> bit-identical scores.
// EXACT=false runs the dense layers on the tensor cores (mma.sync m16n8k8, 3 x TF32) and uses the device expf.
// Both keep every bit the reference computes or stay within 1e-4 of it; the exact dense chain leaves out the
// k steps whose activation is zero for the whole tile (tile_linear_relu), the exact sums of the largest
// vertices are computed in parallel (gvc_px.cuh).
//
// Scheduling.  Real graphs are skewed (the R-MAT benchmark graph: 40 % isolated
// vertices, 80 % of the adjacency in vertices of degree >= 64, one vertex of
// degree 64 452) and the sums are sequential per vertex, so work is organised by
// degree (the vertices are counting-sorted by degree bin at graph upload):
//   ring    deg >= 2048   exact: all 8 warps of a CTA serve ONE vertex: each warp
//                         fetches every 8th 64-row batch, the running sum is
//                         handed from warp to warp through named barriers, so
//                         1024 rows are in flight for a single sequential chain
//                         (w=1: one warp per vertex, from 16384 neighbours on a CTA
//                         with one chain warp and 7 loader warps).
//                         fast: no order to keep -- the lists are cut into chunks of
//                         4096 entries, a CTA (w=16) or warp (w=1) per chunk, the
//                         partial sums are added in chunk order by whoever finishes last
//   mid     64 <= deg     8 vertices per warp task, 4 lanes per vertex, 16 rows in
//           < 2048        flight per vertex (w=1: one lane per vertex, as in tiles)
//   tile    deg < 64      32 vertices per warp, 4 lanes per vertex (1 for w=1)
// ring and mid tasks only produce the 32-float feature vector (side buffer,
// 128 B per vertex); "feature tiles" later run the dense chain on 32 of them, so
// no dense work is wasted on part-filled tiles.  The kernel is persistent: the
// tasks form a two-ended list (heaviest gathers ... lightest tiles); half of the
// warps of every CTA draw from the heavy end, half from the light end, so that
// latency-bound gathers and FMA-bound dense tiles overlap on every SM.
//
// Several GPUs (one process each, a vertex range per GPU): the store epilogue of
// stages 0 and 1 also writes each row into the output buffers of the ranks that
// read it (peer memory over NVLink, PeerOut), so the row exchange between two
// stages happens inside the kernel and only a barrier remains between launches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gvc_expf.h"

namespace gvc {

#ifndef GVC_WARPS_PER_CTA
#define GVC_WARPS_PER_CTA 8
#endif
constexpr int kWarpsPerCta = GVC_WARPS_PER_CTA;   // <= 15 (ring hand-over uses named barriers 1..kWarpsPerCta)
constexpr int kCtaThreads = kWarpsPerCta * 32;
#ifndef GVC_CTAS_PER_SM
#define GVC_CTAS_PER_SM 2
#endif
#ifndef GVC_DENSE_UNROLL
#define GVC_DENSE_UNROLL 2      // k-steps per loop body: small bodies keep the dense code in the instruction cache
#endif
constexpr int kCtasPerSm = GVC_CTAS_PER_SM;
constexpr int kDenseUnroll = GVC_DENSE_UNROLL;
constexpr int kTileVerts = 32;            // vertices per warp tile
constexpr int kTileStride = 36;           // floats per k-row of a tile (32 + 4 pad, keeps float4 alignment)
constexpr int kTileRows = 32;             // widest activation
constexpr int kTileFloats = kTileRows * kTileStride;
constexpr int kWarpSmemFloats = kTileFloats + 64;   // tile + vid[32] + 32 (gvc_px.cuh parks four batch records behind a staged batch)
constexpr int kSyncCounters = 4;          // sync[0..2] task claims, sync[3] ring claims; then the feature-tile counters

#ifndef GVC_MID_SETS
#define GVC_MID_SETS 2            // row sets of 4 in flight per vertex in the mid tasks (3 and 4 measured slower)
#endif
#ifndef GVC_HEAVY_WARPS
#define GVC_HEAVY_WARPS 4
#endif
#ifndef GVC_RING_MIN_DEG
#define GVC_RING_MIN_DEG 2048
#endif
#ifndef GVC_MID_MIN_DEG
#define GVC_MID_MIN_DEG 64
#endif
constexpr int kHeavyWarps = GVC_HEAVY_WARPS;    // warps per CTA that draw tasks from the heavy end of the list
constexpr uint32_t kRingMinDeg = GVC_RING_MIN_DEG;
#ifndef GVC_GIANT1_MIN_DEG
#define GVC_GIANT1_MIN_DEG 2048
#endif
#ifndef GVC_PX_MIN_DEG
#define GVC_PX_MIN_DEG 16384
#endif
constexpr uint32_t kPxMinDeg = GVC_PX_MIN_DEG;           // exact mode: >= this, the sequential sum is emulated in parallel (gvc_px.cuh)
#ifndef GVC_PX_NNZ_PER_NEIGHBOUR
#define GVC_PX_NNZ_PER_NEIGHBOUR 1000
#endif
constexpr uint64_t kPxNnzPerNeighbour = GVC_PX_NNZ_PER_NEIGHBOUR;   // ... and at least (shard's adjacency entries) / this
constexpr uint32_t kGiant1MinDeg = GVC_GIANT1_MIN_DEG;   // stage 0 (w = 1): >= one warp per vertex, below one lane    // >= : ring task (whole CTA)
constexpr uint32_t kMidMinDeg = GVC_MID_MIN_DEG;      // >= : mid task (8 vertices per warp), below: 32-vertex tiles
constexpr int kNumDegBins = 132;

// Task layout of one shard, positions refer to `order` (vertices sorted by degree bin, descending).
struct Schedule {
    uint32_t n_local;       // vertices of the shard
    uint32_t n_ring;        // order[0, n_ring)                ring tasks
    uint32_t n_mid;        // order[n_ring, n_ring + n_mid)  mid-degree vertices, 8 per task
    uint32_t n_ring_ctas;   // CTAs [0, n_ring_ctas) share the ring tasks before joining the task queue
    uint32_t n_giant1;      // order[0, n_giant1): deg >= kGiant1MinDeg, the single-warp tasks of stage 0 (w = 1)
    uint32_t n_px;          // order[0, n_px): deg >= kPxMinDeg, exact mode: sequential sums computed in parallel (gvc_px.cuh)
    uint32_t n_chunks_px;   // their 4096-entry chunks
    uint32_t n_tiles;       // 32-vertex tiles over order[n_ring + n_mid, n_local)
    uint32_t n_feat_tiles;  // 32-vertex feature tiles over order[0, n_ring + n_mid)
    uint32_t n_chunks16;    // fast mode: chunks the ring vertices are cut into (HubSplit), width 16
    uint32_t n_chunks1;     // fast mode: chunks of the stage-0 giants
};

// Fast mode owes the reference no summation order, so one huge vertex need not be one task: its
// adjacency list is cut into chunks of kChunk16 (a CTA each, width 16) or kChunk1 (a warp each,
// width 1) entries, every chunk leaves a partial sum in its slot, and whoever finishes the last
// chunk of a vertex adds the slots up in chunk order (deterministic).  chunk[k] = {position g in
// `order`, chunk index}; info[g] = {first slot, number of chunks}; done[g] counts finished chunks
// (zeroed with the other counters before every launch).
#ifndef GVC_CHUNK16
#define GVC_CHUNK16 4096
#endif
#ifndef GVC_CHUNK1
#define GVC_CHUNK1 4096
#endif
constexpr uint32_t kChunk16 = GVC_CHUNK16;
constexpr uint32_t kChunk1 = GVC_CHUNK1;
// Multi-GPU: the other ranks' copies of this stage's output buffer (peer memory over NVLink).  The
// store epilogue writes every row that another rank can read -- the rows of non-isolated vertices,
// positions [0, n_live) of `order` -- into all of them, so the row exchange between two stages is
// part of the kernel and overlaps its compute; what remains between stages is a barrier.
constexpr int kMaxPeers = 7;
struct PeerOut {
    float *p[kMaxPeers];
    const uint8_t *mask;    // indexed by GLOBAL vertex id (pointer pre-offset by -v_begin): bit q set = peer q owns a
                            // neighbour of the vertex and therefore reads its row; null = every peer gets every row
    int n;                  // 0: single GPU, or rows exchanged by a collective instead
    uint32_t n_live;        // positions of `order` below this hold vertices with neighbours
    // stage 2 only: what the caller's selection order is computed from
    // (src/GNN_VC.cpp:194-206), written beside the scores; null = not wanted
    float *keys;            // [n_local] min(out, 1 - out)
    uint8_t *side;          // [n_local] out > 0.5
};
// std::min(s, 1.0f - s) and s > 0.5f exactly as the caller's comparator evaluates them
__device__ __forceinline__ void store_selection_key(const PeerOut &po, uint32_t i, float s) {
    if (po.keys) {
        const float r = __fsub_rn(1.0f, s);
        po.keys[i] = (r < s) ? r : s;
        po.side[i] = s > 0.5f ? 1 : 0;
    }
}

struct HubSplit {
    const uint4 *chunk;
    const uint2 *info;
    float *partial;
    uint32_t *done;
};

// degree -> bin, monotone in the degree, 4 bins per octave
__host__ __device__ __forceinline__ int degree_bin(uint32_t d) {
    if (d < 4) return (int)d;                  // 0,1,2,3
#if defined(__CUDA_ARCH__)
    const int lg = 31 - __clz(d);
#else
    int lg = 31;
    while (!(d >> lg)) --lg;
#endif
    return 4 * lg + (int)((d >> (lg - 2)) & 3u) - 4;   // d=4 -> 4, contiguous from there; max 123
}

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
__host__ __device__ constexpr int stage_feat_width(int stage) { return stage == 0 ? 5 : 32; }

// Fast mode runs the [32 x K] . [K x N] layers on the tensor cores (mma.sync m16n8k8, TF32 inputs, fp32
// accumulation, every product split 3 ways -- hi.hi + lo.hi + hi.lo -- so that the result keeps fp32
// accuracy: ~1e-6 against the 1e-4 the fast mode promises).  Its parameter block holds the weights in
// FRAGMENT order, already split into TF32 hi and lo parts: for every (k block of 8, n block of 8) 128
// floats [hi b0 x 32 lanes][hi b1][lo b0][lo b1], b0 = W[8 kb + lane % 4][8 nb + lane / 4], b1 four rows
// further down (rows past K are zero); then the bias.  The 16 -> 1 layer of stage 2 stays scalar.
#ifndef GVC_FAST_MMA
#define GVC_FAST_MMA 1
#endif
__host__ __device__ constexpr int mma_layer_floats(int K, int N) { return ((K + 7) / 8) * (N / 8) * 128; }
__host__ __device__ constexpr int stage_mma_floats(int stage) {
    const StageDims D = stage_dims(stage);
    return mma_layer_floats(D.Ka, D.Na) + D.Na + mma_layer_floats(D.Kb, D.Nb) + D.Nb +
           (D.Nc >= 8 ? mma_layer_floats(D.Kc, D.Nc) : D.Kc * D.Nc) + D.Nc;
}
template <int STAGE, bool EXACT>
__host__ __device__ constexpr int stage_param_floats() {
    return (EXACT || !GVC_FAST_MMA) ? stage_dims(STAGE).floats() : stage_mma_floats(STAGE);
}

__device__ __forceinline__ float relu_ref(float v) { return v < 0.0f ? 0.0f : v; }   // std::max(x, 0.0f), :46

template <bool EXACT>
__device__ __forceinline__ float mac(float a, float w, float acc) {
    if constexpr (EXACT) return __fadd_rn(acc, __fmul_rn(a, w));
    else return fmaf(a, w, acc);
}

__device__ __forceinline__ float4 ldg_row4(const float4 *p) { return __ldg(p); }
// neighbour ids are read exactly once: stream them past the caches so they do not evict feature rows
__device__ __forceinline__ uint32_t ld_id(const uint32_t *p) {
    return __ldcs(p);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- dense: one layer + bias + ReLU over the warp's tile, in place ----------------
// T: k-major tile, T[k * kTileStride + i], i = vertex in tile.  Lane (og, vg)
// owns vertices 4vg..4vg+3 and outputs C*og..C*og+C-1.
//
// Zero rows.  The trained network is mostly ReLU-dead: measured with the GNN_VC model, 43-73 % of the
// (tile, k) pairs of the 32-wide layers hold an activation that is exactly zero for ALL 32 vertices of
// the tile (6-19 units per layer are zero for every vertex of a graph, the sums of isolated vertices are
// zero, and the vertices of a tile are alike: they are sorted by degree) -- 55 % of all multiply-adds on
// the R-MAT benchmark graph, 49 % on the grid.  Such a k adds +-0 to every accumulator, and an
// accumulator that starts at +0.0 is never -0.0 (RN: x + y = -0 only for x = y = -0), so leaving the
// step out changes no bit (weights are finite: detect_fused).  `mask` has bit k set when row k of the
// tile holds a non-zero (or NaN) value; the k loop visits the set bits in ascending order, which is the
// reference's order for the terms that remain.  The mask of a layer's OUTPUT is collected in its
// epilogue with one ballot per column group; that of a tile fresh from the gather by tile_row_mask.
__device__ __forceinline__ uint32_t tile_row_mask(const float *__restrict__ T, int lane) {
    const uint4 *row = reinterpret_cast<const uint4 *>(T + lane * kTileStride);      // lane = k
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < kTileVerts / 4; ++j) {
        const uint4 v = row[j];
        any |= v.x | v.y | v.z | v.w;
    }
    return __ballot_sync(0xffffffffu, (any & 0x7FFFFFFFu) != 0u);                    // -0.0 counts as zero
}

template <int K, int NOUT, bool EXACT>
__device__ __forceinline__ uint32_t tile_linear_relu(float *__restrict__ T, const float *__restrict__ Wsm,
                                                     const float *__restrict__ bsm, int lane, uint32_t mask) {
    constexpr int C = NOUT / 4;
    static_assert(C == 8 || C == 4, "NOUT must be 32 or 16");
    const int og = lane >> 3, vg = lane & 7;
    float acc[4][C];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[r][c] = 0.0f;

    // (Software-pipelining this loop by hand -- two operand sets alternating -- made ptxas spill in the gather
    // loops of the kernel, 236-420 bytes; the other warps of the scheduler cover the shared-memory latency.)
    uint32_t m = K >= 32 ? mask : (mask & ((1u << (K & 31)) - 1u));
#pragma unroll 1
    while (m) {
        const int k = __ffs((int)m) - 1;
        m &= m - 1u;
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
    uint32_t out_mask = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float b = bsm[C * og + c];
        float4 o;
        o.x = relu_ref(__fadd_rn(acc[0][c], b));
        o.y = relu_ref(__fadd_rn(acc[1][c], b));
        o.z = relu_ref(__fadd_rn(acc[2][c], b));
        o.w = relu_ref(__fadd_rn(acc[3][c], b));
        *reinterpret_cast<float4 *>(T + (C * og + c) * kTileStride + 4 * vg) = o;
        // output column C * og' + c is non-zero somewhere in the tile iff one of the lanes 8 og' .. 8 og' + 7 says so
        const uint32_t nz = __ballot_sync(0xffffffffu, !(o.x == 0.0f && o.y == 0.0f && o.z == 0.0f && o.w == 0.0f));
#pragma unroll
        for (int g = 0; g < 4; ++g)
            if ((nz >> (8 * g)) & 0xFFu) out_mask |= 1u << (C * g + c);
    }
    __syncwarp();
    return out_mask;
}

// Last dense layer of stages 0/1 (K=32 -> 16) + bias + ReLU, stored straight
// from registers to the stage output rows (64 B per vertex, full sectors).
template <int K, bool EXACT>
__device__ __forceinline__ void tile_linear_relu_store16(const float *__restrict__ T,
                                                         const float *__restrict__ Wsm,
                                                         const float *__restrict__ bsm, int lane,
                                                         float *__restrict__ out /* global row 0 */,
                                                         const uint32_t *__restrict__ vid /* smem: global vertex id per slot */,
                                                         int valid /* slots of the tile that hold a vertex */,
                                                         const PeerOut &peers, int live /* leading slots other ranks read */,
                                                         uint32_t mask /* non-zero rows of T, see tile_linear_relu */) {
    const int og = lane >> 3, vg = lane & 7;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    uint32_t m = K >= 32 ? mask : (mask & ((1u << (K & 31)) - 1u));
#pragma unroll 1
    while (m) {
        const int k = __ffs((int)m) - 1;
        m &= m - 1u;
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
            const size_t at = (size_t)vid[i] * 16 + 4 * og;
            *reinterpret_cast<float4 *>(out + at) = o;
            if (i < live) {
                const uint32_t m = peers.mask ? peers.mask[vid[i]] : 0xFFu;
#pragma unroll 1
                for (int q = 0; q < peers.n; ++q)
                    if (m >> q & 1u) *reinterpret_cast<float4 *>(peers.p[q] + at) = o;
            }
        }
    }
}

// ---- fast mode: the same layers on the tensor cores -------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[mb][nb][.] = sum over k of T[k][vertex] * W[k][n] for the 32 vertices of the tile: two 16-vertex
// row blocks (mb), NOUT / 8 column blocks (nb).  Fragment coordinates of lane l (r = l / 4, c = l % 4):
// A a0 (r, c) a1 (r + 8, c) a2 (r, c + 4) a3 (r + 8, c + 4);  B b0 (c, r) b1 (c + 4, r);
// C c0 (r, 2c) c1 (r, 2c + 1) c2 (r + 8, 2c) c3 (r + 8, 2c + 1).
template <int K, int NOUT>
__device__ __forceinline__ void tile_mma(const float *__restrict__ T, const float *__restrict__ Wf, int lane,
                                         float (&acc)[2][NOUT / 8][4]) {
    constexpr int KB = (K + 7) / 8, NB = NOUT / 8;
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mb][nb][i] = 0.0f;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        uint32_t hi[2][4], lo[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const float *t = T + (8 * kb + c) * kTileStride + 16 * mb + r;
            const float a[4] = {t[0], t[8], t[4 * kTileStride], t[4 * kTileStride + 8]};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                hi[mb][i] = to_tf32(a[i]);
                lo[mb][i] = to_tf32(a[i] - __uint_as_float(hi[mb][i]));
            }
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const float *f = Wf + (kb * NB + nb) * 128 + lane;
            const uint32_t bh0 = __float_as_uint(f[0]), bh1 = __float_as_uint(f[32]);
            const uint32_t bl0 = __float_as_uint(f[64]), bl1 = __float_as_uint(f[96]);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                mma_tf32(acc[mb][nb], lo[mb], bh0, bh1);     // the small terms first
                mma_tf32(acc[mb][nb], hi[mb], bl0, bl1);
                mma_tf32(acc[mb][nb], hi[mb], bh0, bh1);
            }
        }
    }
}

// one layer + bias + ReLU over the warp's tile, in place
template <int K, int NOUT>
__device__ __forceinline__ void tile_linear_relu_mma(float *__restrict__ T, const float *__restrict__ Wf,
                                                     const float *__restrict__ bsm, int lane) {
    float acc[2][NOUT / 8][4];
    tile_mma<K, NOUT>(T, Wf, lane, acc);
    __syncwarp();   // every lane is done reading the input tile
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int nb = 0; nb < NOUT / 8; ++nb) {
        const float b0 = bsm[8 * nb + 2 * c], b1 = bsm[8 * nb + 2 * c + 1];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            float *t = T + (8 * nb + 2 * c) * kTileStride + 16 * mb + r;
            t[0] = relu_ref(acc[mb][nb][0] + b0);
            t[kTileStride] = relu_ref(acc[mb][nb][1] + b1);
            t[8] = relu_ref(acc[mb][nb][2] + b0);
            t[kTileStride + 8] = relu_ref(acc[mb][nb][3] + b1);
        }
    }
    __syncwarp();
}

// last layer of stages 0/1 (K = 32 -> 16) + bias + ReLU.  The accumulator fragments hold two outputs of two rows
// per lane; stored from there, an instruction writes 32 contiguous bytes per row -- fine for local memory, but the
// rows also go to the other GPUs' buffers, and over NVLink half-row writes cost the 8-GPU fast mode 0.3 ms per
// stage (measured: stage 0 at 0.63 ms against 0.27 ms without peers).  So the 32 x 16 outputs make a round trip
// through the warp's tile buffer (free once every lane has read its A fragments) and leave as whole 64-byte rows,
// four lanes per row, exactly like the exact path's epilogue.
template <int K>
__device__ __forceinline__ void tile_linear_relu_store16_mma(float *__restrict__ T, const float *__restrict__ Wf,
                                                             const float *__restrict__ bsm, int lane, float *__restrict__ out,
                                                             const uint32_t *__restrict__ vid, int valid, const PeerOut &peers,
                                                             int live) {
    float acc[2][2][4];
    tile_mma<K, 16>(T, Wf, lane, acc);
    __syncwarp();                                    // every lane is done reading the input tile
    const int r = lane >> 2, c = lane & 3;
    constexpr int kRow = 16;                         // floats per staged row (128-bit reads conflict-free, 64-bit writes two-way)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
        const float b0 = bsm[8 * nb + 2 * c], b1 = bsm[8 * nb + 2 * c + 1];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            float *t = T + (16 * mb + r) * kRow + 8 * nb + 2 * c;
            *reinterpret_cast<float2 *>(t) = make_float2(relu_ref(acc[mb][nb][0] + b0), relu_ref(acc[mb][nb][1] + b1));
            *reinterpret_cast<float2 *>(t + 8 * kRow) = make_float2(relu_ref(acc[mb][nb][2] + b0), relu_ref(acc[mb][nb][3] + b1));
        }
    }
    __syncwarp();
    const int sv = lane >> 2, q = lane & 3;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = 8 * p + sv;
        if (i < valid) {
            const float4 o = *reinterpret_cast<const float4 *>(T + i * kRow + 4 * q);
            const uint32_t v = vid[i];
            const size_t at = (size_t)v * 16 + 4 * q;
            *reinterpret_cast<float4 *>(out + at) = o;
            if (i < live) {
                const uint32_t m = peers.mask ? peers.mask[v] : 0xFFu;
#pragma unroll 1
                for (int g = 0; g < peers.n; ++g)
                    if (m >> g & 1u) *reinterpret_cast<float4 *>(peers.p[g] + at) = o;
            }
        }
    }
}

// sigmoid::forward :49-52
template <bool EXACT>
__device__ __forceinline__ float sigmoid_ref(float v) {
    if constexpr (EXACT) return __fdiv_rn(1.0f, __fadd_rn(1.0f, gvc_expf_glibc(-v)));
    else return 1.0f / (1.0f + expf(-v));
}

// The dense chain of one stage over a filled tile, and the store.
template <int STAGE, bool EXACT>
__device__ __noinline__ void tile_dense_and_store(float *__restrict__ T, const uint32_t *__restrict__ vid,
                                                     int count, const float *__restrict__ P,
                                                     float *__restrict__ out, uint32_t v_begin, int lane,
                                                     const PeerOut &peers, int live) {
    constexpr StageDims D = stage_dims(STAGE);
    const float *Wa = P, *ba = Wa + D.Ka * D.Na;
    const float *Wb = ba + D.Na, *bb = Wb + D.Kb * D.Nb;
    const float *Wc = bb + D.Nb, *bc = Wc + D.Kc * D.Nc;
#ifdef GVC_DEBUG_SKIP_DENSE      // diagnostic build only: how long does the gather take on its own?
    if (lane < count) out[(size_t)(vid[lane] - (STAGE < 2 ? 0u : v_begin)) * (STAGE < 2 ? 16 : 1)] = T[lane];
    return;
#endif
    if constexpr (!EXACT && GVC_FAST_MMA) {
        // fragment-ordered parameter block (see stage_mma_floats)
        const float *Fa = P, *fa = Fa + mma_layer_floats(D.Ka, D.Na);
        const float *Fb = fa + D.Na, *fb = Fb + mma_layer_floats(D.Kb, D.Nb);
        const float *Fc = fb + D.Nb;
        if constexpr (D.Ka % 8 != 0) {          // stage 0: features 5..7 of the first k block must be finite zeros
#pragma unroll
            for (int k = D.Ka; k < (D.Ka + 7) / 8 * 8; ++k) T[k * kTileStride + lane] = 0.0f;
            __syncwarp();
        }
        tile_linear_relu_mma<D.Ka, D.Na>(T, Fa, fa, lane);
        tile_linear_relu_mma<D.Kb, D.Nb>(T, Fb, fb, lane);
        if constexpr (STAGE < 2) {
            tile_linear_relu_store16_mma<D.Kc>(T, Fc, Fc + mma_layer_floats(D.Kc, D.Nc), lane, out, vid, count, peers, live);   // T is scratch from here on
        } else {
            const float *Wc1 = Fc, *bc1 = Fc + D.Kc * D.Nc;
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 16; ++k) s = fmaf(T[k * kTileStride + lane], Wc1[k], s);
            if (lane < count) {
                const float sg = sigmoid_ref<false>(s + bc1[0]);
                out[vid[lane] - v_begin] = sg;
                store_selection_key(peers, vid[lane] - v_begin, sg);
            }
        }
        __syncwarp();
        return;
    }
    uint32_t rows = D.Ka >= 32 ? tile_row_mask(T, lane) : 0xFFFFFFFFu;       // the 5 features of stage 0: not worth a pass
    rows = tile_linear_relu<D.Ka, D.Na, EXACT>(T, Wa, ba, lane, rows);
    rows = tile_linear_relu<D.Kb, D.Nb, EXACT>(T, Wb, bb, lane, rows);
    if constexpr (STAGE < 2) {
        tile_linear_relu_store16<D.Kc, EXACT>(T, Wc, bc, lane, out, vid, count, peers, live, rows);
    } else {
        // 16 -> 1: one lane per vertex.  OpenBLAS' 1-column kernel: even/odd
        // accumulators, C = even + odd (oracle/gnn_oracle.c dot_two_acc).
        float ev = 0.0f, od = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            ev = mac<EXACT>(T[k * kTileStride + lane], Wc[k], ev);
            od = mac<EXACT>(T[(k + 1) * kTileStride + lane], Wc[k + 1], od);
        }
        const float s = __fadd_rn(__fadd_rn(ev, od), bc[0]);
        if (lane < count) {
            const float sg = sigmoid_ref<EXACT>(s);
            out[vid[lane] - v_begin] = sg;
            store_selection_key(peers, vid[lane] - v_begin, sg);
        }
    }
    __syncwarp();
}

// ---- gather, width 16: 4 lanes per vertex ------------------------------------------------------
// One sub-warp (lanes 4s..4s+3, lane q holds floats 4q..4q+3 of every row) sums the rows of
// one vertex in adjacency order.
// the rows of chunk k+1 are in flight while chunk k is added (in-order issue: a load is never
// consumed in the phase that issued it).  Two register sets alternate.
// Neighbour ids are fetched as aligned groups of four (one 128-bit load per sub-warp instead of
// four 32-bit ones: the id loads were half of the kernel's L1 wavefronts).  Group g holds
// col[4g .. 4g+3]; entries before the row's begin or at/after its end are masked, so a row may
// start and stop anywhere.  `col` is 16-byte aligned and readable up to the next multiple of 4
// entries (the host side guarantees both).
__device__ __forceinline__ uint4 ld_id4(const uint32_t *__restrict__ col, uint32_t g) {
    return __ldcs(reinterpret_cast<const uint4 *>(col) + g);
}
__device__ __forceinline__ void load_rows4(float4 (&r)[4], const uint4 id, const float4 *__restrict__ in4, int q,
                                           uint32_t g, uint32_t beg, uint32_t end) {
    const uint32_t e = 4 * g;
    if (e + 0 >= beg && e + 0 < end) r[0] = ldg_row4(in4 + (size_t)id.x * 4 + q);
    if (e + 1 >= beg && e + 1 < end) r[1] = ldg_row4(in4 + (size_t)id.y * 4 + q);
    if (e + 2 >= beg && e + 2 < end) r[2] = ldg_row4(in4 + (size_t)id.z * 4 + q);
    if (e + 3 >= beg && e + 3 < end) r[3] = ldg_row4(in4 + (size_t)id.w * 4 + q);
}
__device__ __forceinline__ void add_rows4(float4 &acc, const float4 (&r)[4], uint32_t g, uint32_t beg,
                                          uint32_t end) {
    const uint32_t e = 4 * g;
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (e + t >= beg && e + t < end) {
            acc.x = __fadd_rn(acc.x, r[t].x); acc.y = __fadd_rn(acc.y, r[t].y);
            acc.z = __fadd_rn(acc.z, r[t].z); acc.w = __fadd_rn(acc.w, r[t].w);
        }
}

// Groups k+1 and k+2's ids and group k+1's rows are in flight while group k is added (in-order
// issue: a load is never consumed in the phase that issued it).  Two register sets alternate.
__device__ __forceinline__ float4 gather16_vertex(const uint32_t *__restrict__ col,
                                                  const float4 *__restrict__ in4, uint32_t beg, uint32_t end,
                                                  int q) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beg >= end) return acc;
    const uint32_t gend = (end + 3) >> 2;
    uint32_t g = beg >> 2;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    uint4 idA = ld_id4(col, g);
    uint4 idB = (g + 1 < gend) ? ld_id4(col, g + 1) : zero;
    float4 rA[4], rB[4];
    load_rows4(rA, idA, in4, q, g, beg, end);
    for (;; g += 2) {
        if (g + 1 < gend) load_rows4(rB, idB, in4, q, g + 1, beg, end);
        idA = (g + 2 < gend) ? ld_id4(col, g + 2) : zero;
        add_rows4(acc, rA, g, beg, end);
        if (g + 1 >= gend) break;
        if (g + 2 < gend) load_rows4(rA, idA, in4, q, g + 2, beg, end);
        idB = (g + 3 < gend) ? ld_id4(col, g + 3) : zero;
        add_rows4(acc, rB, g + 1, beg, end);
        if (g + 2 >= gend) break;
    }
    return acc;
}

// Deeper pipeline for the mid tasks (no tile state to keep in registers there): S row sets and a
// ring of DI id groups.  At the step that adds group G: the rows of G+1 .. G+S-1 and the ids of
// G+S .. G+DI-1 are in flight; rows are requested with ids that were loaded DI-S+1 steps earlier,
// so neither the id nor the row latency sits on the per-group critical path.
template <int S, int DI>
__device__ __forceinline__ float4 gather16_vertex_deep(const uint32_t *__restrict__ col,
                                                       const float4 *__restrict__ in4, uint32_t beg, uint32_t end,
                                                       int q) {
    static_assert(DI % S == 0 && DI >= S, "id ring must be a multiple of the row sets");
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beg >= end) return acc;
    const uint32_t gend = (end + 3) >> 2;
    const uint32_t g0 = beg >> 2;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    uint4 id[DI];
    float4 r[S][4];
#pragma unroll
    for (int d = 0; d < DI; ++d) id[d] = (g0 + d < gend) ? ld_id4(col, g0 + d) : zero;
#pragma unroll
    for (int d = 0; d < S - 1; ++d)
        if (g0 + d < gend) load_rows4(r[d], id[d], in4, q, g0 + d, beg, end);
    for (uint32_t g = g0; g < gend; g += DI) {
#pragma unroll
        for (int j = 0; j < DI; ++j) {
            const uint32_t G = g + j;
            if (G < gend) {
                if (G + S - 1 < gend) load_rows4(r[(j + S - 1) % S], id[(j + S - 1) % DI], in4, q, G + S - 1, beg, end);
                add_rows4(acc, r[j % S], G, beg, end);
                id[j] = (G + DI < gend) ? ld_id4(col, G + DI) : zero;
            }
        }
    }
    return acc;
}

// self features with the :38-40 quirk: D, W/s, NW/s overwrite self features 1..3
__device__ __forceinline__ float4 self_features16(const float4 *__restrict__ in4, uint32_t ul, uint32_t deg,
                                                  const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                                  uint32_t v_begin, float scale, int q) {
    float4 self = ldg_row4(in4 + (size_t)(v_begin + ul) * 4 + q);                     // :37
    if (q == 0) {
        self.y = __uint2float_rn(deg);
        self.z = __fdiv_rn(__uint2float_rn(__ldg(Wv + ul)), scale);
        self.w = __fdiv_rn(__uint2float_rn(__ldg(NWv + ul)), scale);
    }
    return self;
}

// tiles (deg < 64): 8 vertices per pass, 4 passes; vid[i] receives the GLOBAL id of slot i.
__device__ __forceinline__ void gather16_tile(float *__restrict__ T, uint32_t *__restrict__ vid,
                                              const uint4 *__restrict__ vrec, uint32_t pos0, int count,
                                              const uint32_t *__restrict__ col,
                                              const uint32_t *__restrict__ NWv,
                                              const float4 *__restrict__ in4, uint32_t v_begin, float scale,
                                              int lane) {
    const int sv = lane >> 2, q = lane & 3;
    // vertex records {id, row begin, row end, W} of this lane's 4 vertices: one round trip for all
    // four passes instead of the chain order -> row_ptr -> W per pass
    uint4 rec[4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
        rec[p] = (p * 8 + sv < count) ? __ldg(vrec + pos0 + p * 8 + sv) : make_uint4(0u, 0u, 0u, 0u);
    // self rows and NW of all four passes are requested before the first gather starts
    float4 self[4];
    uint32_t nw[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const bool on = p * 8 + sv < count;
        self[p] = on ? ldg_row4(in4 + (size_t)(v_begin + rec[p].x) * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);   // :37
        nw[p] = (on && q == 0) ? __ldg(NWv + rec[p].x) : 0u;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = p * 8 + sv;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < count) {
            acc = gather16_vertex(col, in4, rec[p].y, rec[p].z, q);
            if (q == 0) {   // the quirk: D, W/s, NW/s overwrite self features 1..3 (:38-40)
                self[p].y = __uint2float_rn(rec[p].z - rec[p].y);
                self[p].z = __fdiv_rn(__uint2float_rn(rec[p].w), scale);
                self[p].w = __fdiv_rn(__uint2float_rn(nw[p]), scale);
                vid[i] = v_begin + rec[p].x;
            }
        }
        float *t = T + (4 * q) * kTileStride + i;
        t[0] = acc.x; t[kTileStride] = acc.y; t[2 * kTileStride] = acc.z; t[3 * kTileStride] = acc.w;
        t += 16 * kTileStride;
        t[0] = self[p].x; t[kTileStride] = self[p].y; t[2 * kTileStride] = self[p].z; t[3 * kTileStride] = self[p].w;
    }
    __syncwarp();
}

// mid task (64 <= deg < 2048): 8 vertices, one per sub-warp; the feature vectors go to the
// side buffer (128 B per vertex, two 64 B halves written by the 4 lanes of the sub-warp).
__device__ __forceinline__ void gather16_mid_task(float *__restrict__ feat, uint32_t *__restrict__ ready,
                                               const uint32_t *__restrict__ order, uint32_t pos0, int count,
                                               const uint32_t *__restrict__ row_ptr,
                                               const uint32_t *__restrict__ col,
                                               const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                               const float4 *__restrict__ in4, uint32_t v_begin, float scale,
                                               int lane) {
    const int sv = lane >> 2, q = lane & 3;
    if (sv < count) {
        const uint32_t pos = pos0 + sv;
        const uint32_t ul = __ldg(order + pos);
        const uint32_t beg = __ldg(row_ptr + ul), end = __ldg(row_ptr + ul + 1);
        const float4 acc = gather16_vertex_deep<GVC_MID_SETS, 2 * GVC_MID_SETS>(col, in4, beg, end, q);
        const float4 self = self_features16(in4, ul, end - beg, Wv, NWv, v_begin, scale, q);
        float4 *f = reinterpret_cast<float4 *>(feat + (size_t)pos * 32);
        __stcg(f + q, acc);
        __stcg(f + 4 + q, self);
        __threadfence();
    }
    __syncwarp();
    if (sv < count && q == 0) atomicAdd(ready + ((pos0 + sv) >> 5), 1u);
}

// ---- gather, width 1, tiles: one lane per vertex -------------------------------------------
__device__ __noinline__ void gather1_tile(float *__restrict__ T, uint32_t *__restrict__ vid,
                                             const uint4 *__restrict__ vrec, uint32_t pos0, int count,
                                             const uint32_t *__restrict__ col,
                                             const uint32_t *__restrict__ NWv,
                                             const float *__restrict__ x, uint32_t v_begin, float scale,
                                             int lane) {
    float agg = 0.0f, xs = 0.0f, fd = 0.0f, fw = 0.0f, fnw = 0.0f;
    if (lane < count) {
        const uint4 rec = __ldg(vrec + pos0 + lane);     // {id, row begin, row end, W}
        const uint32_t ul = rec.x;
        uint32_t e = rec.y;
        const uint32_t end = rec.z;
        const uint32_t nw = __ldg(NWv + ul);
        xs = __ldg(x + v_begin + ul);
        fd = __uint2float_rn(end - e);
        // ids as aligned groups of four (see ld_id4), two groups per trip, next trip's ids ahead
        if (e < end) {
            const uint32_t beg = e, gend = (end + 3) >> 2;
            uint32_t g = beg >> 2;
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            uint4 i0 = ld_id4(col, g), i1 = (g + 1 < gend) ? ld_id4(col, g + 1) : zero;
            for (; g < gend; g += 2) {
                const uint4 n0 = (g + 2 < gend) ? ld_id4(col, g + 2) : zero;
                const uint4 n1 = (g + 3 < gend) ? ld_id4(col, g + 3) : zero;
                const uint32_t ids[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
                float a[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t idx = 4 * g + t;
                    a[t] = (idx >= beg && idx < end) ? __ldg(x + ids[t]) : 0.0f;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t idx = 4 * g + t;
                    if (idx >= beg && idx < end) agg = __fadd_rn(agg, a[t]);
                }
                i0 = n0; i1 = n1;
            }
        }
        fw = __fdiv_rn(__uint2float_rn(rec.w), scale);
        fnw = __fdiv_rn(__uint2float_rn(nw), scale);
        vid[lane] = v_begin + ul;
    }
    T[0 * kTileStride + lane] = agg;    // [agg | x | D | W/s | NW/s], :33-40 with w = 1
    T[1 * kTileStride + lane] = xs;
    T[2 * kTileStride + lane] = fd;
    T[3 * kTileStride + lane] = fw;
    T[4 * kTileStride + lane] = fnw;
    __syncwarp();
}

// ---- 64-row batches of the ring tasks ----------------------------------------------------------
struct BatchIds { uint32_t lo, hi; };   // lane l holds neighbour ids e0+l and e0+32+l

__device__ __forceinline__ BatchIds coop_load_ids(const uint32_t *__restrict__ col, uint32_t e0, uint32_t end,
                                                  int lane) {
    BatchIds b;
    b.lo = (e0 + lane < end) ? ld_id(col + e0 + lane) : 0u;
    b.hi = (e0 + 32 + lane < end) ? ld_id(col + e0 + 32 + lane) : 0u;
    return b;
}

// 64 neighbour rows: every lane fetches 8 x 16 B (8 rows per load wave, each row coalesced)
__device__ __forceinline__ void coop_load_rows16(float4 (&r)[8], const BatchIds ids,
                                                 const float4 *__restrict__ in4, uint32_t e0, uint32_t end,
                                                 int lane) {
    const int sv = lane >> 2, q = lane & 3;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int j = 8 * w + sv;
        const uint32_t id = __shfl_sync(0xffffffffu, (w < 4) ? ids.lo : ids.hi, j & 31);
        r[w] = (e0 + j < end) ? ldg_row4(in4 + (size_t)id * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// A parked batch is stored column-major, S[c * kRingColStride + row]: the lane that chains column
// c reads four consecutive rows with one 128-bit load (the stride of 68 floats keeps the eight
// lanes of a quarter-warp on different banks).
constexpr int kRingColStride = 68;
static_assert(16 * kRingColStride <= kTileFloats, "a parked batch must fit the warp's tile buffer");

__device__ __forceinline__ void coop_stage_rows16(float *__restrict__ S, const float4 (&r)[8], int lane) {
    float *dst = S + (4 * (lane & 3)) * kRingColStride + (lane >> 2);      // column 4q, row sv
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        dst[0 * kRingColStride + 8 * w] = r[w].x;
        dst[1 * kRingColStride + 8 * w] = r[w].y;
        dst[2 * kRingColStride + 8 * w] = r[w].z;
        dst[3 * kRingColStride + 8 * w] = r[w].w;
    }
}

// The reference's sequential chain over the rows parked in S: lane c (< 16, mirrored in lanes
// 16..31) adds column c in adjacency order, one dependent FADD per neighbour.  A window of eight
// 128-bit loads (32 rows) stays ahead of the adds: with the other warps of the SM issuing global
// gathers a shared-memory load takes far longer than the 16 cycles its four adds cover
// (tools/microbench/chain_lat.cu: 5.6 ns per add with two loads ahead, 2.6 ns with eight).
// `q` holds the first 4 * kRingWindow rows on entry (loaded before the running sum arrived).
constexpr int kRingWindow = 5;            // 128-bit loads ahead (6 and more spill next to the two row buffers)
__device__ __forceinline__ float chain_add16_full(const float *__restrict__ S, float4 (&q)[kRingWindow], float acc, int lane) {
    const float4 *s4 = reinterpret_cast<const float4 *>(S + (lane & 15) * kRingColStride);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float4 a = q[k % kRingWindow];
        if (k + kRingWindow < 16) q[k % kRingWindow] = s4[k + kRingWindow];
        acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y);
        acc = __fadd_rn(acc, a.z); acc = __fadd_rn(acc, a.w);
    }
    return acc;
}

// partial batch (the end of the list)
__device__ __forceinline__ float chain_add16(const float *__restrict__ S, int cnt, float acc, int lane) {
    const float *s = S + (lane & 15) * kRingColStride;
    for (int j = 0; j < cnt; ++j) acc = __fadd_rn(acc, s[j]);
    return acc;
}

// ring task, width 16: the 8 warps of the CTA serve one vertex.  Warp w fetches batches
// w, w+8, ... (64 rows each) into its own tile buffer; the running sum travels from warp
// to warp through shared memory, the hand-over for batch b is named barrier 1 + b % 8
// (arrive by the warp that summed b-1, sync by the warp that sums b).  All 8 warps call this;
// the caller reads the 16 sums from ring_acc after a __syncthreads().
// Measured on a 262144-neighbour star (tools/chain_probe.py): 3.9 ns per neighbour against 2.07 ns
// for a bare chain of dependent FADDs (tools/microbench/chain_lat.cu: 4 cycles).  Tried and not
// faster: polling a {value, sequence number} slot instead of the barrier (4.5 ns); one warp that
// only chains, fed by 7 loader warps through full/empty barriers (4.4 ns; 4.0 ns with the parked
// batch stored column-major for 128-bit loads; 4.0 ns with both register buffers of every loader
// in flight).  The same microbenchmark shows why: a chain warp whose neighbours on the SM issue
// global gathers drops to 5.5 ns per add -- the loaders themselves are the interference.
__device__ __noinline__ void ring_gather16_exact(float *__restrict__ S, float *__restrict__ ring_acc,
                                              const uint32_t *__restrict__ col,
                                              const float4 *__restrict__ in4, uint32_t beg, uint32_t end,
                                              int warp, int lane) {
    constexpr uint32_t kStep = 64 * kWarpsPerCta;
    const uint32_t nb = (end - beg + 63) / 64;
    // Two batches of rows per warp are in flight: the rows of this warp's NEXT batch are requested
    // before it waits for the running sum of the current one, so the memory latency of a batch is
    // spread over two trips of the token around the ring (ids run two batches ahead of the rows).
    // `endk` = end while batch b + 8k exists, else 0: loads past the list are predicated off.
    auto end_of = [&](uint32_t bb) { return bb < nb ? end : 0u; };
    float4 r[8], rn[8];
    uint32_t b = warp, e0 = beg + 64 * b;
    BatchIds ids = coop_load_ids(col, e0, end_of(b), lane);
    BatchIds idn = coop_load_ids(col, e0 + kStep, end_of(b + kWarpsPerCta), lane);
    coop_load_rows16(r, ids, in4, e0, end_of(b), lane);
    ids = coop_load_ids(col, e0 + 2 * kStep, end_of(b + 2 * kWarpsPerCta), lane);
#pragma unroll 1
    for (; b < nb; b += kWarpsPerCta, e0 += kStep) {
        coop_load_rows16(rn, idn, in4, e0 + kStep, end_of(b + kWarpsPerCta), lane);
        idn = ids;
        ids = coop_load_ids(col, e0 + 3 * kStep, end_of(b + 3 * kWarpsPerCta), lane);
        coop_stage_rows16(S, r, lane);
        __syncwarp();
        const int cnt = (int)min(64u, end - e0);
        float4 va[kRingWindow];
        if (cnt == 64) {                       // first rows of the batch: on hand before the sum arrives
            const float4 *s4 = reinterpret_cast<const float4 *>(S + (lane & 15) * kRingColStride);
#pragma unroll
            for (int t = 0; t < kRingWindow; ++t) va[t] = s4[t];
        }
        float acc = 0.0f;
        if (b > 0) {
            named_bar_sync(1 + (int)(b % kWarpsPerCta), 64);
            acc = ring_acc[lane & 15];
        }
        acc = cnt == 64 ? chain_add16_full(S, va, acc, lane) : chain_add16(S, cnt, acc, lane);
        __syncwarp();
        if (lane < 16) ring_acc[lane] = acc;
        if (b + 1 < nb) named_bar_arrive(1 + (int)((b + 1) % kWarpsPerCta), 64);
#pragma unroll
        for (int w = 0; w < 8; ++w) r[w] = rn[w];
    }
}

// Fast mode does not owe the reference its summation order: every warp sums its own batches in
// registers (no staging, no hand-over), the 8 sub-warps and then the 8 warps are combined in a
// fixed order, so the result is deterministic but not the sequential chain.  ring_acc gets the
// 16 sums after the caller's __syncthreads(); `part` is kWarpsPerCta x 16 floats of shared memory.
__device__ __noinline__ void ring_gather16_fast(float *__restrict__ part, const uint32_t *__restrict__ col,
                                                const float4 *__restrict__ in4, uint32_t beg, uint32_t end,
                                                int warp, int lane) {
    constexpr uint32_t kStep = 64 * kWarpsPerCta;
    const uint32_t nb = (end - beg + 63) / 64;
    auto end_of = [&](uint32_t bb) { return bb < nb ? end : 0u; };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r[8], rn[8];
    uint32_t b = warp, e0 = beg + 64 * b;
    BatchIds ids = coop_load_ids(col, e0, end_of(b), lane);
    BatchIds idn = coop_load_ids(col, e0 + kStep, end_of(b + kWarpsPerCta), lane);
    coop_load_rows16(r, ids, in4, e0, end_of(b), lane);          // rows past `end` come back as zeros
    ids = coop_load_ids(col, e0 + 2 * kStep, end_of(b + 2 * kWarpsPerCta), lane);
#pragma unroll 1
    for (; b < nb; b += kWarpsPerCta, e0 += kStep) {
        coop_load_rows16(rn, idn, in4, e0 + kStep, end_of(b + kWarpsPerCta), lane);   // next batch in flight
        idn = ids;
        ids = coop_load_ids(col, e0 + 3 * kStep, end_of(b + 3 * kWarpsPerCta), lane);
#pragma unroll
        for (int w = 0; w < 8; ++w) { acc.x += r[w].x; acc.y += r[w].y; acc.z += r[w].z; acc.w += r[w].w; }
#pragma unroll
        for (int w = 0; w < 8; ++w) r[w] = rn[w];
    }
#pragma unroll
    for (int m = 4; m < 32; m <<= 1) {                         // the 8 sub-warps hold the same 4 columns
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, m); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, m);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, m); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, m);
    }
    if (lane < 4) *reinterpret_cast<float4 *>(part + warp * 16 + 4 * lane) = acc;
}

// ---- width 1 ---------------------------------------------------------------------------------------
// single-warp task, width 1 (stage 0 giants): blocks of 256 neighbours, lane l holds elements
// 32 t + l (coalesced ids, gathered x).  The chain needs ~4 cycles per neighbour (one dependent
// FADD), about 0.5 us per block, a gather under load takes several times that: the values of the
// next kGiant1Depth blocks and the ids of two blocks beyond those are in flight while a block is
// summed.  The block is parked in shared memory and every lane walks it with broadcast 128-bit
// loads.  Measured alternatives, both slower (3.4 ns per neighbour here): cp.async of the single
// values straight into a shared-memory ring (4.4 ns), and the prefetch code placed behind the
// fully unrolled chain's first loads so that the scheduler can interleave them (3.95 ns); a window
// of eight shared-memory loads ahead of the adds as in the width-16 chain (3.5 ns).
constexpr int kGiant1Depth = 4;

__device__ __noinline__ float coop_gather1(float *__restrict__ S /* >= 256 floats */,
                                              const uint32_t *__restrict__ col, const float *__restrict__ x,
                                              uint32_t beg, uint32_t end, int lane) {
    float acc = 0.0f;
    if (beg >= end) return acc;
    constexpr int K = kGiant1Depth;
    const uint32_t nb = (end - beg + 255) / 256;
    uint32_t ida[8], idb[8];           // ids of blocks b + K and b + K + 1
    float v[K][8];                     // values of blocks b .. b + K - 1
    auto ld_ids = [&](uint32_t (&id)[8], uint32_t blk) {
        const uint32_t e0 = beg + 256 * blk, lim = blk < nb ? end : 0u;
#pragma unroll
        for (int t = 0; t < 8; ++t) { const uint32_t e = e0 + 32 * t + lane; id[t] = (e < lim) ? ld_id(col + e) : 0u; }
    };
    auto ld_x = [&](float (&val)[8], const uint32_t (&id)[8], uint32_t blk) {
        const uint32_t e0 = beg + 256 * blk, lim = blk < nb ? end : 0u;
#pragma unroll
        for (int t = 0; t < 8; ++t) { const uint32_t e = e0 + 32 * t + lane; val[t] = (e < lim) ? __ldg(x + id[t]) : 0.0f; }
    };
#pragma unroll
    for (int k = 0; k < K; ++k) {      // prologue: ids then values of the first K blocks
        ld_ids(ida, k);
        ld_x(v[k], ida, k);
    }
    ld_ids(ida, K);
    ld_ids(idb, K + 1);
#pragma unroll 1
    for (uint32_t blk = 0; blk < nb; ++blk) {
        const uint32_t e0 = beg + 256 * blk;
#pragma unroll
        for (int t = 0; t < 8; ++t) S[32 * t + lane] = v[0][t];
#pragma unroll
        for (int k = 0; k + 1 < K; ++k) {
#pragma unroll
            for (int t = 0; t < 8; ++t) v[k][t] = v[k + 1][t];
        }
        if (blk + K + 4 <= nb) {
            // blocks blk + K and blk + K + 2 are full: no bounds checks, one base address per block
            // (every instruction spent on fetching is time the chain of this single warp stands still)
#pragma unroll
            for (int t = 0; t < 8; ++t) v[K - 1][t] = __ldg(x + ida[t]);
            const uint32_t *cp = col + e0 + 256 * (K + 2) + lane;
#pragma unroll
            for (int t = 0; t < 8; ++t) { ida[t] = idb[t]; idb[t] = ld_id(cp + 32 * t); }
        } else {
            ld_x(v[K - 1], ida, blk + K);         // values K blocks ahead
#pragma unroll
            for (int t = 0; t < 8; ++t) ida[t] = idb[t];
            ld_ids(idb, blk + K + 2);
        }
        __syncwarp();
        const int cnt = (int)min(256u, end - e0);
        const float4 *s4 = reinterpret_cast<const float4 *>(S);
        int j = 0;
        if (cnt == 256) {
            // the next 16 values are loaded while the current 16 are added: the chain runs at the
            // FADD latency (4 cycles per neighbour), not at shared-memory latency
            float4 a = s4[0], b = s4[1], c = s4[2], d = s4[3];
#pragma unroll 4
            for (; j < 256; j += 16) {
                const int n4 = (j + 16 < 256) ? (j + 16) / 4 : 0;
                const float4 na = s4[n4], nb4 = s4[n4 + 1], nc = s4[n4 + 2], nd = s4[n4 + 3];
                acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y); acc = __fadd_rn(acc, a.z); acc = __fadd_rn(acc, a.w);
                acc = __fadd_rn(acc, b.x); acc = __fadd_rn(acc, b.y); acc = __fadd_rn(acc, b.z); acc = __fadd_rn(acc, b.w);
                acc = __fadd_rn(acc, c.x); acc = __fadd_rn(acc, c.y); acc = __fadd_rn(acc, c.z); acc = __fadd_rn(acc, c.w);
                acc = __fadd_rn(acc, d.x); acc = __fadd_rn(acc, d.y); acc = __fadd_rn(acc, d.z); acc = __fadd_rn(acc, d.w);
                a = na; b = nb4; c = nc; d = nd;
            }
        }
        for (; j + 16 <= cnt; j += 16) {
            const float4 a = s4[j / 4], b = s4[j / 4 + 1], c = s4[j / 4 + 2], d = s4[j / 4 + 3];
            acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y); acc = __fadd_rn(acc, a.z); acc = __fadd_rn(acc, a.w);
            acc = __fadd_rn(acc, b.x); acc = __fadd_rn(acc, b.y); acc = __fadd_rn(acc, b.z); acc = __fadd_rn(acc, b.w);
            acc = __fadd_rn(acc, c.x); acc = __fadd_rn(acc, c.y); acc = __fadd_rn(acc, c.z); acc = __fadd_rn(acc, c.w);
            acc = __fadd_rn(acc, d.x); acc = __fadd_rn(acc, d.y); acc = __fadd_rn(acc, d.z); acc = __fadd_rn(acc, d.w);
        }
        for (; j < cnt; ++j) acc = __fadd_rn(acc, S[j]);
        __syncwarp();
    }
    return acc;
}

// fast mode: every lane sums its own elements, one shuffle reduction at the end; the values of
// the next block and the ids of the one after are in flight while a block is added up
__device__ __noinline__ float coop_gather1_fast(const uint32_t *__restrict__ col, const float *__restrict__ x,
                                                uint32_t beg, uint32_t end, int lane) {
    float acc = 0.0f;
    if (beg >= end) return acc;
    const uint32_t nb = (end - beg + 255) / 256;
    uint32_t id[8];
    float v[8], vn[8];
    auto ld_ids = [&](uint32_t blk) {
        const uint32_t e0 = beg + 256 * blk, lim = blk < nb ? end : 0u;
#pragma unroll
        for (int t = 0; t < 8; ++t) { const uint32_t e = e0 + 32 * t + lane; id[t] = (e < lim) ? ld_id(col + e) : 0u; }
    };
    auto ld_x = [&](float (&val)[8], uint32_t blk) {
        const uint32_t e0 = beg + 256 * blk, lim = blk < nb ? end : 0u;
#pragma unroll
        for (int t = 0; t < 8; ++t) { const uint32_t e = e0 + 32 * t + lane; val[t] = (e < lim) ? __ldg(x + id[t]) : 0.0f; }
    };
    ld_ids(0);
    ld_x(v, 0);
    ld_ids(1);
#pragma unroll 1
    for (uint32_t blk = 0; blk < nb; ++blk) {
        ld_x(vn, blk + 1);
        ld_ids(blk + 2);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc += v[t];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = vn[t];
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    return acc;
}

}  // namespace gvc
#include "gvc_px.cuh"
namespace gvc {

// ---- feature vectors of ring/mid vertices: feat[pos * 32 + k] -------------------------------------
// width 16: lanes 0-15 hold agg[c], lanes 16-31 supply self[c] with the :38-40 quirk
__device__ __forceinline__ void put_features16(float *__restrict__ feat, uint32_t pos, float acc,
                                               const float *__restrict__ in, uint32_t ul, uint32_t deg,
                                               const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                               uint32_t v_begin, float scale, int lane) {
    float f = acc;
    if (lane >= 16) {
        const int c = lane - 16;
        f = __ldg(in + (size_t)(v_begin + ul) * 16 + c);
        if (c == 1) f = __uint2float_rn(deg);
        if (c == 2) f = __fdiv_rn(__uint2float_rn(__ldg(Wv + ul)), scale);
        if (c == 3) f = __fdiv_rn(__uint2float_rn(__ldg(NWv + ul)), scale);
    }
    __stcg(feat + (size_t)pos * 32 + lane, f);
}

__device__ __forceinline__ void put_features1(float *__restrict__ feat, uint32_t pos, float acc,
                                              const float *__restrict__ x, uint32_t ul, uint32_t deg,
                                              const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
                                              uint32_t v_begin, float scale, int lane) {
    float f = acc;                                                          // [agg | x | D | W/s | NW/s]
    if (lane == 1) f = __ldg(x + v_begin + ul);
    if (lane == 2) f = __uint2float_rn(deg);
    if (lane == 3) f = __fdiv_rn(__uint2float_rn(__ldg(Wv + ul)), scale);
    if (lane == 4) f = __fdiv_rn(__uint2float_rn(__ldg(NWv + ul)), scale);
    if (lane < 5) __stcg(feat + (size_t)pos * 32 + lane, f);
}

// one more feature vector of feature tile pos/32 is complete (release)
__device__ __forceinline__ void publish_feature(uint32_t *__restrict__ ready, uint32_t pos, int lane) {
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(ready + (pos >> 5), 1u);
}

// ---- the fused stage kernel (persistent) --------------------------------------------------------------
// STAGE 0: in = x [n_global],      out = h rows [n_global x 16]
// STAGE 1: in = h [n_global x 16], out = h rows [n_global x 16]
// STAGE 2: in = h [n_global x 16], out = scores [n_local]
// sync[0..2] = task counters, sync[3] = ring claims, sync[4 + t] = finished feature vectors of feature
// tile t; zeroed before launch.
template <int STAGE, bool EXACT>
__global__ void __launch_bounds__(kCtaThreads, kCtasPerSm)
stage_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
             const uint32_t *__restrict__ Wv, const uint32_t *__restrict__ NWv,
             const uint32_t *__restrict__ order, const uint4 *__restrict__ vrec, const Schedule sc,
             const HubSplit hub, const PxArgs px, const PeerOut peers_arg, float *__restrict__ feat,
             uint32_t *__restrict__ sync, const float *__restrict__ in, float *__restrict__ out,
             const float *__restrict__ params, uint32_t v_begin, float scale) {
    constexpr StageDims D = stage_dims(STAGE);
    extern __shared__ __align__(16) float smem[];
    float *P = smem;                                             // packed parameters
    constexpr int kParams = stage_param_floats<STAGE, EXACT>();
    constexpr int kParamFloats = (kParams + 3) / 4 * 4;
    float *ring_acc = smem + kParamFloats;                       // 16 sums, the CTA's ring claim [16], a chunk record [20..23],
    PeerOut &peers = *reinterpret_cast<PeerOut *>(ring_acc + 24);   // the peer table [24..47] (shared memory: passed by reference)
    static_assert(sizeof(PeerOut) <= 96, "PeerOut is laid out in 24 floats of shared memory");
    float *warp_mem = ring_acc + 48;

    for (int i = threadIdx.x; i < kParams; i += kCtaThreads) P[i] = __ldg(params + i);
    if (threadIdx.x == 0) peers = peers_arg;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *T = warp_mem + warp * kWarpSmemFloats;
    uint32_t *vid = reinterpret_cast<uint32_t *>(T + kTileFloats);
    uint32_t *ready = sync + kSyncCounters;

    // ---- ring tasks (width 16 only): the whole CTA, largest vertices first --------------------
    // (stage 0 runs its giants below 16384 neighbours as single-warp tasks of the queue instead)
    // ---- exact mode: the vertices of degree >= kPxMinDeg ------------------------------------------
    // Their sequential sums are emulated in parallel (gvc_px.cuh): the chunks of their lists are gathered
    // and quantised here by whole CTAs (phases A and B); the in-order walk over the batches (phase C)
    // is a single-warp task at the very front of the queue below, so that it starts at once on the
    // CTAs that do not take part here and overlaps everything else the stage has to do.
    if constexpr (EXACT) {
        if (blockIdx.x < sc.n_ring_ctas && sc.n_px)
            px_phases_ab<STAGE == 0 ? 1 : 16>(px, col, in, reinterpret_cast<uint32_t *>(ring_acc) + 16, warp_mem, warp, lane);
    }
    if constexpr (STAGE != 0) {
        // claimed one at a time, largest first: a CTA that drew a huge vertex takes fewer of them
        uint32_t *claim = reinterpret_cast<uint32_t *>(ring_acc) + 16;
        if constexpr (EXACT) {
#pragma unroll 1
            while (blockIdx.x < sc.n_ring_ctas && sc.n_ring > sc.n_px) {
                if (threadIdx.x == 0) *claim = sc.n_px + atomicAdd(sync + 3, 1u);     // [0, n_px) go the parallel way
                __syncthreads();
                const uint32_t g = *claim;
                if (g >= sc.n_ring) break;
                const uint32_t ul = __ldg(order + g);
                const uint32_t beg = __ldg(row_ptr + ul), end = __ldg(row_ptr + ul + 1);
                ring_gather16_exact(T, ring_acc, col, reinterpret_cast<const float4 *>(in), beg, end, warp, lane);
                __syncthreads();
                if (warp == 0) {
                    put_features16(feat, g, ring_acc[lane & 15], in, ul, end - beg, Wv, NWv, v_begin, scale, lane);
                    publish_feature(ready, g, lane);
                }
                __syncthreads();
            }
        } else if (blockIdx.x < sc.n_ring_ctas && sc.n_chunks16) {
            // fast mode claims chunks (HubSplit).  Thread 0 keeps one claim and one chunk record
            // ahead: the atomic for chunk i + 2 and the record of chunk i + 1 are in flight while
            // the CTA gathers chunk i.
            uint4 *rec_s = reinterpret_cast<uint4 *>(ring_acc + 20);
            const uint4 none = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
            uint32_t k_next = 0;
            uint4 rec_next = none;
            if (threadIdx.x == 0) {
                const uint32_t k0 = atomicAdd(sync + 3, 1u);
                *rec_s = k0 < sc.n_chunks16 ? __ldg(hub.chunk + k0) : none;
                k_next = atomicAdd(sync + 3, 1u);
            }
#pragma unroll 1
            for (;;) {
                __syncthreads();
                const uint4 ck = *rec_s;                               // {g, first entry, end entry, chunk index}
                if (ck.x == 0xFFFFFFFFu) break;
                if (threadIdx.x == 0) {
                    rec_next = k_next < sc.n_chunks16 ? __ldg(hub.chunk + k_next) : none;
                    k_next = atomicAdd(sync + 3, 1u);
                }
                const uint32_t g = ck.x;
                const uint2 hi = __ldg(hub.info + g);                  // {first slot, chunks}
                float *part = warp_mem;                    // warp 0's tile buffer: free during the ring phase
                ring_gather16_fast(part, col, reinterpret_cast<const float4 *>(in), ck.y, ck.z, warp, lane);
                __syncthreads();
                if (threadIdx.x == 0) *rec_s = rec_next;               // every thread has read the old record
                bool last = true;
                if (threadIdx.x < 16) {
                    float sum = 0.0f;
                    for (int w = 0; w < kWarpsPerCta; ++w) sum += part[w * 16 + threadIdx.x];
                    ring_acc[threadIdx.x] = sum;
                    if (hi.y > 1) {
                        __stcg(hub.partial + (size_t)(hi.x + ck.w) * 16 + threadIdx.x, sum);
                        __threadfence();
                    }
                }
                if (hi.y > 1) {                                        // uniform over the CTA
                    __syncthreads();
                    if (threadIdx.x == 0) *claim = atomicAdd(hub.done + g, 1u);
                    __syncthreads();
                    last = *claim == hi.y - 1;
                    if (last && threadIdx.x < 16) {
                        __threadfence();
                        float sum = 0.0f;
                        for (uint32_t c = 0; c < hi.y; ++c) sum += __ldcg(hub.partial + (size_t)(hi.x + c) * 16 + threadIdx.x);
                        ring_acc[threadIdx.x] = sum;
                    }
                }
                __syncthreads();
                if (last && warp == 0) {
                    const uint32_t ul = __ldg(order + g);
                    const uint32_t deg = __ldg(row_ptr + ul + 1) - __ldg(row_ptr + ul);
                    put_features16(feat, g, ring_acc[lane & 15], in, ul, deg, Wv, NWv, v_begin, scale, lane);
                    publish_feature(ready, g, lane);
                }
            }
            __syncthreads();
        }
    }

    // ---- dynamic tasks, one warp each ----------------------------------------------------------------
    // stage 0: front = the giants (exact: those the ring phase did not take, one warp each; fast:
    // their chunks); everything else is a 32-vertex tile.
    // stages 1/2: front = mid tasks (8 vertices each); tiles hold the vertices of degree < 64.
    const uint32_t n_pre = STAGE == 0 ? sc.n_giant1 : sc.n_ring + sc.n_mid;   // positions that go through feature tiles
    const uint32_t n_walk = EXACT ? sc.n_px : 0u;           // front of the front: the phase-C walks of the largest vertices
    const uint32_t n_front = STAGE == 0 ? (EXACT ? sc.n_giant1 : sc.n_chunks1) : n_walk + (sc.n_mid + 7) / 8;
    const uint32_t n_tiles = (sc.n_local - n_pre + kTileVerts - 1) / kTileVerts;
    const uint32_t n_heavy = n_front + n_tiles;             // dealt alternately from both ends
    const uint32_t n_tasks = n_heavy + (n_pre + kTileVerts - 1) / kTileVerts;
    // Two-ended task list: [front tasks, heaviest first | tiles, heavier to lighter].  Warps with a
    // "heavy" role draw from the front, the others from the back, so that every SM always has warps
    // stalled on gathers AND warps running dense tiles (with a single alternating counter all warps
    // end up holding long gather tasks at the same time and the FMA pipes idle).  sync[0] counts all
    // claims, sync[1] the front claims, sync[2] the back claims: front ids 0..F-1 and back ids
    // n-1..n-B can never meet because F + B <= n.
    const bool heavy_role = warp < kHeavyWarps;
#pragma unroll 1
    for (;;) {
        uint32_t k = 0, g = 0;
        if (lane == 0) {
            k = atomicAdd(sync, 1u);
            if (k < n_heavy) g = heavy_role ? atomicAdd(sync + 1, 1u) : n_heavy - 1 - atomicAdd(sync + 2, 1u);
        }
        k = __shfl_sync(0xffffffffu, k, 0);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (k >= n_tasks) break;
        if (k < n_heavy) {
            if (g < n_front) {
                if constexpr (STAGE == 0 && EXACT) {
                    const uint32_t pos = g;
                    const uint32_t ul = __ldg(order + pos);
                    const uint32_t beg = __ldg(row_ptr + ul), end = __ldg(row_ptr + ul + 1);
                    const float acc = g < n_walk ? px_walk1(px, g, end - beg, lane) : coop_gather1(T, col, in, beg, end, lane);
                    put_features1(feat, pos, acc, in, ul, end - beg, Wv, NWv, v_begin, scale, lane);
                    publish_feature(ready, pos, lane);
                } else if constexpr (STAGE == 0) {
                    const uint4 ck = __ldg(hub.chunk + g);             // front task = one chunk of a giant
                    const uint32_t pos = ck.x;
                    const uint2 hi = __ldg(hub.info + pos);
                    float acc = coop_gather1_fast(col, in, ck.y, ck.z, lane);
                    bool last = true;
                    if (hi.y > 1) {
                        uint32_t prev = 0;
                        if (lane == 0) {
                            __stcg(hub.partial + hi.x + ck.w, acc);
                            __threadfence();
                            prev = atomicAdd(hub.done + pos, 1u);
                        }
                        last = __shfl_sync(0xffffffffu, prev, 0) == hi.y - 1;
                        if (last) {
                            __threadfence();
                            acc = 0.0f;                                // slots in chunk order: lane-strided, then a fixed tree
                            for (uint32_t c = lane; c < hi.y; c += 32) acc += __ldcg(hub.partial + hi.x + c);
#pragma unroll
                            for (int m = 1; m < 32; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                        }
                    }
                    if (last) {
                        const uint32_t ul = __ldg(order + pos);
                        const uint32_t deg = __ldg(row_ptr + ul + 1) - __ldg(row_ptr + ul);
                        put_features1(feat, pos, acc, in, ul, deg, Wv, NWv, v_begin, scale, lane);
                        publish_feature(ready, pos, lane);
                    }
                } else if (EXACT && g < n_walk) {
                    const uint32_t ul = __ldg(order + g);
                    const uint32_t deg = __ldg(row_ptr + ul + 1) - __ldg(row_ptr + ul);
                    const float acc = px_walk16(px, g, deg, T, lane);
                    put_features16(feat, g, acc, in, ul, deg, Wv, NWv, v_begin, scale, lane);
                    publish_feature(ready, g, lane);
                } else {
                    const uint32_t pos0 = sc.n_ring + 8 * (g - n_walk);
                    gather16_mid_task(feat, ready, order, pos0, (int)min(8u, n_pre - pos0), row_ptr, col, Wv, NWv,
                                      reinterpret_cast<const float4 *>(in), v_begin, scale, lane);
                }
            } else {
                // 32-vertex tile: gather + dense + store
                const uint32_t pos0 = n_pre + (g - n_front) * kTileVerts;
                const int count = (int)min((uint32_t)kTileVerts, sc.n_local - pos0);
                if constexpr (STAGE == 0)
                    gather1_tile(T, vid, vrec, pos0, count, col, NWv, in, v_begin, scale, lane);
                else
                    gather16_tile(T, vid, vrec, pos0, count, col, NWv,
                                  reinterpret_cast<const float4 *>(in), v_begin, scale, lane);
                const int live = peers.n_live > pos0 ? (int)min((uint32_t)count, peers.n_live - pos0) : 0;
                tile_dense_and_store<STAGE, EXACT>(T, vid, count, P, out, v_begin, lane, peers, live);
            }
        } else {
            // feature tile: 32 precomputed feature vectors -> dense + store
            const uint32_t ft = k - n_heavy;
            const uint32_t pos0 = ft * kTileVerts;
            const int count = (int)min((uint32_t)kTileVerts, n_pre - pos0);
            if (lane == 0) {
                while (*reinterpret_cast<volatile uint32_t *>(ready + ft) < (uint32_t)count) __nanosleep(200);
                __threadfence();
            }
            __syncwarp();
            if (lane < count) vid[lane] = v_begin + __ldg(order + pos0 + lane);
            constexpr int FW = stage_feat_width(STAGE);
            if (lane < FW) {
#pragma unroll 4
                for (int i = 0; i < kTileVerts; ++i)
                    T[lane * kTileStride + i] = (i < count) ? __ldcg(feat + (size_t)(pos0 + i) * 32 + lane) : 0.0f;
            }
            __syncwarp();
            tile_dense_and_store<STAGE, EXACT>(T, vid, count, P, out, v_begin, lane, peers, count);   // degree >= 64: all live
        }
    }
}

template <int STAGE, bool EXACT>
constexpr size_t stage_smem_bytes() {
    return ((stage_param_floats<STAGE, EXACT>() + 3) / 4 * 4 + 48 + kWarpsPerCta * kWarpSmemFloats) * sizeof(float);
}

// ---- schedule construction (graph upload time) -------------------------------------------
// counting sort of the local vertices by degree bin: histogram, (host) scan, scatter
__global__ void degree_hist_kernel(const uint32_t *__restrict__ row_ptr, uint32_t n_local,
                                   uint32_t *__restrict__ hist /* kNumDegBins, zeroed */) {
    __shared__ uint32_t sh[kNumDegBins];
    for (int i = threadIdx.x; i < kNumDegBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_local; u += gridDim.x * blockDim.x)
        atomicAdd(&sh[degree_bin(row_ptr[u + 1] - row_ptr[u])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kNumDegBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void degree_scatter_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ Wv,
                                      uint32_t n_local,
                                      uint32_t *__restrict__ cursor /* kNumDegBins: start of each bin */,
                                      uint32_t *__restrict__ order, uint4 *__restrict__ vrec) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_local; u += gridDim.x * blockDim.x) {
        const uint32_t beg = row_ptr[u], end = row_ptr[u + 1];
        const uint32_t pos = atomicAdd(&cursor[degree_bin(end - beg)], 1u);
        order[pos] = u;
        vrec[pos] = make_uint4(u, beg, end, Wv[u]);      // everything a tile needs about the vertex, 16 B
    }
}

// Which peers read the row of local vertex u?  Those that own one of its neighbours (the adjacency
// is symmetric).  bounds[0..n_parts] are the vertex ranges of the parts, peer_of_part[k] the index
// of part k's owner in the peer tables (-1: this rank).  One thread per vertex; a hub is done as
// soon as every peer has been seen.
struct PartMap {
    uint32_t bounds[kMaxPeers + 2];
    int peer_of_part[kMaxPeers + 1];
    int n_parts;
};
__global__ void peer_mask_kernel(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                 uint32_t n_local, const PartMap pm, uint32_t all_peers, uint8_t *__restrict__ mask) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_local; u += gridDim.x * blockDim.x) {
        uint32_t m = 0;
        for (uint32_t e = row_ptr[u], end = row_ptr[u + 1]; e < end && m != all_peers; ++e) {
            const uint32_t v = col[e];
            int k = 0;
            while (k + 1 < pm.n_parts && v >= pm.bounds[k + 1]) ++k;
            const int q = pm.peer_of_part[k];
            if (q >= 0) m |= 1u << q;
        }
        mask[u] = (uint8_t)m;
    }
}

// fast-mode hub chunks of order[0, n_class): see HubSplit.  counter[0] ends up as the number of chunks.
__global__ void hub_chunks_kernel(const uint4 *__restrict__ vrec, uint32_t n_class, uint32_t chunk_len,
                                  uint32_t *__restrict__ counter, uint2 *__restrict__ info,
                                  uint4 *__restrict__ chunk, uint32_t capacity /* entries of `chunk` */) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_class; g += gridDim.x * blockDim.x) {
        const uint4 r = vrec[g];
        const uint32_t nch = max(1u, (r.z - r.y + chunk_len - 1) / chunk_len);
        const uint32_t base = atomicAdd(counter, nch);
        info[g] = make_uint2(base, nch);
        // (capacity = nnz / chunk_len + n_class covers every consistent CSR; the check keeps offsets
        // a caller adopted without validation from writing past the list)
        for (uint32_t c = 0; c < nch && base + c < capacity; ++c)
            chunk[base + c] = make_uint4(g, r.y + c * chunk_len, min(r.z, r.y + (c + 1) * chunk_len), c);
    }
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
                  uint32_t ul, uint32_t v_begin, float scale, const PeerOut peers) {
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
        if (lane < 16) {
            const float o = relu_ref(__fadd_rn(dot2(f[0], Wc, D.Kc, 16, lane), bc[lane]));
            out[(size_t)(v_begin + ul) * 16 + lane] = o;
            const uint32_t m = peers.mask ? peers.mask[v_begin + ul] : 0xFFu;
            if (end > beg)
                for (int q = 0; q < peers.n; ++q)
                    if (m >> q & 1u) peers.p[q][(size_t)(v_begin + ul) * 16 + lane] = o;
        }
    } else {
        if (lane == 0) {   // 1 row x 1 column kernel: four accumulators, (c0+c1)+(c2+c3)
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < 16; ++k) c[k & 3] = __fadd_rn(c[k & 3], __fmul_rn(f[0][k], Wc[k]));
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(c[0], c[1]), __fadd_rn(c[2], c[3])), bc[0]);
            const float sg = sigmoid_ref<true>(s);
            out[ul] = sg;
            store_selection_key(peers, ul, sg);
        }
    }
}

}  // namespace gvc
