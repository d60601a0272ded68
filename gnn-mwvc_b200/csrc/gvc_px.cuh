// gvc_px.cuh -- the reference's SEQUENTIAL fp32 neighbour sum of a huge vertex, computed in parallel,
// bit for bit (exact mode; included by gvc_kernels.cuh).
//
// graph_layer::forward adds the rows of a vertex's neighbours one after the other into one fp32
// accumulator per column (reference src/gnn_inference.cpp:33-36).  For a hub of 258 306 neighbours
// (R-MAT scale 23) that is a chain of 258 306 dependent additions: 0.9 ms at the 4-cycle latency of an
// FADD -- longer than everything else a GPU has to do in that stage -- and no rearrangement of a
// floating-point sum is allowed if the scores are to stay bit-identical.
//
// What makes it parallel after all: the addends are NOT arbitrary.  x = W/s > 0, and h1/h2 are ReLU
// outputs >= 0, so the running sum only grows, and while it stays inside one binade [2^e, 2^(e+1))
// every intermediate sum is a multiple of u = 2^(e-23).  There "acc = RN(acc + v)" is integer
// arithmetic on multiples of u:
//     acc' = acc + d(v),   d(v) = v rounded to the nearest multiple of u   (= RN(2^e + v) - 2^e, exact)
// unless v lies exactly half way between two multiples (then the tie goes to the even neighbour, which
// depends on acc itself).  Sums of multiples of u below 2^(e+1) are exact in fp32 IN ANY ORDER.  So,
// for a batch of consecutive neighbours, with e the binade of the running sum at the batch's entry:
//     if no addend is a tie, all are in [0, 2^e), and acc + sum d(v_i) < 2^(e+1)
//     then the sequential result after the batch is exactly acc + D,  D = sum d(v_i)  (any order).
// e is not known when the batches are processed in parallel -- it is PREDICTED from approximate prefix
// sums, and the prediction is VERIFIED when the batches are finally put together in order: a batch whose
// prediction fails, or that holds a tie, a negative/NaN/huge addend or crosses into the next binade, is
// simply added up the reference's way, element by element (its rows are kept).  Nothing is assumed
// that is not checked, so the result is the reference's for ANY input; only the speed depends on the
// inputs being the non-negative ones this network produces (about 1 batch in 15-50 goes the slow way).
//
// Three phases per hub, all inside the stage kernel (cooperative launch: every CTA is resident):
//   A  4096-entry chunks of the adjacency list, one CTA each, any CTA: gather the rows into a scratch
//      copy (coalesced from then on), leave one approximate sum per batch and column and one per chunk;
//      whoever finishes the last chunk of a hub turns the chunk sums into the entry value of every chunk
//   B  again per chunk: entry value P of every batch (prefix inside the chunk), then D and a "clean"
//      flag per batch under the binade predicted from P; one 128-byte record {P, D} per batch
//   C  one warp per hub walks the batches in order: one FADD and three checks per clean batch, the
//      element-wise chain for the others (rows staged through shared memory)
// Everything that crosses CTAs goes through L2 (st.cg / ld.cg) behind fences and counters.
// Validated against the element-wise chain on the CPU (tests/test_px_model.py restates it in numpy) and
// on the GPU against the oracle (tests/test_gpu_parity.py: hubs up to 262 144 neighbours, adversarial
// inputs).
#pragma once

namespace gvc {

constexpr uint32_t kPxChunk = 4096;                       // entries per chunk (phases A and B: one CTA per chunk)
template <int W> struct PxGeom;                           // W = floats per row: 16 (h rows) or 1 (x)
template <> struct PxGeom<16> { static constexpr int kBatch = 64; };
template <> struct PxGeom<1> { static constexpr int kBatch = 256; };
template <int W> __host__ __device__ constexpr int px_batches_per_chunk() { return (int)kPxChunk / PxGeom<W>::kBatch; }

struct PxArgs {
    const uint4 *chunk;      // [n_chunks] {position g in `order`, first entry, end entry, chunk index within the hub}
    const uint2 *info;       // [n_hubs]   {first chunk of the hub, number of chunks}
    float *scratch;          // [n_chunks][4096][W] rows in adjacency order (zero rows past the end of a list)
    float *S;                // [n_chunks][batches per chunk][W] approximate sum of every batch (phase A)
    float *T;                // [n_chunks][W] approximate sum of every chunk (phase A); the scan turns it into the chunk's entry value
    float *rec;              // [n_chunks][batches per chunk][2][W] {P = predicted entry value (-1: all-zero column), D = exact increment}
    uint32_t *flag;          // [n_chunks][batches per chunk]: 1 = add this batch element by element
    uint32_t *ctr;           // [0..1] claim counters of phases A, B; [2..7] statistics; from [8] per hub {A done, scan done, B done}
    uint32_t n_chunks, n_hubs;
};

// Binade of the running sum predicted for a batch that is entered at about P and adds about S:
// false if the interval [P, P + S] widened by 1/4096 either way touches a power of two, or P is not a
// positive normal number well inside the exponent range.  mbits = bit pattern of 2^e.  (The margin only
// trades rare failed verifications -- the prefix sums are good to about 1e-5 -- against batches that
// are sent the slow way needlessly; exactness never depends on it.)
constexpr float kPxLo = 1.0f - 1.0f / 4096.0f, kPxHi = 1.0f + 1.0f / 4096.0f;
__device__ __forceinline__ bool px_predict(float P, float S, uint32_t &mbits) {
    const float lo = __fmul_rn(P, kPxLo), hi = __fmul_rn(__fadd_rn(P, S), kPxHi);
    mbits = __float_as_uint(lo) & 0x7F800000u;
    const uint32_t hb = __float_as_uint(hi) & 0x7F800000u;
    return lo > 0.0f && hi >= lo && mbits == hb && mbits >= (26u << 23) && mbits <= (252u << 23);
}
__device__ __forceinline__ uint32_t px_entry_binade(float P) { return __float_as_uint(__fmul_rn(P, kPxLo)) & 0x7F800000u; }

// d(v) for the binade of M = 2^e, and whether v may take the fast way (see the header comment)
__device__ __forceinline__ float px_quantum(float v, float M, float half_u, bool &bad) {
    const float s = __fadd_rn(M, v);
    const float d = __fsub_rn(s, M);
    const float t = __fsub_rn(v, d);              // exact: the rounding error of M + v
    bad |= !(v >= 0.0f && v < M) || fabsf(t) == half_u;
    return d;
}

__device__ __forceinline__ void px_wait(const uint32_t *flag, uint32_t want) {
    while (*reinterpret_cast<const volatile uint32_t *>(flag) < want) __nanosleep(100);
    __threadfence();
}

// ---- phase A, width 16: the CTA gathers one chunk -----------------------------------------------------
// (same 64-row batches and register layout as the ring: lane (sv, q) holds floats 4q..4q+3 of rows
// 8w + sv; the next batch's rows and the ids of the one after are in flight).  sh: 8 x 16 floats.
__device__ __noinline__ void px_gather16(const PxArgs &px, uint32_t k, const uint32_t *__restrict__ col,
                                         const float4 *__restrict__ in4, float *__restrict__ sh, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t beg = ck.y, end = ck.z;
    const uint32_t nb = (end - beg + 63) / 64;                       // batches that hold entries (<= 64)
    float4 *rows = reinterpret_cast<float4 *>(px.scratch) + (size_t)k * kPxChunk * 4;
    float4 *S4 = reinterpret_cast<float4 *>(px.S) + (size_t)k * 64 * 4;
    auto end_of = [&](uint32_t bb) { return bb < nb ? end : 0u; };
    float4 r[8], rn[8];
    uint32_t b = warp;
    BatchIds ids = coop_load_ids(col, beg + 64 * b, end_of(b), lane);
    BatchIds idn = coop_load_ids(col, beg + 64 * (b + kWarpsPerCta), end_of(b + kWarpsPerCta), lane);
    coop_load_rows16(r, ids, in4, beg + 64 * b, end_of(b), lane);
    const int sv = lane >> 2, q = lane & 3;
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (; b < 64; b += kWarpsPerCta) {
        coop_load_rows16(rn, idn, in4, beg + 64 * (b + kWarpsPerCta), end_of(b + kWarpsPerCta), lane);
        idn = coop_load_ids(col, beg + 64 * (b + 2 * kWarpsPerCta), end_of(b + 2 * kWarpsPerCta), lane);
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < nb) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                __stcg(rows + (size_t)(64 * b + 8 * w + sv) * 4 + q, r[w]);
                sum.x += r[w].x; sum.y += r[w].y; sum.z += r[w].z; sum.w += r[w].w;
            }
#pragma unroll
            for (int m = 4; m < 32; m <<= 1) {
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, m); sum.y += __shfl_xor_sync(0xffffffffu, sum.y, m);
                sum.z += __shfl_xor_sync(0xffffffffu, sum.z, m); sum.w += __shfl_xor_sync(0xffffffffu, sum.w, m);
            }
        }
        if (lane < 4) __stcg(S4 + (size_t)b * 4 + q, sum);           // unused batch slots of a last chunk: 0
        tot.x += sum.x; tot.y += sum.y; tot.z += sum.z; tot.w += sum.w;
#pragma unroll
        for (int w = 0; w < 8; ++w) r[w] = rn[w];
    }
    if (lane < 4) *reinterpret_cast<float4 *>(sh + warp * 16 + 4 * q) = tot;
    __syncthreads();
    if (threadIdx.x < 16) {                                          // the chunk's sum per column
        float t = 0.0f;
        for (int w = 0; w < kWarpsPerCta; ++w) t += sh[w * 16 + threadIdx.x];
        __stcg(px.T + (size_t)k * 16 + threadIdx.x, t);
    }
}

// phase A, width 1: 256-entry batches, lane l holds entries 32 t + l
__device__ __noinline__ void px_gather1(const PxArgs &px, uint32_t k, const uint32_t *__restrict__ col,
                                        const float *__restrict__ x, float *__restrict__ sh, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t beg = ck.y, end = ck.z;
    float *vals = px.scratch + (size_t)k * kPxChunk;
    float *S = px.S + (size_t)k * 16;
    float tot = 0.0f;
#pragma unroll 1
    for (uint32_t b = warp; b < 16; b += kWarpsPerCta) {
        const uint32_t e0 = beg + 256 * b;
        uint32_t id[8];
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) { const uint32_t e = e0 + 32 * t + lane; id[t] = e < end ? ld_id(col + e) : 0xFFFFFFFFu; }
        float sum = 0.0f;
#pragma unroll
        for (int t = 0; t < 8; ++t) { v[t] = id[t] != 0xFFFFFFFFu ? __ldg(x + id[t]) : 0.0f; }
#pragma unroll
        for (int t = 0; t < 8; ++t) { __stcg(vals + 256 * b + 32 * t + lane, v[t]); sum += v[t]; }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
        if (lane == 0) __stcg(S + b, sum);
        tot += sum;
    }
    if (lane == 0) sh[warp] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < kWarpsPerCta; ++w) t += sh[w];
        __stcg(px.T + k, t);
    }
}

// ---- the scan over a hub's chunk sums: T[chunk] := sum of the chunks before it (double inside) -----------
// by warp 0 of whichever CTA finished the hub's last chunk; lane c = column (W = 1: lane 0)
template <int W>
__device__ __forceinline__ void px_scan_chunks(const PxArgs &px, uint32_t first, uint32_t count, int lane) {
    if (lane >= W) return;
    float *t = px.T + (size_t)first * W + lane;
    double run = 0.0;
    for (uint32_t i0 = 0; i0 < count; i0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = i0 + j < count ? __ldcg(t + (size_t)(i0 + j) * W) : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j < count) { __stcg(t + (size_t)(i0 + j) * W, (float)run); run += (double)v[j]; }
    }
}

// ---- phase B ----------------------------------------------------------------------------------------------
// sh: 64 x 16 floats.  First the entry value of every batch of the chunk (chunk entry + sums of the
// batches before it), then per batch D and the flag.
__device__ __noinline__ void px_quantise16(const PxArgs &px, uint32_t k, float *__restrict__ sh, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t nb = (ck.z - ck.y + 63) / 64;
    const float4 *rows = reinterpret_cast<const float4 *>(px.scratch) + (size_t)k * kPxChunk * 4;
    float4 *R4 = reinterpret_cast<float4 *>(px.rec) + (size_t)k * 64 * 8;
    // the chunk's 64 x 16 batch sums: one 128-bit load per thread, then 16 lanes walk the columns
    // (plain loads below where a 128-byte line belongs to this chunk alone and is complete before this CTA
    // may run -- nothing stale can sit in L1; ld.cg costs ~1400 cycles a piece here and does not overlap)
    reinterpret_cast<float4 *>(sh)[threadIdx.x] = (reinterpret_cast<const float4 *>(px.S) + (size_t)k * 64 * 4)[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 16) {
        float run = __ldcg(px.T + (size_t)k * 16 + threadIdx.x);     // entry value of the chunk
        for (int b = 0; b < 64; ++b) {
            const float v = sh[b * 16 + threadIdx.x];
            sh[b * 16 + threadIdx.x] = run;
            run += v;
        }
    }
    __syncthreads();
    const int sv = lane >> 2, q = lane & 3;
#pragma unroll 1
    for (uint32_t b = warp; b < nb; b += kWarpsPerCta) {
        float4 r[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) r[w] = rows[(size_t)(64 * b + 8 * w + sv) * 4 + q];
        float4 Pq = *reinterpret_cast<const float4 *>(sh + b * 16 + 4 * q);
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) { S.x += r[w].x; S.y += r[w].y; S.z += r[w].z; S.w += r[w].w; }
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {
            S.x += __shfl_xor_sync(0xffffffffu, S.x, m); S.y += __shfl_xor_sync(0xffffffffu, S.y, m);
            S.z += __shfl_xor_sync(0xffffffffu, S.z, m); S.w += __shfl_xor_sync(0xffffffffu, S.w, m);
        }
        // A column whose 64 addends are all zero adds nothing whatever the running sum is (dead ReLU units
        // are zero for EVERY vertex): always clean, no prediction needed; marked with P = -1 for the walk.
        uint32_t nz = 0;                                             // bit i: column 4q + i has a non-zero addend in this lane's rows
#pragma unroll
        for (int w = 0; w < 8; ++w)
            nz |= (r[w].x != 0.0f ? 1u : 0u) | (r[w].y != 0.0f ? 2u : 0u) | (r[w].z != 0.0f ? 4u : 0u) | (r[w].w != 0.0f ? 8u : 0u);
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) nz |= __shfl_xor_sync(0xffffffffu, nz, m);
        uint32_t mx = 0, my = 0, mz = 0, mw = 0;
        const bool cx = (nz & 1u) != 0, cy = (nz & 2u) != 0, cz = (nz & 4u) != 0, cw = (nz & 8u) != 0;   // columns that need a binade
        bool bad = cx && !px_predict(Pq.x, S.x, mx);
        bad |= cy && !px_predict(Pq.y, S.y, my);
        bad |= cz && !px_predict(Pq.z, S.z, mz);
        bad |= cw && !px_predict(Pq.w, S.w, mw);
        const float Mx = __uint_as_float(mx), My = __uint_as_float(my), Mz = __uint_as_float(mz), Mw = __uint_as_float(mw);
        const float hx = __uint_as_float(mx - (24u << 23)), hy = __uint_as_float(my - (24u << 23)),
                    hz = __uint_as_float(mz - (24u << 23)), hw = __uint_as_float(mw - (24u << 23));
        float4 Dq = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool unpredictable = __any_sync(0xffffffffu, bad);
        if (!bad) {                                                  // (garbage exponents otherwise; the batch is dirty anyway)
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (cx) Dq.x = __fadd_rn(Dq.x, px_quantum(r[w].x, Mx, hx, bad));
                if (cy) Dq.y = __fadd_rn(Dq.y, px_quantum(r[w].y, My, hy, bad));
                if (cz) Dq.z = __fadd_rn(Dq.z, px_quantum(r[w].z, Mz, hz, bad));
                if (cw) Dq.w = __fadd_rn(Dq.w, px_quantum(r[w].w, Mw, hw, bad));
            }
        }
        bad = __any_sync(0xffffffffu, bad);
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {                           // multiples of u below 2^(e+1): exact in any order
            Dq.x = __fadd_rn(Dq.x, __shfl_xor_sync(0xffffffffu, Dq.x, m)); Dq.y = __fadd_rn(Dq.y, __shfl_xor_sync(0xffffffffu, Dq.y, m));
            Dq.z = __fadd_rn(Dq.z, __shfl_xor_sync(0xffffffffu, Dq.z, m)); Dq.w = __fadd_rn(Dq.w, __shfl_xor_sync(0xffffffffu, Dq.w, m));
        }
        if (lane < 4) {                                              // the batch's record: {P[16], D[16]}
            if (!cx) Pq.x = -1.0f;
            if (!cy) Pq.y = -1.0f;
            if (!cz) Pq.z = -1.0f;
            if (!cw) Pq.w = -1.0f;
            __stcg(R4 + (size_t)b * 8 + q, Pq);
            __stcg(R4 + (size_t)b * 8 + 4 + q, Dq);
        }
        if (lane == 0) {
            __stcg(px.flag + (size_t)k * 64 + b, bad ? 1u : 0u);
            if (bad) atomicAdd(px.ctr + 7, 1u);                      // statistics
        }
    }
}

__device__ __noinline__ void px_quantise1(const PxArgs &px, uint32_t k, float *__restrict__ sh, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t nb = (ck.z - ck.y + 255) / 256;
    const float *vals = px.scratch + (size_t)k * kPxChunk;
    if (threadIdx.x < 16) sh[threadIdx.x] = __ldcg(px.S + (size_t)k * 16 + threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        float run = __ldcg(px.T + k);
        for (int b = 0; b < 16; ++b) { const float v = sh[b]; sh[b] = run; run += v; }
    }
    __syncthreads();
#pragma unroll 1
    for (uint32_t b = warp; b < nb; b += kWarpsPerCta) {
        float v[8], S = 0.0f;
#pragma unroll
        for (int t = 0; t < 8; ++t) { v[t] = __ldcg(vals + 256 * b + 32 * t + lane); S += v[t]; }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) S += __shfl_xor_sync(0xffffffffu, S, m);
        const float Pb = sh[b];
        bool nzl = false;
#pragma unroll
        for (int t = 0; t < 8; ++t) nzl |= v[t] != 0.0f;
        const bool nonzero = __any_sync(0xffffffffu, nzl);           // an all-zero batch adds nothing: clean without a binade
        uint32_t mb = 0;
        bool bad = nonzero && !px_predict(Pb, S, mb);
        const float M = __uint_as_float(mb), half_u = __uint_as_float(mb - (24u << 23));
        float D = 0.0f;
        if (!bad && nonzero) {
#pragma unroll
            for (int t = 0; t < 8; ++t) D = __fadd_rn(D, px_quantum(v[t], M, half_u, bad));
        }
        bad = __any_sync(0xffffffffu, bad);
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) D = __fadd_rn(D, __shfl_xor_sync(0xffffffffu, D, m));
        if (lane == 0) {
            __stcg(reinterpret_cast<float2 *>(px.rec) + (size_t)k * 16 + b, make_float2(nonzero ? Pb : -1.0f, D));
            __stcg(px.flag + (size_t)k * 16 + b, bad ? 1u : 0u);
        }
    }
}

// ---- phases A and B as seen by a CTA of the stage kernel: claim chunks until there are none left --------
// `claim` is a word of shared memory; sh holds 64 x 16 floats.  Ends with every thread of the CTA past
// a __syncthreads().
template <int W>
__device__ __forceinline__ void px_phases_ab(const PxArgs &px, const uint32_t *__restrict__ col, const float *__restrict__ in,
                                             uint32_t *claim, float *sh, int warp, int lane) {
    // A: gather
#pragma unroll 1
    for (;;) {
        if (threadIdx.x == 0) *claim = atomicAdd(px.ctr + 0, 1u);
        __syncthreads();
        const uint32_t k = *claim;
        if (k >= px.n_chunks) break;
        if constexpr (W == 16) px_gather16(px, k, col, reinterpret_cast<const float4 *>(in), sh, warp, lane);
        else px_gather1(px, k, col, in, sh, warp, lane);
        const uint32_t g = __ldg(&px.chunk[k].x);
        const uint2 hi = __ldg(px.info + g);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) *claim = atomicAdd(px.ctr + 8 + 3 * g, 1u);
        __syncthreads();
        if (*claim == hi.y - 1 && warp == 0) {                       // the hub's last chunk is in: entry values of its chunks
            __threadfence();
            px_scan_chunks<W>(px, hi.x, hi.y, lane);
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(px.ctr + 8 + 3 * g + 1, 1u);
        }
        __syncthreads();
    }
    // B: quantise
#pragma unroll 1
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) *claim = atomicAdd(px.ctr + 1, 1u);
        __syncthreads();
        const uint32_t k = *claim;
        if (k >= px.n_chunks) break;
        const uint32_t g = __ldg(&px.chunk[k].x);
        if (threadIdx.x == 0) px_wait(px.ctr + 8 + 3 * g + 1, 1u);
        __syncthreads();
        if constexpr (W == 16) px_quantise16(px, k, sh, warp, lane);
        else px_quantise1(px, k, sh, warp, lane);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(px.ctr + 8 + 3 * g + 2, 1u);
    }
    __syncthreads();
}

// a batch the slow way, width 16: its 64 rows staged column-major in the warp's tile buffer, then the chain
// (out of line on purpose: unrolled into the walk it pushed the loop out of the instruction cache and
// every batch, clean or not, paid ~1400 cycles of instruction fetch)
__device__ __noinline__ float px_slow16(float *__restrict__ T, const float4 *__restrict__ rows, uint32_t b, float acc, int lane) {
    float4 r[8];
    const int sv = lane >> 2, q = lane & 3;
#pragma unroll
    for (int w = 0; w < 8; ++w) r[w] = rows[(size_t)(64 * b + 8 * w + sv) * 4 + q];
    coop_stage_rows16(T, r, lane);
    __syncwarp();
    float4 va[kRingWindow];
    const float4 *s4 = reinterpret_cast<const float4 *>(T + (lane & 15) * kRingColStride);
#pragma unroll
    for (int t = 0; t < kRingWindow; ++t) va[t] = s4[t];
    acc = chain_add16_full(T, va, acc, lane);
    __syncwarp();
    return acc;
}
// the 64 rows of batch b (4 KB = 32 lines, one per lane) on their way into L1: no register, no scoreboard
__device__ __forceinline__ void px_prefetch_batch16(const float4 *__restrict__ rows, uint32_t b, int lane) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(rows + (size_t)64 * b * 4 + 8 * lane));
}

// ---- phase C: one warp walks the batches of hub g in order; returns the sums (lane c, c < W; mirrored) ---
// T: the warp's tile buffer in shared memory (slow batches are staged there, column-major, like the ring's).
// The walk is one dependent chain, so what counts is the latency per batch.  Measured on a 262 144-neighbour
// star (tools/chain_probe.py) a batch-by-batch loop cost 184 cycles per clean batch (shared-memory load,
// shuffle, vote and branch all behind the running sum) and ~1000 per dirty one (its rows were only
// requested when it was reached).  Hence:
//   * four batches are checked at once: their P / D come out of shared memory before the sum is needed,
//     the four dependent FADDs follow each other directly, ONE vote decides whether all four were clean
//     (else the quad is redone batch by batch, which is always right);
//   * the rows of the next batch flagged dirty by phase B are prefetched into L1 as soon as it is known to
//     be next, i.e. while the clean batches before it are walked.
// With this the 262 144-neighbour star walks at 119-138 cycles per clean batch (stage 2: 0.52 -> 0.41 ms), and
// the shard of an 8-GPU run that owns the 258 306-neighbour hub of R-MAT scale 23 takes 0.75 / 0.72 ms for
// stages 1 / 2 instead of 0.86 / 0.83 (tools/shard_probe.py; fast mode, which may split the sum: 0.55 / 0.50).
// Tried and measured slower (profiles/r2_walk_variants.json): a ring of record words in registers, sixteen
// batches ahead, P / D by shuffle -- next to a call ptxas keeps the ring in local memory, and a load that is
// stored straight to the stack is waited for at once (0.88 / 0.83); eight batches per vote with the records
// of eight batches in shared memory (0.84 / 0.81): the walk is bound by the instructions per batch that one
// warp gets issued next to three busy ones, not by a fixed cost per vote; four quads of records in flight in
// four named registers instead of one (1.03 / 1.01, and the whole scale-20 forward 1.13 -> 1.25 ms); the walk
// organised in blocks of 32 batches with the flag bookkeeping hoisted out of the quad loop (0.68 / 0.64 on that
// shard -- the best walk -- but the single-GPU scale-20 forward 1.13 -> 1.19 ms: more spills in the stage kernels).
__device__ __noinline__ float px_walk16(const PxArgs &px, uint32_t g, uint32_t deg, float *__restrict__ T, int lane) {
    const uint2 hi = __ldg(px.info + g);
    const long long t_in = clock64();
    px_wait(px.ctr + 8 + 3 * g + 2, hi.y);                           // all chunks quantised (every lane polls: no divergence)
    const long long t_go = clock64();
    const int c = lane & 15;
    const uint32_t nb = (deg + 63) / 64;
    // records of 4 batches at a time (4 x {P[16], D[16]} = 128 floats = one 128-bit load per lane), parked
    // behind the slow path's staging area in the warp's tile buffer; the next four are in flight
    const float4 *rec4 = reinterpret_cast<const float4 *>(px.rec + (size_t)hi.x * 64 * 32) + lane;
    const uint32_t *F = px.flag + (size_t)hi.x * 64;
    const float4 *rows = reinterpret_cast<const float4 *>(px.scratch) + (size_t)hi.x * kPxChunk * 4;
    float *R = T + 16 * kRingColStride;                              // 128 floats
    static_assert(16 * kRingColStride + 128 <= kWarpSmemFloats, "records of four batches must fit behind a parked batch");
    const uint32_t nq = (nb + 3) / 4;                                // quads of batches (the arrays are padded to whole chunks)
    float acc = 0.0f;
    uint32_t slow = 0;
    // The records were written by other SMs (st.cg): a load of them is an L2 access, 500-600 cycles under load,
    // and with only the next quad in flight that latency was the walk's pace (138 cycles per batch measured).
    // So the record stream is prefetched into L1 sixteen to twenty-four quads ahead -- one 128-byte line per
    // lane covers eight quads -- and the register copy of the next quad then comes out of L1.
    const float *rec_lines = px.rec + (size_t)hi.x * 64 * 32 + 32 * lane;      // lane's line of an 8-quad group (8 x 128 floats)
    auto prefetch_quads = [&](uint32_t q0) {                                   // quads [q0, q0 + 8)
        if (q0 < nq) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec_lines + (size_t)q0 * 128));
    };
    prefetch_quads(0); prefetch_quads(8); prefetch_quads(16);
    float4 rnext = rec4[0];
    // dirty: bit i = batch (32-block base) + i was flagged by phase B; the next block's flags are in flight
    uint32_t dirty = 0, fl_next = F[min((uint32_t)lane, nb - 1)];
    uint32_t pre_b = 0xFFFFFFFFu;                                    // the batch whose rows were requested ahead (L1 prefetch)
#pragma unroll 1
    for (uint32_t qd = 0; qd < nq; ++qd) {
        if ((qd & 7u) == 0) {
            dirty = __ballot_sync(0xffffffffu, fl_next != 0u);
            fl_next = F[min(4 * qd + 32 + lane, nb - 1)];
            prefetch_quads(qd + 24);
        }
        __syncwarp();
        reinterpret_cast<float4 *>(R)[lane] = rnext;
        rnext = rec4[(size_t)min(qd + 1, nq - 1) * 32];
        // the next dirty batch of this 32-block at or after this quad: get its rows under way
        const uint32_t ahead = dirty >> ((4 * qd) & 31u);
        if (ahead != 0u && pre_b == 0xFFFFFFFFu) {
            const uint32_t nd = 4 * qd + (uint32_t)__ffs((int)ahead) - 1;
            if (nd < nb) { pre_b = nd; px_prefetch_batch16(rows, nd, lane); }
        }
        __syncwarp();
        if ((ahead & 0xFu) == 0u && 4 * qd + 4 <= nb) {
            // all four at once
            const float P0 = R[c], D0 = R[16 + c], P1 = R[32 + c], D1 = R[48 + c];
            const float P2 = R[64 + c], D2 = R[80 + c], P3 = R[96 + c], D3 = R[112 + c];
            const uint32_t m0 = px_entry_binade(P0), m1 = px_entry_binade(P1), m2 = px_entry_binade(P2), m3 = px_entry_binade(P3);
            const float a1 = __fadd_rn(acc, D0), a2 = __fadd_rn(a1, D1), a3 = __fadd_rn(a2, D2), a4 = __fadd_rn(a3, D3);
            const bool ok = ((P0 < 0.0f) | (((__float_as_uint(acc) & 0x7F800000u) == m0) & (a1 < __uint_as_float(m0 + (1u << 23))))) &
                            ((P1 < 0.0f) | (((__float_as_uint(a1) & 0x7F800000u) == m1) & (a2 < __uint_as_float(m1 + (1u << 23))))) &
                            ((P2 < 0.0f) | (((__float_as_uint(a2) & 0x7F800000u) == m2) & (a3 < __uint_as_float(m2 + (1u << 23))))) &
                            ((P3 < 0.0f) | (((__float_as_uint(a3) & 0x7F800000u) == m3) & (a4 < __uint_as_float(m3 + (1u << 23)))));
            if (__all_sync(0xffffffffu, ok)) { acc = a4; continue; }
        }
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const uint32_t b = 4 * qd + j;
            if (b >= nb) break;
            const float Pj = R[32 * j + c], Dj = R[32 * j + 16 + c];
            const uint32_t Fj = (dirty >> (b & 31u)) & 1u;
            const uint32_t mb = px_entry_binade(Pj);
            const float nxt = __fadd_rn(acc, Dj);
            const bool ok = (Fj == 0u) & ((Pj < 0.0f) |              // an all-zero column: D = 0, nothing to verify
                                          (((__float_as_uint(acc) & 0x7F800000u) == mb) & (nxt < __uint_as_float(mb + (1u << 23)))));
            if (__all_sync(0xffffffffu, ok)) {
                acc = nxt;
            } else {                                                 // the reference's way: element by element
                ++slow;
                acc = px_slow16(T, rows, b, acc, lane);
                if (pre_b <= b) pre_b = 0xFFFFFFFFu;
            }
        }
    }
    if (lane == 0) {                                                 // statistics: slow batches, cycles waited / walked (>> 10)
        atomicAdd(px.ctr + 3, slow);
        atomicAdd(px.ctr + 5, (uint32_t)((t_go - t_in) >> 10)); atomicAdd(px.ctr + 6, (uint32_t)((clock64() - t_go) >> 10));
    }
    return acc;
}

__device__ __noinline__ float px_walk1(const PxArgs &px, uint32_t g, uint32_t deg, int lane) {
    const uint2 hi = __ldg(px.info + g);
    px_wait(px.ctr + 8 + 3 * g + 2, hi.y);                           // every lane polls: no divergence
    const uint32_t nb = (deg + 255) / 256;
    const float2 *rec = reinterpret_cast<const float2 *>(px.rec) + (size_t)hi.x * 16;
    const uint32_t *F = px.flag + (size_t)hi.x * 16;
    const float4 *vals = reinterpret_cast<const float4 *>(px.scratch + (size_t)hi.x * kPxChunk);
    float acc = 0.0f;                                                // every lane computes the same (loads are broadcasts)
    uint32_t slow = 0;
#pragma unroll 1
    for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
        const uint32_t bl = min(b0 + lane, nb - 1);                  // lane l holds the records of batch b0 + l
        const float2 Rl = __ldcg(rec + bl);
        const uint32_t Fl = __ldcg(F + bl);
        const int cnt = (int)min(32u, nb - b0);
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const float Pj = __shfl_sync(0xffffffffu, Rl.x, j), Dj = __shfl_sync(0xffffffffu, Rl.y, j);
            const uint32_t Fj = __shfl_sync(0xffffffffu, Fl, j);
            const uint32_t mb = px_entry_binade(Pj);
            const float nxt = __fadd_rn(acc, Dj);
            if (Fj == 0u && (Pj < 0.0f || ((__float_as_uint(acc) & 0x7F800000u) == mb && nxt < __uint_as_float(mb + (1u << 23))))) {
                acc = nxt;
            } else {
                ++slow;
                const float4 *vb = vals + (size_t)(b0 + j) * 64;
#pragma unroll 1
                for (int i = 0; i < 64; i += 4) {
                    const float4 a = __ldcg(vb + i), b = __ldcg(vb + i + 1), cc = __ldcg(vb + i + 2), d = __ldcg(vb + i + 3);
                    acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y); acc = __fadd_rn(acc, a.z); acc = __fadd_rn(acc, a.w);
                    acc = __fadd_rn(acc, b.x); acc = __fadd_rn(acc, b.y); acc = __fadd_rn(acc, b.z); acc = __fadd_rn(acc, b.w);
                    acc = __fadd_rn(acc, cc.x); acc = __fadd_rn(acc, cc.y); acc = __fadd_rn(acc, cc.z); acc = __fadd_rn(acc, cc.w);
                    acc = __fadd_rn(acc, d.x); acc = __fadd_rn(acc, d.y); acc = __fadd_rn(acc, d.z); acc = __fadd_rn(acc, d.w);
                }
            }
        }
    }
    if (lane == 0) atomicAdd(px.ctr + 3, slow);
    return acc;
}

}  // namespace gvc
