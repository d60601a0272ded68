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
//      copy (coalesced from then on) and leave one approximate sum per batch and column; whoever
//      finishes the last chunk of a hub turns those into prefix sums P (entry value of every batch)
//   B  again per chunk: D and a "clean" flag per batch, under the binade predicted from P
//   C  one warp per hub walks the batches in order: one FADD and three checks per clean batch, the
//      element-wise chain for the others
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
    float *P, *D;            // [n_chunks][batches per chunk][W]: predicted entry value / exact increment
    uint32_t *flag;          // [n_chunks][batches per chunk]: 1 = add this batch element by element
    uint32_t *ctr;           // [0..2] claim counters of phases A, B (CTAs); then per hub {A done, scan done, B done}
    uint32_t n_chunks, n_hubs;
};

// Binade of the running sum predicted for a batch that is entered at about P and adds about S:
// false if the interval [P, P + S] widened by 0.1 % touches a power of two, or P is not a positive
// normal number well inside the exponent range.  mbits = bit pattern of 2^e.
__device__ __forceinline__ bool px_predict(float P, float S, uint32_t &mbits) {
    const float lo = __fmul_rn(P, 0.999f), hi = __fmul_rn(__fadd_rn(P, S), 1.001f);
    mbits = __float_as_uint(lo) & 0x7F800000u;
    const uint32_t hb = __float_as_uint(hi) & 0x7F800000u;
    return lo > 0.0f && hi >= lo && mbits == hb && mbits >= (26u << 23) && mbits <= (252u << 23);
}
__device__ __forceinline__ uint32_t px_entry_binade(float P) { return __float_as_uint(__fmul_rn(P, 0.999f)) & 0x7F800000u; }

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
// 8w + sv; the next batch's rows and the ids of the one after are in flight)
__device__ __noinline__ void px_gather16(const PxArgs &px, uint32_t k, const uint32_t *__restrict__ col,
                                         const float4 *__restrict__ in4, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t beg = ck.y, end = ck.z;
    const uint32_t nb = (end - beg + 63) / 64;                       // batches that hold entries (<= 64)
    float4 *rows = reinterpret_cast<float4 *>(px.scratch) + (size_t)k * kPxChunk * 4;
    float4 *S4 = reinterpret_cast<float4 *>(px.P) + (size_t)k * 64 * 4;
    auto end_of = [&](uint32_t bb) { return bb < nb ? end : 0u; };
    float4 r[8], rn[8];
    uint32_t b = warp;
    BatchIds ids = coop_load_ids(col, beg + 64 * b, end_of(b), lane);
    BatchIds idn = coop_load_ids(col, beg + 64 * (b + kWarpsPerCta), end_of(b + kWarpsPerCta), lane);
    coop_load_rows16(r, ids, in4, beg + 64 * b, end_of(b), lane);
    const int sv = lane >> 2, q = lane & 3;
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
#pragma unroll
        for (int w = 0; w < 8; ++w) r[w] = rn[w];
    }
}

// phase A, width 1: 256-entry batches, lane l holds entries 32 t + l
__device__ __noinline__ void px_gather1(const PxArgs &px, uint32_t k, const uint32_t *__restrict__ col,
                                        const float *__restrict__ x, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t beg = ck.y, end = ck.z;
    float *vals = px.scratch + (size_t)k * kPxChunk;
    float *S = px.P + (size_t)k * 16;
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
    }
}

// ---- the scan: per-batch sums of a whole hub -> exclusive prefix sums, in place, by one CTA ------------
// slots [first, first + count) of `P`, `W` columns each (column-interleaved); sums run in double
template <int W>
__device__ __noinline__ void px_scan(float *__restrict__ P, size_t first, uint32_t count, double *__restrict__ sh /* kCtaThreads */) {
    constexpr int kSeg = kCtaThreads / W;                            // segments per column
    const int c = threadIdx.x % W, seg = threadIdx.x / W;
    const uint32_t per = (count + kSeg - 1) / kSeg;
    const uint32_t lo = min(count, (uint32_t)seg * per), hi = min(count, lo + per);
    float *base = P + first * W + c;
    double s = 0.0;
    for (uint32_t i = lo; i < hi; ++i) s += (double)__ldcg(base + (size_t)i * W);
    sh[threadIdx.x] = s;
    __syncthreads();
    double run = 0.0;
    for (int t = 0; t < seg; ++t) run += sh[t * W + c];
    for (uint32_t i = lo; i < hi; ++i) {
        const float v = __ldcg(base + (size_t)i * W);
        __stcg(base + (size_t)i * W, (float)run);
        run += (double)v;
    }
    __syncthreads();
}

// ---- phase B ----------------------------------------------------------------------------------------------
__device__ __noinline__ void px_quantise16(const PxArgs &px, uint32_t k, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t nb = (ck.z - ck.y + 63) / 64;
    const float4 *rows = reinterpret_cast<const float4 *>(px.scratch) + (size_t)k * kPxChunk * 4;
    const float4 *P4 = reinterpret_cast<const float4 *>(px.P) + (size_t)k * 64 * 4;
    float4 *D4 = reinterpret_cast<float4 *>(px.D) + (size_t)k * 64 * 4;
    const int sv = lane >> 2, q = lane & 3;
#pragma unroll 1
    for (uint32_t b = warp; b < nb; b += kWarpsPerCta) {
        float4 r[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) r[w] = __ldcg(rows + (size_t)(64 * b + 8 * w + sv) * 4 + q);
        const float4 Pq = __ldcg(P4 + (size_t)b * 4 + q);
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) { S.x += r[w].x; S.y += r[w].y; S.z += r[w].z; S.w += r[w].w; }
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {
            S.x += __shfl_xor_sync(0xffffffffu, S.x, m); S.y += __shfl_xor_sync(0xffffffffu, S.y, m);
            S.z += __shfl_xor_sync(0xffffffffu, S.z, m); S.w += __shfl_xor_sync(0xffffffffu, S.w, m);
        }
        uint32_t mx, my, mz, mw;
        bool bad = !px_predict(Pq.x, S.x, mx);
        bad |= !px_predict(Pq.y, S.y, my);
        bad |= !px_predict(Pq.z, S.z, mz);
        bad |= !px_predict(Pq.w, S.w, mw);
        const float Mx = __uint_as_float(mx), My = __uint_as_float(my), Mz = __uint_as_float(mz), Mw = __uint_as_float(mw);
        const float hx = __uint_as_float(mx - (24u << 23)), hy = __uint_as_float(my - (24u << 23)),
                    hz = __uint_as_float(mz - (24u << 23)), hw = __uint_as_float(mw - (24u << 23));
        float4 Dq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!bad) {                                                  // (garbage exponents otherwise; the batch is dirty anyway)
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                Dq.x = __fadd_rn(Dq.x, px_quantum(r[w].x, Mx, hx, bad));
                Dq.y = __fadd_rn(Dq.y, px_quantum(r[w].y, My, hy, bad));
                Dq.z = __fadd_rn(Dq.z, px_quantum(r[w].z, Mz, hz, bad));
                Dq.w = __fadd_rn(Dq.w, px_quantum(r[w].w, Mw, hw, bad));
            }
        }
        bad = __any_sync(0xffffffffu, bad);
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {                           // multiples of u below 2^(e+1): exact in any order
            Dq.x = __fadd_rn(Dq.x, __shfl_xor_sync(0xffffffffu, Dq.x, m)); Dq.y = __fadd_rn(Dq.y, __shfl_xor_sync(0xffffffffu, Dq.y, m));
            Dq.z = __fadd_rn(Dq.z, __shfl_xor_sync(0xffffffffu, Dq.z, m)); Dq.w = __fadd_rn(Dq.w, __shfl_xor_sync(0xffffffffu, Dq.w, m));
        }
        if (lane < 4) __stcg(D4 + (size_t)b * 4 + q, Dq);
        if (lane == 0) __stcg(px.flag + (size_t)k * 64 + b, bad ? 1u : 0u);
    }
}

__device__ __noinline__ void px_quantise1(const PxArgs &px, uint32_t k, int warp, int lane) {
    const uint4 ck = __ldg(px.chunk + k);
    const uint32_t nb = (ck.z - ck.y + 255) / 256;
    const float *vals = px.scratch + (size_t)k * kPxChunk;
#pragma unroll 1
    for (uint32_t b = warp; b < nb; b += kWarpsPerCta) {
        float v[8], S = 0.0f;
#pragma unroll
        for (int t = 0; t < 8; ++t) { v[t] = __ldcg(vals + 256 * b + 32 * t + lane); S += v[t]; }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) S += __shfl_xor_sync(0xffffffffu, S, m);
        const float Pb = __ldcg(px.P + (size_t)k * 16 + b);
        uint32_t mb;
        bool bad = !px_predict(Pb, S, mb);
        const float M = __uint_as_float(mb), half_u = __uint_as_float(mb - (24u << 23));
        float D = 0.0f;
        if (!bad) {
#pragma unroll
            for (int t = 0; t < 8; ++t) D = __fadd_rn(D, px_quantum(v[t], M, half_u, bad));
        }
        bad = __any_sync(0xffffffffu, bad);
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) D = __fadd_rn(D, __shfl_xor_sync(0xffffffffu, D, m));
        if (lane == 0) { __stcg(px.D + (size_t)k * 16 + b, D); __stcg(px.flag + (size_t)k * 16 + b, bad ? 1u : 0u); }
    }
}

// ---- phases A and B as seen by a CTA of the stage kernel: claim chunks until there are none left --------
// `claim` is a word of shared memory; sh holds kCtaThreads doubles (the scan).  Ends with every thread
// of the CTA past a __syncthreads().
template <int W>
__device__ __forceinline__ void px_phases_ab(const PxArgs &px, const uint32_t *__restrict__ col, const float *__restrict__ in,
                                             uint32_t *claim, double *sh, int warp, int lane) {
    constexpr int kBpc = px_batches_per_chunk<W>();
    // A: gather
#pragma unroll 1
    for (;;) {
        if (threadIdx.x == 0) *claim = atomicAdd(px.ctr + 0, 1u);
        __syncthreads();
        const uint32_t k = *claim;
        if (k >= px.n_chunks) break;
        if constexpr (W == 16) px_gather16(px, k, col, reinterpret_cast<const float4 *>(in), warp, lane);
        else px_gather1(px, k, col, in, warp, lane);
        const uint32_t g = __ldg(&px.chunk[k].x);
        const uint2 hi = __ldg(px.info + g);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) *claim = atomicAdd(px.ctr + 4 + 3 * g, 1u);
        __syncthreads();
        if (*claim == hi.y - 1) {                                    // the hub's last chunk is in: its prefix sums
            __threadfence();
            px_scan<W>(px.P, (size_t)hi.x * kBpc, hi.y * kBpc, sh);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(px.ctr + 4 + 3 * g + 1, 1u);
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
        if (threadIdx.x == 0) px_wait(px.ctr + 4 + 3 * g + 1, 1u);
        __syncthreads();
        if constexpr (W == 16) px_quantise16(px, k, warp, lane);
        else px_quantise1(px, k, warp, lane);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(px.ctr + 4 + 3 * g + 2, 1u);
    }
    __syncthreads();
}

// ---- phase C: one warp walks the batches of hub g in order; returns the sums (lane c, c < W; mirrored) ---
__device__ __noinline__ float px_walk16(const PxArgs &px, uint32_t g, uint32_t deg, int lane) {
    const uint2 hi = __ldg(px.info + g);
    if (lane == 0) px_wait(px.ctr + 4 + 3 * g + 2, hi.y);            // all chunks quantised
    __syncwarp();
    __threadfence();
    const int c = lane & 15;
    const uint32_t nb = (deg + 63) / 64;
    const float *P = px.P + (size_t)hi.x * 64 * 16 + c, *D = px.D + (size_t)hi.x * 64 * 16 + c;
    const uint32_t *F = px.flag + (size_t)hi.x * 64;
    const float *rows = px.scratch + (size_t)hi.x * kPxChunk * 16 + c;
    float acc = 0.0f;
    constexpr int G = 4;                                             // batches per group, the next group's records in flight
    float Pn[G], Dn[G];
    uint32_t Fn[G];
    auto load = [&](uint32_t b0) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t b = min(b0 + j, nb - 1);
            Pn[j] = __ldcg(P + (size_t)b * 16); Dn[j] = __ldcg(D + (size_t)b * 16); Fn[j] = __ldcg(F + b);
        }
    };
    load(0);
#pragma unroll 1
    for (uint32_t b0 = 0; b0 < nb; b0 += G) {
        float Pc[G], Dc[G];
        uint32_t Fc[G];
#pragma unroll
        for (int j = 0; j < G; ++j) { Pc[j] = Pn[j]; Dc[j] = Dn[j]; Fc[j] = Fn[j]; }
        load(b0 + G);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t b = b0 + j;
            if (b >= nb) break;
            const uint32_t mb = px_entry_binade(Pc[j]);
            const float nxt = __fadd_rn(acc, Dc[j]);
            const bool ok = Fc[j] == 0u && (__float_as_uint(acc) & 0x7F800000u) == mb &&
                            nxt < __uint_as_float(mb + (1u << 23));
            if (__all_sync(0xffffffffu, ok)) {
                acc = nxt;
            } else {                                                 // the reference's way: element by element
                const float *rb = rows + (size_t)b * 64 * 16;
#pragma unroll 1
                for (int i = 0; i < 64; i += 16) {
                    float v[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = __ldcg(rb + (size_t)(i + t) * 16);
#pragma unroll
                    for (int t = 0; t < 16; ++t) acc = __fadd_rn(acc, v[t]);
                }
            }
        }
    }
    return acc;
}

__device__ __noinline__ float px_walk1(const PxArgs &px, uint32_t g, uint32_t deg, int lane) {
    const uint2 hi = __ldg(px.info + g);
    if (lane == 0) px_wait(px.ctr + 4 + 3 * g + 2, hi.y);
    __syncwarp();
    __threadfence();
    const uint32_t nb = (deg + 255) / 256;
    const float *P = px.P + (size_t)hi.x * 16, *D = px.D + (size_t)hi.x * 16;
    const uint32_t *F = px.flag + (size_t)hi.x * 16;
    const float4 *vals = reinterpret_cast<const float4 *>(px.scratch + (size_t)hi.x * kPxChunk);
    float acc = 0.0f;                                                // every lane computes the same (loads are broadcasts)
#pragma unroll 1
    for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
        const uint32_t bl = min(b0 + lane, nb - 1);                  // lane l holds the records of batch b0 + l
        const float Pl = __ldcg(P + bl), Dl = __ldcg(D + bl);
        const uint32_t Fl = __ldcg(F + bl);
        const int cnt = (int)min(32u, nb - b0);
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const float Pj = __shfl_sync(0xffffffffu, Pl, j), Dj = __shfl_sync(0xffffffffu, Dl, j);
            const uint32_t Fj = __shfl_sync(0xffffffffu, Fl, j);
            const uint32_t mb = px_entry_binade(Pj);
            const float nxt = __fadd_rn(acc, Dj);
            if (Fj == 0u && (__float_as_uint(acc) & 0x7F800000u) == mb && nxt < __uint_as_float(mb + (1u << 23))) {
                acc = nxt;
            } else {
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
    return acc;
}

}  // namespace gvc
