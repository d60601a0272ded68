// How fast can one warp run a chain of dependent fp32 adds?  (the exact-mode hub sums)
//   A: operands in registers                      -> pure FADD dependent-issue latency
//   B: operands streamed from shared memory with 128-bit broadcast loads, two loads ahead
//   C: B, while 7 more warps of the CTA wait at a named barrier
//   D: B, while 7 more warps of the CTA spin on loads from global memory (busy neighbours),
//      with 2..16 shared-memory loads ahead and with fewer gathering warps
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_lat chain_lat.cu ; run: ./chain_lat
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;   // x 256 adds

__global__ void chain_regs(float *out, float seed) {
    float v[16];
    for (int k = 0; k < 16; ++k) v[k] = seed * (k + 1);
    float acc = 0.f;
    for (int it = 0; it < kIters * 16; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = __fadd_rn(acc, v[k]);
    }
    out[threadIdx.x] = acc;
}

template <int MODE, int AHEAD = 2>
__global__ void chain_smem(float *out, const float *g, float seed) {
    __shared__ __align__(16) float S[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        for (int k = lane; k < 256; k += 32) S[k] = seed * (k + 1);
        __syncwarp();
        const float4 *s4 = reinterpret_cast<const float4 *>(S);
        float acc = 0.f;
#pragma unroll 1
        for (int it = 0; it < kIters; ++it) {
            float4 q[AHEAD];                       // AHEAD 128-bit loads (4 adds each) in flight
#pragma unroll
            for (int j = 0; j < AHEAD; ++j) q[j] = s4[j];
#pragma unroll 16
            for (int k = 0; k < 64; ++k) {
                const float4 a = q[0];
#pragma unroll
                for (int j = 0; j + 1 < AHEAD; ++j) q[j] = q[j + 1];
                q[AHEAD - 1] = s4[(k + AHEAD) & 63];
                acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y);
                acc = __fadd_rn(acc, a.z); acc = __fadd_rn(acc, a.w);
            }
        }
        out[lane] = acc;
        if (MODE == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
        if (MODE == 2) *reinterpret_cast<volatile float *>(&S[0]) = -1.f;
    } else if (MODE == 1) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
    } else if (MODE == 2) {
        float t = 0.f;
        size_t i = threadIdx.x;
        while (*reinterpret_cast<volatile float *>(&S[0]) != -1.f) {
            for (int r = 0; r < 64; ++r) { t += __ldcg(g + i); i = (i * 1664525u + 1013904223u) & ((1u << 24) - 1); }
        }
        if (t == 123.f) out[threadIdx.x] = t;
    }
}

int main() {
    float *out, *g;
    cudaMalloc(&out, 4096);
    cudaMalloc(&g, sizeof(float) << 24);
    cudaMemset(g, 0, sizeof(float) << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double adds = (double)kIters * 256;
    auto report = [&](const char *name, float ms) { printf("%-44s %.3f ms  %.2f ns/add\n", name, ms, ms * 1e6 / adds); };
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); chain_regs<<<1, 32>>>(out, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("A registers", ms);
        cudaEventRecord(e0); chain_smem<0><<<1, 32>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("B shared memory, LDS.128 two ahead", ms);
        cudaEventRecord(e0); chain_smem<1><<<1, 256>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("C + 7 warps waiting at a named barrier", ms);
        cudaEventRecord(e0); chain_smem<2><<<1, 256>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("D + 7 warps gathering from global memory", ms);
        cudaEventRecord(e0); chain_smem<2, 4><<<1, 256>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("D with 4 loads ahead", ms);
        cudaEventRecord(e0); chain_smem<2, 8><<<1, 256>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("D with 8 loads ahead", ms);
        cudaEventRecord(e0); chain_smem<2, 16><<<1, 256>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("D with 16 loads ahead", ms);
        cudaEventRecord(e0); chain_smem<2, 2><<<1, 64>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("D with 1 gathering warp only (other SMSP)", ms);
        cudaEventRecord(e0); chain_smem<2, 2><<<1, 160>>>(out, g, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("D with 4 gathering warps (one on the chain's SMSP)", ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
