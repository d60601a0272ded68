// gather_bw.cu -- what can a B200 deliver for random 64-byte row gathers out of an L2-resident table?
// (ceiling for the neighbour aggregation of stages 1/2: 31.4 M rows of 64 B per stage on R-MAT scale 20)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu ; run: ./gather_bw
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int U>
__global__ void gather_kernel(const float4* __restrict__ table, const unsigned* __restrict__ ids, size_t m,
                              float4* __restrict__ out) {
    // 4 lanes per row; each sub-warp walks a contiguous slice of ids, U rows in flight per lane
    const size_t sub = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int q = threadIdx.x & 3;
    const size_t nsub = ((size_t)gridDim.x * blockDim.x) >> 2;
    const size_t per = (m + nsub - 1) / nsub;
    size_t e = sub * per, end = e + per < m ? e + per : m;
    float4 acc = make_float4(0, 0, 0, 0);
    for (; e + U <= end; e += U) {
        unsigned id[U];
        float4 r[U];
#pragma unroll
        for (int t = 0; t < U; ++t) id[t] = __ldg(ids + e + t);
#pragma unroll
        for (int t = 0; t < U; ++t) r[t] = __ldg(table + (size_t)id[t] * 4 + q);
#pragma unroll
        for (int t = 0; t < U; ++t) { acc.x += r[t].x; acc.y += r[t].y; acc.z += r[t].z; acc.w += r[t].w; }
    }
    if (acc.x == 12345.678f) out[sub * 4 + q] = acc;   // keep the loads alive
}

int main() {
    const size_t n = 1 << 20, m = 32u << 20;
    std::vector<unsigned> h(m);
    unsigned s = 12345;
    for (size_t i = 0; i < m; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % n; }
    float4* table; unsigned* ids; float4* out;
    cudaMalloc(&table, n * 64); cudaMalloc(&ids, m * 4); cudaMalloc(&out, 64 << 20);
    cudaMemset(table, 0, n * 64);
    cudaMemcpy(ids, h.data(), m * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](auto kern, const char* name, int blocks_per_sm, int threads) {
        for (int rep = 0; rep < 3; ++rep) {
            if (rep == 1) cudaEventRecord(a);
            kern<<<148 * blocks_per_sm, threads>>>(table, ids, m, out);
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 2;
        printf("%-12s blocks/SM=%d threads=%d : %.3f ms  rows %.1f G/s  row bytes %.0f GB/s (+ids %.0f GB/s)\n", name,
               blocks_per_sm, threads, ms, m / ms / 1e6, m * 64.0 / ms / 1e6, m * 4.0 / ms / 1e6);
    };
    for (int bps : {2, 4, 8}) {
        run(gather_kernel<4>, "U=4", bps, 256);
        run(gather_kernel<8>, "U=8", bps, 256);
        run(gather_kernel<16>, "U=16", bps, 256);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
