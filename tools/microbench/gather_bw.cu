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

// software-pipelined: ids two chunks ahead, rows one chunk ahead (what the stage kernel does)
template <int U>
__global__ void gather_pipelined(const float4* __restrict__ table, const unsigned* __restrict__ ids, size_t m,
                                 float4* __restrict__ out) {
    const size_t sub = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int q = threadIdx.x & 3;
    const size_t nsub = ((size_t)gridDim.x * blockDim.x) >> 2;
    const size_t per = (m + nsub - 1) / nsub / (2 * U) * (2 * U);
    size_t e = sub * per, end = e + per < m ? e + per : m;
    if (e + 4 * U > end) return;
    float4 acc = make_float4(0, 0, 0, 0);
    unsigned idA[U], idB[U];
    float4 rA[U], rB[U];
#pragma unroll
    for (int t = 0; t < U; ++t) { idA[t] = __ldg(ids + e + t); idB[t] = __ldg(ids + e + U + t); }
#pragma unroll
    for (int t = 0; t < U; ++t) rA[t] = __ldg(table + (size_t)idA[t] * 4 + q);
    for (; e + 4 * U <= end; e += 2 * U) {
#pragma unroll
        for (int t = 0; t < U; ++t) rB[t] = __ldg(table + (size_t)idB[t] * 4 + q);
#pragma unroll
        for (int t = 0; t < U; ++t) idA[t] = __ldg(ids + e + 2 * U + t);
#pragma unroll
        for (int t = 0; t < U; ++t) { acc.x += rA[t].x; acc.y += rA[t].y; acc.z += rA[t].z; acc.w += rA[t].w; }
#pragma unroll
        for (int t = 0; t < U; ++t) rA[t] = __ldg(table + (size_t)idA[t] * 4 + q);
#pragma unroll
        for (int t = 0; t < U; ++t) idB[t] = __ldg(ids + e + 3 * U + t);
#pragma unroll
        for (int t = 0; t < U; ++t) { acc.x += rB[t].x; acc.y += rB[t].y; acc.z += rB[t].z; acc.w += rB[t].w; }
    }
    if (acc.x == 12345.678f) out[sub * 4 + q] = acc;
}

// cp.async ring in shared memory: D rows in flight per sub-warp, no registers held by loads in flight
template <int D>
__global__ void gather_cpasync(const float4* __restrict__ table, const unsigned* __restrict__ ids, size_t m,
                               float4* __restrict__ out) {
    extern __shared__ float4 ring[];   // [threads/4][D][4]
    const size_t sub = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int q = threadIdx.x & 3;
    float4* my = ring + (size_t)(threadIdx.x >> 2) * D * 4 + q;
    const size_t nsub = ((size_t)gridDim.x * blockDim.x) >> 2;
    const size_t per = (m + nsub - 1) / nsub;
    size_t e = sub * per, end = e + per < m ? e + per : m;
    float4 acc = make_float4(0, 0, 0, 0);
    // prologue: D rows in flight, one commit group per row
    size_t issued = e;
    for (int d = 0; d < D && issued < end; ++d, ++issued) {
        const unsigned id = __ldg(ids + issued);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(my + d * 4);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(table + (size_t)id * 4 + q));
        asm volatile("cp.async.commit_group;");
    }
    int slot = 0;
    for (; e < end; ++e) {
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
        const float4 r = my[slot * 4];
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
        if (issued < end) {
            const unsigned id = __ldg(ids + issued);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(my + slot * 4);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(table + (size_t)id * 4 + q));
            ++issued;
        }
        asm volatile("cp.async.commit_group;");
        slot = slot + 1 == D ? 0 : slot + 1;
    }
    if (acc.x == 12345.678f) out[sub * 4 + q] = acc;
}

int main(int argc, char** argv) {
    // GATHER_ROWS=8388608: a 512 MB table, four times the L2 -- the regime of the 100 M-edge graph (R-MAT scale 23)
    const size_t n = getenv("GATHER_ROWS") ? (size_t)atoll(getenv("GATHER_ROWS")) : (size_t)1 << 20;
    size_t m = 32u << 20;
    printf("table: %zu rows of 64 B = %.0f MB\n", n, n * 64.0 / 1e6);
    std::vector<unsigned> h(m);
    unsigned s = 12345;
    for (size_t i = 0; i < m; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % n; }
    if (argc > 1) {   // ids from a file (uint32): e.g. the col array of the R-MAT benchmark graph
        FILE* f = fopen(argv[1], "rb");
        if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
        m = fread(h.data(), 4, m, f) / 1024 * 1024;
        fclose(f);
        printf("ids from %s: %zu entries\n", argv[1], m);
    }
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
    for (int bps : {2, 4}) {
        run(gather_pipelined<4>, "pipe U=4", bps, 256);
        run(gather_pipelined<8>, "pipe U=8", bps, 256);
    }
    auto run_cp = [&](auto kern, const char* name, int blocks_per_sm, int threads, int depth) {
        const size_t smem = (size_t)threads / 4 * depth * 64;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 3; ++rep) {
            if (rep == 1) cudaEventRecord(a);
            kern<<<148 * blocks_per_sm, threads, smem>>>(table, ids, m, out);
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 2;
        printf("%-12s blocks/SM=%d threads=%d smem=%zu KB : %.3f ms  rows %.1f G/s  row bytes %.0f GB/s\n", name,
               blocks_per_sm, threads, smem / 1024, ms, m / ms / 1e6, m * 64.0 / ms / 1e6);
    };
    run_cp(gather_cpasync<8>, "cpasync D=8", 2, 256, 8);
    run_cp(gather_cpasync<16>, "cpasync D=16", 2, 256, 16);
    run_cp(gather_cpasync<32>, "cpasync D=32", 2, 256, 32);
    run_cp(gather_cpasync<16>, "cpasync D=16", 4, 256, 16);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
