// fp32_rate.cu -- what bounds the exact-mode dense chain?  One [32 vertices x 32] . [32 x 32] layer per
// warp tile out of shared memory, in place, repeated; four ways to do the multiply-adds:
//   mul_add   FMUL + FADD per product, as the exact mode does today (reference order: product and sum
//             rounded separately, src/matrix.cpp:112 through OpenBLAS' SSE kernel)
//   ffma      one FFMA per product (fast mode before the tensor cores)
//   ffma2     packed FFMA2, two products per instruction (fast)
//   mul_add2  the exact arithmetic on packed instructions: p = fma2(a, w, -0.0) is the correctly rounded
//             product (x*y + -0 == x*y for every x*y incl. both zeros), acc = fma2(p, 1.0, acc) the correctly
//             rounded sum -- bit-identical to FMUL + FADD, half the instructions.  (mul.rn.f32x2 followed by
//             add.rn.f32x2 is NOT usable: ptxas 12.9 contracts the pair into one FFMA2.)
// Prints ns per tile-layer per SM-resident warp and the MAC rate.  Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/fp32_rate tools/microbench/fp32_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kStride = 36;

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// MODE 0 mul_add, 1 ffma, 2 ffma2, 3 mul_add2
template <int MODE>
__global__ void __launch_bounds__(256, 2) layer_kernel(const float *__restrict__ Wg, float *__restrict__ out, int reps, float one, float nzero) {
    extern __shared__ float smem[];
    float *Wsm = smem;                        // 32 x 32 (MODE >= 2: every weight twice, 32 x 64)
    float *T = smem + 2048 + (threadIdx.x >> 5) * (32 * kStride);
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        if (MODE >= 2) { Wsm[2 * i] = Wg[i]; Wsm[2 * i + 1] = Wg[i]; }
        else Wsm[i] = Wg[i];
    }
    for (int k = 0; k < 32; ++k) T[k * kStride + lane] = 0.001f * (float)((lane * 7 + k * 3 + blockIdx.x) % 97);
    __syncthreads();
    const int og = lane >> 3, vg = lane & 7;
    for (int rep = 0; rep < reps; ++rep) {
        if constexpr (MODE < 2) {
            float acc[4][8];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
#pragma unroll 2
            for (int k = 0; k < 32; ++k) {
                const float4 a = *reinterpret_cast<const float4 *>(T + k * kStride + 4 * vg);
                const float4 w0 = *reinterpret_cast<const float4 *>(Wsm + k * 32 + 8 * og);
                const float4 w1 = *reinterpret_cast<const float4 *>(Wsm + k * 32 + 8 * og + 4);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        acc[r][c] = MODE == 0 ? __fadd_rn(acc[r][c], __fmul_rn(av[r], w[c])) : fmaf(av[r], w[c], acc[r][c]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 o;
                o.x = fminf(acc[0][c], 1.0f); o.y = fminf(acc[1][c], 1.0f); o.z = fminf(acc[2][c], 1.0f); o.w = fminf(acc[3][c], 1.0f);
                *reinterpret_cast<float4 *>(T + (8 * og + c) * kStride + 4 * vg) = o;
            }
            __syncwarp();
        } else {
            // pairs along the vertices: acc2[h][c] = outputs c of vertices (2h, 2h + 1); the activations come as
            // natural pairs out of the 128-bit load, the weights are stored twice
            unsigned long long acc2[2][8];
            // 1.0 and -0.0 arrive as kernel arguments: constants would let ptxas fold the pair back into ONE FFMA2
            const unsigned long long one2 = pack2(one, one), nzero2 = pack2(nzero, nzero);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc2[h][c] = 0ull;
#pragma unroll 2
            for (int k = 0; k < 32; ++k) {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(T + k * kStride + 4 * vg);
                unsigned long long w[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const ulonglong2 ww = *reinterpret_cast<const ulonglong2 *>(Wsm + k * 64 + 16 * og + 4 * j);
                    w[2 * j] = ww.x; w[2 * j + 1] = ww.y;
                }
                const unsigned long long av[2] = {a.x, a.y};
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if constexpr (MODE == 2) acc2[h][c] = fma2(av[h], w[c], acc2[h][c]);
                        else acc2[h][c] = fma2(fma2(av[h], w[c], nzero2), one2, acc2[h][c]);
                    }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 o;
                unpack2(acc2[0][c], o.x, o.y);
                unpack2(acc2[1][c], o.z, o.w);
                o.x = fminf(o.x, 1.0f); o.y = fminf(o.y, 1.0f); o.z = fminf(o.z, 1.0f); o.w = fminf(o.w, 1.0f);
                *reinterpret_cast<float4 *>(T + (8 * og + c) * kStride + 4 * vg) = o;
            }
            __syncwarp();
        }
    }
    float s = 0.0f;
    for (int k = 0; k < 32; ++k) s += T[k * kStride + lane];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const float *dW, float *dout, int reps, const char *name, std::vector<float> *keep) {
    const int ctas = 148 * 2, smem = (2048 + 8 * 32 * kStride) * sizeof(float);
    cudaFuncSetAttribute(layer_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    layer_kernel<MODE><<<ctas, 256, smem>>>(dW, dout, 10, 1.0f, -0.0f);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    layer_kernel<MODE><<<ctas, 256, smem>>>(dW, dout, reps, 1.0f, -0.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double tiles = (double)ctas * 8 * reps, macs = tiles * 32 * 32 * 32;
    std::printf("%-9s %8.3f ms  %6.1f ns per tile-layer per SM  %6.2f T MAC/s  (%s)\n", name, ms, ms * 1e6 / (tiles / 148), macs / ms * 1e-9,
                cudaGetErrorString(cudaGetLastError()));
    if (keep) { keep->resize((size_t)ctas * 256); cudaMemcpy(keep->data(), dout, keep->size() * 4, cudaMemcpyDeviceToHost); }
    return ms;
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? std::atoi(argv[1]) : 2000;
    std::vector<float> W(1024);
    for (int i = 0; i < 1024; ++i) W[i] = 0.03f * (float)((i * 37 % 61) - 30) / 30.0f;
    float *dW, *dout;
    cudaMalloc(&dW, 4096);
    cudaMalloc(&dout, 148 * 2 * 256 * 4);
    cudaMemcpy(dW, W.data(), 4096, cudaMemcpyHostToDevice);
    std::vector<float> r0, r3, r1, r2;
    run<0>(dW, dout, reps, "mul_add", &r0);
    run<1>(dW, dout, reps, "ffma", &r1);
    run<2>(dW, dout, reps, "ffma2", &r2);
    run<3>(dW, dout, reps, "mul_add2", &r3);
    size_t diff = 0, diff_f = 0;
    for (size_t i = 0; i < r0.size(); ++i) { diff += r0[i] != r3[i]; diff_f += r1[i] != r2[i]; }
    std::printf("mul_add2 vs mul_add: %zu of %zu outputs differ (must be 0); ffma2 vs ffma: %zu differ\n", diff, r0.size(), diff_f);
    return diff != 0;
}
