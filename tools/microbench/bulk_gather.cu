// bulk_gather.cu -- can the 64-byte neighbour rows be gathered with the bulk-copy engine?
// MEASURED in round 2 (profiles/r2_microbench_bulk_gather_rmat_ids.txt, the id stream of the R-MAT
// benchmark graph): NO.  3.0 TB/s of row bytes with 2-4 batches in flight per warp, 1.6 TB/s with 8
// (LDG.128 gathers of the same ids: 9-13 TB/s, gather_bw.cu), and 4.7 ns per neighbour for a single
// hub on one CTA (the exact chain runs at 2.9-3.4).  One 64-byte copy per bulk instruction is far below
// the granularity the engine is built for; the row gathers stay on LDG.128.
//
// Question: can the 64-byte neighbour rows be gathered with the bulk-copy engine
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx, "TMA 1-D") instead of LDG.128 per
// lane?  A row is 64 contiguous, 16-byte aligned bytes -- a legal bulk copy.  What it would buy:
// no registers for rows in flight, no LSU instruction per lane (tools/microbench/chain_lat.cu shows
// that the exact-mode hub chain drops from 2.1 to 5.6 ns per add when its SM neighbours issue
// gathers), arbitrarily deep prefetch limited only by shared memory.
//
//   A  whole GPU: every warp owns a contiguous slice of the id array, a ring of STAGES batches of
//      32 rows (2 KB) per warp, lane l issues the copy of row l of a batch, the warp waits on the
//      batch's mbarrier and adds the rows up (fast-mode style reduction: 4 lanes per row)
//   B  one CTA, one "hub": the same with all 8 warps of a single CTA on one long list
//      (compare tools/chain_probe.py: 1.8 ns per neighbour in fast mode before the hub chunks)
// The LDG baseline for A is gather_bw.cu (9-10 TB/s pipelined at 16 warps per SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_gather bulk_gather.cu ; run: ./bulk_gather [ids.u32]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_row(void *dst_smem, const void *src_gmem, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(smem_u32(bar)) : "memory");
}

// One warp: sum the rows table[ids[e]] for e in [beg, end) (end - beg a multiple of 32), STAGES batches in flight.
template <int STAGES>
__device__ __forceinline__ float4 warp_bulk_sum(const float4 *__restrict__ table, const unsigned *__restrict__ ids,
                                                size_t beg, size_t end, float4 *ring /* STAGES x 32 rows x 4 float4 */,
                                                uint64_t *bars /* STAGES */, int lane) {
    const int sv = lane >> 2, q = lane & 3;
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t nb = (end - beg) / 32;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](size_t b) {
        const int s = (int)(b % STAGES);
        if (lane == 0) mbar_expect_tx(bars + s, 32 * 64);
        __syncwarp();
        const unsigned id = __ldg(ids + beg + 32 * b + lane);
        bulk_row(ring + ((size_t)s * 32 + lane) * 4, table + (size_t)id * 4, bars + s);
    };
    for (size_t b = 0; b < (size_t)STAGES && b < nb; ++b) issue(b);
    for (size_t b = 0; b < nb; ++b) {
        const int s = (int)(b % STAGES);
        mbar_wait(bars + s, (uint32_t)((b / STAGES) & 1));
        const float4 *rows = ring + (size_t)s * 32 * 4;
#pragma unroll
        for (int w = 0; w < 4; ++w) {                          // 8 rows per step, 4 lanes per row
            const float4 r = rows[(8 * w + sv) * 4 + q];
            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
        }
        __syncwarp();                                          // everybody has read the slot
        if (b + STAGES < nb) issue(b + STAGES);
    }
    return acc;
}

template <int STAGES>
__global__ void bulk_all(const float4 *__restrict__ table, const unsigned *__restrict__ ids, size_t m, float4 *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float4 *ring = reinterpret_cast<float4 *>(smem) + (size_t)warp * STAGES * 32 * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * STAGES * 32 * 64) + warp * STAGES;
    const size_t gw = (size_t)blockIdx.x * nw + warp, tw = (size_t)gridDim.x * nw;
    const size_t per = m / tw / 32 * 32;
    const float4 acc = warp_bulk_sum<STAGES>(table, ids, gw * per, gw * per + per, ring, bars, lane);
    if (acc.x == 12345.678f) out[gw * 32 + lane] = acc;
}

template <int STAGES>
__global__ void bulk_hub(const float4 *__restrict__ table, const unsigned *__restrict__ ids, size_t deg, float4 *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float4 *ring = reinterpret_cast<float4 *>(smem) + (size_t)warp * STAGES * 32 * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * STAGES * 32 * 64) + warp * STAGES;
    const size_t per = deg / nw / 32 * 32;
    const float4 acc = warp_bulk_sum<STAGES>(table, ids, warp * per, warp * per + per, ring, bars, lane);
    if (acc.x == 12345.678f) out[warp * 32 + lane] = acc;
}

int main(int argc, char **argv) {
    const size_t n = 1 << 20;
    size_t m = 32u << 20;
    std::vector<unsigned> h(m);
    unsigned s = 12345;
    for (size_t i = 0; i < m; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % n; }
    if (argc > 1) {   // ids from a file (uint32): e.g. the col array of the R-MAT benchmark graph (dump_rmat_col.py)
        FILE *f = fopen(argv[1], "rb");
        if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
        m = fread(h.data(), 4, m, f) / 1024 * 1024;
        fclose(f);
        printf("ids from %s: %zu entries\n", argv[1], m);
    }
    float4 *table, *out;
    unsigned *ids;
    cudaMalloc(&table, n * 64); cudaMalloc(&ids, m * 4); cudaMalloc(&out, 64 << 20);
    cudaMemset(table, 0, n * 64);
    cudaMemcpy(ids, h.data(), m * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto smem_for = [](int warps, int stages) { return (size_t)warps * stages * (32 * 64 + 8); };
    auto run_all = [&](auto kern, int stages, int blocks_per_sm, int threads) {
        const size_t smem = smem_for(threads / 32, stages);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 3; ++rep) {
            if (rep == 1) cudaEventRecord(a);
            kern<<<148 * blocks_per_sm, threads, smem>>>(table, ids, m, out);
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 2;
        printf("A bulk stages=%d blocks/SM=%d threads=%d smem=%zu KB : %.3f ms  rows %.1f G/s  row bytes %.0f GB/s  (%s)\n",
               stages, blocks_per_sm, threads, smem / 1024, ms, m / ms / 1e6, m * 64.0 / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    };
    run_all(bulk_all<2>, 2, 2, 256); run_all(bulk_all<4>, 4, 2, 256); run_all(bulk_all<8>, 8, 2, 256);
    run_all(bulk_all<4>, 4, 4, 256); run_all(bulk_all<8>, 8, 1, 512);
    auto run_hub = [&](auto kern, int stages, size_t deg) {
        const size_t smem = smem_for(8, stages);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 3; ++rep) {
            if (rep == 1) cudaEventRecord(a);
            kern<<<1, 256, smem>>>(table, ids, deg, out);
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 2;
        printf("B hub of %zu rows on one CTA, stages=%d : %.3f ms  %.2f ns per neighbour  %.1f GB/s  (%s)\n", deg, stages, ms,
               ms * 1e6 / deg, deg * 64.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    run_hub(bulk_hub<2>, 2, 262144); run_hub(bulk_hub<4>, 4, 262144); run_hub(bulk_hub<8>, 8, 262144);
    run_hub(bulk_hub<16>, 16, 262144);
    return 0;
}
