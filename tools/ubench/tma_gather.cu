// Micro-benchmark (tuning aid, not product): throughput of many small cp.async.bulk (TMA 1-D) row gathers
// from an L2-resident buffer into shared memory, vs the same gather done with LDG.128 into registers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather tma_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Each CTA: `rounds` times { issue `rows` bulk copies of `rowbytes` each (row ids from idx), wait }.
// Issue is spread over the 32 lanes of warp 0.
__global__ void tma_gather_kernel(const char *X, const int *idx, int nidx, int rows, int rowbytes, size_t ld_bytes,
                                  int rounds, float *sink) {
    extern __shared__ __align__(128) char stage[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        if (threadIdx.x == 0) mbar_expect_tx(&bar, (uint32_t)rows * rowbytes);
        __syncthreads();
        if (threadIdx.x < 32) {
            const int base = (blockIdx.x * rounds + r) * rows;
            for (int i = threadIdx.x; i < rows; i += 32) {
                const int row = idx[(base + i) % nidx];
                bulk_g2s(stage + (size_t)i * rowbytes, X + (size_t)row * ld_bytes, rowbytes, &bar);
            }
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        acc += ((float *)stage)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 123.456f) sink[0] = acc;
}

// Reference: same rows gathered with LDG.128 by all threads (each warp reads one row segment of 512 B per step).
__global__ void ldg_gather_kernel(const char *X, const int *idx, int nidx, int rows, int rowbytes, size_t ld_bytes,
                                  int rounds, float *sink) {
    float4 acc = make_float4(0, 0, 0, 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int per_row = rowbytes / 16;  // lanes needed per row
    for (int r = 0; r < rounds; ++r) {
        const int base = (blockIdx.x * rounds + r) * rows;
        for (int i = warp * (32 / per_row) + lane / per_row; i < rows; i += nw * (32 / per_row)) {
            const int row = idx[(base + i) % nidx];
            const float4 v = __ldg((const float4 *)(X + (size_t)row * ld_bytes) + (lane % per_row));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

int main() {
    const int n = 50000;
    const size_t ld_bytes = 2048 * 4;  // a 2048-column fp32 tile: rows 8 KB apart, we fetch `rowbytes` of each
    char *X;
    cudaMalloc(&X, (size_t)n * ld_bytes);
    cudaMemset(X, 0, (size_t)n * ld_bytes);
    const int nidx = 1 << 22;
    int *h = (int *)malloc(sizeof(int) * nidx);
    srand(1);
    for (int i = 0; i < nidx; ++i) h[i] = rand() % n;
    int *idx;
    cudaMalloc(&idx, sizeof(int) * nidx);
    cudaMemcpy(idx, h, sizeof(int) * nidx, cudaMemcpyHostToDevice);
    float *sink;
    cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(tma_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int rowbytes_opts[] = {128, 256, 512};
    for (int rb : rowbytes_opts) {
        for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
            const int stage_bytes = ctas_per_sm == 1 ? 192 * 1024 : 96 * 1024;
            const int rows = stage_bytes / rb;
            const int rounds = 64;
            const int grid = 148 * ctas_per_sm;
            for (int mode = 0; mode < 2; ++mode) {
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0)
                        tma_gather_kernel<<<grid, 256, stage_bytes>>>(X, idx, nidx, rows, rb, ld_bytes, rounds, sink);
                    else
                        ldg_gather_kernel<<<grid * 4, 256>>>(X, idx, nidx, rows, rb, ld_bytes, rounds / 4, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)grid * rounds * rows * rb;
                printf("%s rowbytes=%d ctas/sm=%d rows/stage=%d: %.3f ms  %.2f TB/s  (%.1f B/clk/SM @1.92GHz)  err=%s\n",
                       mode == 0 ? "TMA-bulk" : "LDG.128 ", rb, ctas_per_sm, rows, ms, bytes / ms / 1e9,
                       bytes / (ms * 1e-3) / 148 / 1.92e9, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
