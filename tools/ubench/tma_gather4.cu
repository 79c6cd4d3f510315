// Micro-benchmark (B200, sm_100a): gathering random ROWS of an L2-resident iterate tile into shared memory with the
// tensor-map TMA in tile::gather4 mode (one instruction = 4 arbitrary rows x `boxW` columns), against the same
// gather with LDG.128 into registers. Question (VERDICT r01, n1 / weak #2-iii): can TMA staging of the distribution
// tile feed the batched propagation faster than the L1 path? Round 1 only measured 1-D cp.async.bulk row copies
// (request-rate bound at 0.6-4 TB/s chip-wide).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_gather4 tools/ubench/tma_gather4.cu
//   ./tma_gather4 [boxH]      boxH = rows of the tensor map's box: the PTX ISA wants 1 for gather4; other values are tried
//                             only to document the encoding
// Every configuration is verified (the gathered rows are compared with the source) before it is timed. A TMA that never
// completes must not hang the device: the mbarrier wait gives up after ~1 s and the run reports "timeout".
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int kRows = 50000, kCols = 2048;  // one column tile of config 2: 410 MB of uint32

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, uint64_t *bar, int col, int r0, int r1,
                                            int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

// One CTA: `rounds` batches of `per_batch` gather4 operations (4 rows each), all issued by thread 0 and awaited by
// everybody on one mbarrier. rows: [gridDim.x][rounds][per_batch][4] row ids. verify: compare with the source.
template <int BOXW>
__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap map, const uint32_t *src,
                                              const int *rows, int rounds, int per_batch, int col_tiles, int verify,
                                              int *status, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t *buf = reinterpret_cast<uint32_t *>(smem);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long acc = 0;
    const int *my = rows + (size_t)blockIdx.x * rounds * per_batch * 4;
    for (int r = 0; r < rounds; ++r) {
        const int col = ((blockIdx.x + r) % col_tiles) * BOXW;
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (uint32_t)per_batch * 4 * BOXW * 4);
            for (int i = 0; i < per_batch; ++i) {
                const int *q = my + ((size_t)r * per_batch + i) * 4;
                tma_gather4(buf + (size_t)i * 4 * BOXW, &map, &bar, col, q[0], q[1], q[2], q[3]);
            }
        }
        const long long t0 = clock64();
        while (!mbar_try_wait(&bar, r & 1)) {
            if (clock64() - t0 > 200000000LL) {
                if (threadIdx.x == 0) atomicExch(status, 2);  // timeout
                return;
            }
        }
        if (verify) {
            for (int e = threadIdx.x; e < per_batch * 4 * BOXW; e += blockDim.x) {
                const int i = e / (4 * BOXW), rr = (e / BOXW) % 4, c = e % BOXW;
                const int row = my[((size_t)r * per_batch + i) * 4 + rr];
                if (buf[e] != src[(size_t)row * kCols + col + c]) atomicExch(status, 1);
            }
        } else {
            acc += buf[threadIdx.x % (per_batch * 4 * BOXW)];
        }
        __syncthreads();  // the buffer is reused by the next batch
    }
    if (acc == 0x1234567ull) sink[0] = acc;
}

// The same gather with LDG.128: a warp per row segment of BOXW uint32 (BOXW / 4 lanes x 16 bytes), several in flight.
template <int BOXW>
__global__ void __launch_bounds__(256) k_ldg(const uint32_t *src, const int *rows, int per_cta, int col_tiles,
                                              unsigned long long *sink) {
    constexpr int LANES = BOXW / 4;          // lanes per row segment
    constexpr int SEGS = 32 / LANES;         // row segments per warp-level load
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, seg = lane / LANES, l = lane % LANES;
    const int *my = rows + (size_t)blockIdx.x * per_cta;
    const int col = (blockIdx.x % col_tiles) * BOXW;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = warp * SEGS * 4; i + SEGS * 4 <= per_cta; i += 8 * SEGS * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = my[i + u * SEGS + seg];
            v[u] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)row * kCols + col) + l);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc.x ^= v[u].x, acc.y ^= v[u].y, acc.z ^= v[u].z, acc.w ^= v[u].w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = acc.x;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BOXW>
void run(EncodeFn encode, uint32_t *d_src, int boxH, int num_sms, int window_rows) {
    CUtensorMap map;
    cuuint64_t dims[2] = {kCols, kRows}, strides[1] = {(cuuint64_t)kCols * 4};
    cuuint32_t box[2] = {BOXW, (cuuint32_t)boxH}, estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        printf("boxW %4d boxH %d: cuTensorMapEncodeTiled failed (%d)\n", BOXW, boxH, (int)cr);
        return;
    }
    const int ctas_per_sm = 4, grid = (getenv("TMA_PROBE") ? 1 : num_sms * ctas_per_sm);
    const int batch_bytes = 48 * 1024, per_batch = batch_bytes / (4 * BOXW * 4), rounds = 64;
    const int col_tiles = kCols / BOXW;
    std::vector<int> rows((size_t)grid * rounds * per_batch * 4);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < rows.size(); ++i) {  // rows of a window of the tile (the L2-resident band of a wave of CTAs)
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        rows[i] = (int)(s % (uint64_t)window_rows);
    }
    int *d_rows, *d_status;
    unsigned long long *d_sink;
    CK(cudaMalloc(&d_rows, rows.size() * 4));
    CK(cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_status, 4));
    CK(cudaMemset(d_status, 0, 4));
    CK(cudaMalloc(&d_sink, 8));
    CK(cudaFuncSetAttribute(k_tma<BOXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, batch_bytes));
    k_tma<BOXW><<<grid, 128, batch_bytes>>>(map, d_src, d_rows, 2, per_batch, col_tiles, 1, d_status, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    int status = 0;
    if (e == cudaSuccess) CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
    if (e != cudaSuccess || status != 0) {
        printf("boxW %4d boxH %d: %s\n", BOXW, boxH,
               e != cudaSuccess ? cudaGetErrorString(e) : (status == 2 ? "timeout (the copies never completed)" : "WRONG DATA"));
        fflush(stdout);
        if (e != cudaSuccess) exit(3);  // sticky error: nothing else can run in this process
        return;
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms_tma = 0, ms_ldg = 0;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        k_tma<BOXW><<<grid, 128, batch_bytes>>>(map, d_src, d_rows, rounds, per_batch, col_tiles, 0, d_status, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_tma, e0, e1));
    }
    const double bytes = (double)grid * rounds * per_batch * 4 * BOXW * 4;
    const int per_cta = rounds * per_batch * 4;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        k_ldg<BOXW><<<grid, 256>>>(d_src, d_rows, per_cta, col_tiles, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_ldg, e0, e1));
    }
    printf("row segment %4d B (boxW %3d, boxH %d), rows from a window of %6d: TMA gather4 %.2f TB/s (%.0f M gather4/s, "
           "%.1f clk per instruction per SM) | LDG.128 %.2f TB/s\n",
           BOXW * 4, BOXW, boxH, window_rows, bytes / (ms_tma * 1e-3) / 1e12, grid * (double)rounds * per_batch / (ms_tma * 1e-3) / 1e6,
           1.9e9 * (ms_tma * 1e-3) / ((double)ctas_per_sm * rounds * per_batch), bytes / (ms_ldg * 1e-3) / 1e12);
    fflush(stdout);
    cudaFree(d_rows), cudaFree(d_status), cudaFree(d_sink);
}

int main(int argc, char **argv) {
    const int boxH = argc > 1 ? atoi(argv[1]) : 1;
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    uint32_t *d_src;
    CK(cudaMalloc(&d_src, (size_t)kRows * kCols * 4));
    std::vector<uint32_t> h((size_t)kRows * kCols);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)(i * 2654435761u);
    CK(cudaMemcpy(d_src, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    setvbuf(stdout, nullptr, _IONBF, 0);
    printf("init done\n");
    const int only = argc > 2 ? atoi(argv[2]) : 0;  // run one box width only (process isolation in the driver script)
    for (int window : {8192, kRows}) {  // 67 MB (L2-resident) and the whole 410 MB tile
        if (!only || only == 32) run<32>(encode, d_src, boxH, p.multiProcessorCount, window);
        if (!only || only == 64) run<64>(encode, d_src, boxH, p.multiProcessorCount, window);
        if (!only || only == 128) run<128>(encode, d_src, boxH, p.multiProcessorCount, window);
        if (only == 256) run<256>(encode, d_src, boxH, p.multiProcessorCount, window);
    }
    return 0;
}
