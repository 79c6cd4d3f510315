// Micro-benchmark (tuning aid): cost of cluster.sync() per iteration for a 16-CTA cluster, with and without
// global / distributed-shared stores in between.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void k(int iters, float *g, long long *out) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float sm[];
    const int rank = cluster.block_rank(), nc = cluster.num_blocks();
    float v = threadIdx.x;
    cluster.sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) g[(it & 1) * 65536 + rank * blockDim.x + threadIdx.x] = v;  // global store
        if (MODE == 2) {                                                            // one DSMEM store per thread
            cluster.map_shared_rank(sm, (rank + 1 + (threadIdx.x & 7)) % nc)[threadIdx.x] = v;
        }
        if (MODE == 3 && threadIdx.x < 256) {                                        // 16 DSMEM stores from 256 threads
            for (int d = 0; d < nc; ++d) cluster.map_shared_rank(sm, d)[threadIdx.x + 1024] = v;
        }
        cluster.sync();
        v += sm[threadIdx.x];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && rank == 0) out[0] = t1 - t0;
    if (v == 1234.5f) g[0] = v;
}

template <int MODE>
void run(int c, int threads, const char *name) {
    float *g;
    long long *out;
    cudaMalloc(&g, 4 * 131072 * 4);
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = 32768;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, k<MODE>, iters, g, out);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-28s cluster=%2d threads=%4d: %7.1f cycles / iteration  (%s)\n", name, c, threads, (double)h / iters,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    for (int c : {8, 16})
        for (int th : {256, 1024}) {
            run<0>(c, th, "sync only");
            run<1>(c, th, "global store + sync");
            run<2>(c, th, "1 dsmem store/thread + sync");
            run<3>(c, th, "16 dsmem stores x256 + sync");
        }
    return 0;
}
