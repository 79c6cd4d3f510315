// K4 — pairwise distance matrices between sampled trajectories: dynamic time warping and
// Hausdorff distance over C series of `len` points in `d` dimensions, the O(C^2 len^2 d) step of the reference's
// lineage clustering (cytopath.py:30-45 compute_Hausdorff_distance, :90-96 dtaidistance.dtw_ndim.distance_matrix_fast).
//
// The arithmetic of the reference lives in two third-party packages that are not in this image (dtaidistance,
// hausdorff; unpinned in pyproject.toml:39-40); the kernels follow their published algorithms as restated in
// oracle/traj.py — float64 throughout, squared-Euclidean local cost summed in dimension order.
//
// One warp per pair of series (A = rows, B = columns), lane = one column of a 32-column band, the band's 32 points of
// B held in registers (d padded to DP, a template parameter). For every block of 32 rows:
//   phase 1  cost[r][lane] = sum_k (A[r][k] - B[c][k])^2 — A[r][k] is a warp-uniform (broadcast) load, so the phase
//            is float64-FMA bound; the 32 x 32 costs go to shared memory
//   phase 2  (DTW only) the recurrence D = cost + min(up, left, diagonal) as a skewed wavefront over the tile:
//            lane l handles row s - l at step s, takes `left` and `diagonal` from lane l-1 by shuffle; lane 0 reads the
//            previous band's last column from a per-warp buffer, lane 31 overwrites it for the next band.
#include <float.h>

#include "common.cuh"

namespace cy2p {

constexpr int kTrajWarps = 4;

__device__ __forceinline__ void pair_from_index(int64_t p, int C, int &i, int &j) {
    // p enumerates i < j row by row: row i starts at i*(2C-i-1)/2
    double fi = floor(((2.0 * C - 1.0) - sqrt((2.0 * C - 1.0) * (2.0 * C - 1.0) - 8.0 * (double)p)) * 0.5);
    int ii = (int)fi;
    if (ii < 0) ii = 0;
    while (ii > 0 && (int64_t)ii * (2 * C - ii - 1) / 2 > p) --ii;
    while ((int64_t)(ii + 1) * (2 * C - ii - 2) / 2 <= p) ++ii;
    i = ii;
    j = (int)(p - (int64_t)ii * (2 * C - ii - 1) / 2) + ii + 1;
}

template <int DP>
__device__ __forceinline__ double sq_dist(const double *__restrict__ arow, const double (&b)[DP]) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k += 2) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(arow + k));  // same address in every lane
        const double d0 = a.x - b[k], d1 = a.y - b[k + 1];
        acc = fma(d0, d0, acc);
        acc = fma(d1, d1, acc);
    }
    return acc;
}

// Costs of `nr` rows against the lane's column (coordinates b of the dimensions [K0, K0 + KN)): written to
// (ACC = false) or added to (ACC = true) the warp's shared-memory tile. Two rows at a time = two independent
// accumulation chains.
template <int DP, int K0, int KN, bool ACC>
__device__ __forceinline__ void cost_rows(const double *__restrict__ Arows, const double (&b)[KN], int nr,
                                          double (*t)[32], int lane) {
    int rr = 0;
    for (; rr + 1 < nr; rr += 2) {
        const double *a0 = Arows + (size_t)rr * DP + K0;
        double acc0 = ACC ? t[rr][lane] : 0.0, acc1 = ACC ? t[rr + 1][lane] : 0.0;
#pragma unroll
        for (int k = 0; k < KN; k += 2) {
            const double2 x0 = __ldg(reinterpret_cast<const double2 *>(a0 + k));
            const double2 x1 = __ldg(reinterpret_cast<const double2 *>(a0 + DP + k));
            const double d00 = x0.x - b[k], d01 = x0.y - b[k + 1];
            const double d10 = x1.x - b[k], d11 = x1.y - b[k + 1];
            acc0 = fma(d00, d00, acc0);
            acc1 = fma(d10, d10, acc1);
            acc0 = fma(d01, d01, acc0);
            acc1 = fma(d11, d11, acc1);
        }
        t[rr][lane] = acc0;
        t[rr + 1][lane] = acc1;
    }
    if (rr < nr) {
        const double *a0 = Arows + (size_t)rr * DP + K0;
        double acc0 = ACC ? t[rr][lane] : 0.0;
#pragma unroll
        for (int k = 0; k < KN; k += 2) {
            const double2 x0 = __ldg(reinterpret_cast<const double2 *>(a0 + k));
            const double d00 = x0.x - b[k], d01 = x0.y - b[k + 1];
            acc0 = fma(d00, d00, acc0);
            acc0 = fma(d01, d01, acc0);
        }
        t[rr][lane] = acc0;
    }
}

template <int DP>
__global__ void __launch_bounds__(kTrajWarps * 32)
dtw_pairs_kernel(int C, int len, const double *__restrict__ S, double *__restrict__ out, double *__restrict__ left_ws) {
    __shared__ double tile[kTrajWarps][32][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t npairs = (int64_t)C * (C - 1) / 2;
    const int64_t wglobal = (int64_t)blockIdx.x * kTrajWarps + warp, wstride = (int64_t)gridDim.x * kTrajWarps;
    double *left = left_ws + wglobal * len;
    double(*t)[32] = tile[warp];
    for (int64_t p = wglobal; p < npairs; p += wstride) {
        int i, j;
        pair_from_index(p, C, i, j);
        const double *A = S + (size_t)i * len * DP, *B = S + (size_t)j * len * DP;
        double result = 0.0;
        for (int c0 = 0; c0 < len; c0 += 32) {
            const int c = c0 + lane;
            const bool colok = c < len;
            const double *Bcol = B + (size_t)(colok ? c : 0) * DP;
            double bband[DP <= 32 ? DP : 2];  // the whole column when it fits the register budget
            if constexpr (DP <= 32) {
#pragma unroll
                for (int k = 0; k < DP; ++k) bband[k] = colok ? Bcol[k] : 0.0;
            }
            double v1 = DBL_MAX, v2 = DBL_MAX;       // D[r][c] of the lane's latest row and of the row before it
            double lf_prev = c0 == 0 ? 0.0 : DBL_MAX;  // lane 0: D[r-1][c0-1]; D[-1][-1] = 0 starts the recurrence
            for (int r0 = 0; r0 < len; r0 += 32) {
                const int nr = min(32, len - r0);
                // the previous band's last column for these 32 rows: one coalesced load, handed to lane 0 by shuffle
                const double left_in = (c0 > 0 && lane < nr) ? __ldcg(left + r0 + lane) : DBL_MAX;
                if constexpr (DP <= 32) {
                    cost_rows<DP, 0, DP, false>(A + (size_t)r0 * DP, bband, nr, t, lane);
                } else {  // two passes over the dimensions: at most 32 coordinates of the column live in registers
                    {
                        double b[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) b[k] = colok ? Bcol[k] : 0.0;
                        cost_rows<DP, 0, 32, false>(A + (size_t)r0 * DP, b, nr, t, lane);
                    }
                    {
                        double b[DP - 32];
#pragma unroll
                        for (int k = 0; k < DP - 32; ++k) b[k] = colok ? Bcol[32 + k] : 0.0;
                        cost_rows<DP, 32, DP - 32, true>(A + (size_t)r0 * DP, b, nr, t, lane);
                    }
                }
                __syncwarp();
                double left_out = 0.0;  // lane q collects D[r0 + q][c0 + 31] for the next band
                for (int s = 0; s < nr + 31; ++s) {
                    const int q = s - lane;
                    const bool active = colok && q >= 0 && q < nr;
                    double lf = __shfl_up_sync(0xffffffffu, v1, 1);
                    double dg = __shfl_up_sync(0xffffffffu, v2, 1);
                    const double l0 = __shfl_sync(0xffffffffu, left_in, s & 31);
                    if (lane == 0) {
                        dg = lf_prev;
                        lf = l0;
                    }
                    if (active) {
                        const double best = fmin(fmin(v1, lf), dg);  // v1 is D[r-1][c]
                        v2 = v1;
                        v1 = t[q][lane] + best;
                        if (lane == 0) lf_prev = lf;
                    }
                    const double last = __shfl_sync(0xffffffffu, v1, 31);  // lane 31 has just finished row s - 31
                    if (lane == s - 31) left_out = last;
                }
                if (lane < nr && c0 + 31 < len) __stcg(left + r0 + lane, left_out);
                __syncwarp();
            }
            if (c == len - 1) result = v1;
            __syncwarp();  // the band's last column is complete in `left` before lane 0 of the next band reads it
        }
        result = __shfl_sync(0xffffffffu, result, (len - 1) & 31);
        if (lane == 0) {
            const double dist = sqrt(result);
            out[(size_t)i * C + j] = dist;
            out[(size_t)j * C + i] = dist;
        }
    }
}

// g[i][j] = max over the points b of series j of min over the points a of series i of |a - b|   (i != j)
template <int DP>
__global__ void __launch_bounds__(kTrajWarps * 32)
hausdorff_directed_kernel(int C, int len, const double *__restrict__ S, double *__restrict__ g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t npairs = (int64_t)C * C;
    const int64_t wglobal = (int64_t)blockIdx.x * kTrajWarps + warp, wstride = (int64_t)gridDim.x * kTrajWarps;
    for (int64_t p = wglobal; p < npairs; p += wstride) {
        const int i = (int)(p / C), j = (int)(p % C);
        if (i == j) {
            if (lane == 0) g[p] = 0.0;
            continue;
        }
        const double *A = S + (size_t)i * len * DP, *B = S + (size_t)j * len * DP;
        double worst = 0.0;
        for (int c0 = 0; c0 < len; c0 += 32) {
            const int c = c0 + lane;
            const bool colok = c < len;
            double b[DP];
#pragma unroll
            for (int k = 0; k < DP; ++k) b[k] = colok ? B[(size_t)c * DP + k] : 0.0;
            double colmin = DBL_MAX;
            for (int r = 0; r < len; ++r) colmin = fmin(colmin, sq_dist<DP>(A + (size_t)r * DP, b));
            if (colok) worst = fmax(worst, colmin);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        if (lane == 0) g[p] = sqrt(worst);
    }
}

__global__ void hausdorff_symmetrise_kernel(int C, const double *__restrict__ g, double *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)C * C) return;
    const int i = (int)(p / C), j = (int)(p % C);
    out[p] = fmax(g[p], g[(size_t)j * C + i]);
}

__global__ void zero_diagonal_kernel(int C, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) out[(size_t)i * C + i] = 0.0;
}

static int traj_grid(const void *kernel, int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTrajWarps * 32, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return sms * per_sm;
}

constexpr int kTrajMaxCtas = 148 * 16;  // bound of the persistent grid (sizes the per-warp work space)

}  // namespace cy2p

using namespace cy2p;

extern "C" {

size_t cy2p_traj_ws_bytes(int32_t C, int32_t len) {
    if (C < 1 || len < 1) return 0;
    const size_t dtw = sizeof(double) * (size_t)kTrajMaxCtas * kTrajWarps * (size_t)len;
    const size_t hd = sizeof(double) * (size_t)C * (size_t)C;
    return (dtw > hd ? dtw : hd) + 256;
}

#define CY2P_TRAJ_DISPATCH(KERNEL, ...)                                                   \
    switch (dpad) {                                                                       \
        case 8: KERNEL<8> __VA_ARGS__; break;                                             \
        case 16: KERNEL<16> __VA_ARGS__; break;                                           \
        case 24: KERNEL<24> __VA_ARGS__; break;                                           \
        case 32: KERNEL<32> __VA_ARGS__; break;                                           \
        case 40: KERNEL<40> __VA_ARGS__; break;                                           \
        case 48: KERNEL<48> __VA_ARGS__; break;                                           \
        case 56: KERNEL<56> __VA_ARGS__; break;                                           \
        default: KERNEL<64> __VA_ARGS__; break;                                           \
    }

static int32_t traj_check(const double *d_series, int32_t C, int32_t len, int32_t dpad, double *d_out, void *d_ws,
                          size_t ws_bytes) {
    CY2P_REQUIRE(d_series && d_out, "null pointer");
    CY2P_REQUIRE(C >= 1 && len >= 1, "C >= 1 and len >= 1 required");
    CY2P_REQUIRE(dpad >= 8 && dpad <= 64 && dpad % 8 == 0, "point dimension must be padded to a multiple of 8, at most 64");
    CY2P_REQUIRE(((uintptr_t)d_series & 15) == 0, "series must be 16-byte aligned");
    if (!d_ws || ws_bytes < cy2p_traj_ws_bytes(C, len)) {
        set_error("work space too small: need %zu bytes", cy2p_traj_ws_bytes(C, len));
        return CY2P_ERR_WORKSPACE;
    }
    return CY2P_OK;
}

int32_t cy2p_dtw_matrix(const double *d_series, int32_t C, int32_t len, int32_t dpad, double *d_out, void *d_ws,
                        size_t ws_bytes, void *stream) {
    int32_t rc = traj_check(d_series, C, len, dpad, d_out, d_ws, ws_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    CY2P_CUDA(cudaGetDevice(&device));
    count_launch();
    zero_diagonal_kernel<<<(C + 255) / 256, 256, 0, st>>>(C, d_out);
    if (C > 1) {
        const int64_t npairs = (int64_t)C * (C - 1) / 2;
        int grid = 0;
#define CY2P_GRID(DPV) grid = traj_grid((const void *)dtw_pairs_kernel<DPV>, device)
        switch (dpad) {
            case 8: CY2P_GRID(8); break;
            case 16: CY2P_GRID(16); break;
            case 24: CY2P_GRID(24); break;
            case 32: CY2P_GRID(32); break;
            case 40: CY2P_GRID(40); break;
            case 48: CY2P_GRID(48); break;
            case 56: CY2P_GRID(56); break;
            default: CY2P_GRID(64); break;
        }
#undef CY2P_GRID
        const int64_t need = (npairs + kTrajWarps - 1) / kTrajWarps;
        if (grid > need) grid = (int)need;
        if (grid > kTrajMaxCtas) grid = kTrajMaxCtas;
        count_launch();
        CY2P_TRAJ_DISPATCH(dtw_pairs_kernel, <<<grid, kTrajWarps * 32, 0, st>>>(C, len, d_series, d_out, (double *)d_ws));
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_hausdorff_matrix(const double *d_series, int32_t C, int32_t len, int32_t dpad, double *d_out, void *d_ws,
                              size_t ws_bytes, void *stream) {
    int32_t rc = traj_check(d_series, C, len, dpad, d_out, d_ws, ws_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t npairs = (int64_t)C * C;
    int64_t grid = (npairs + kTrajWarps - 1) / kTrajWarps;
    if (grid > kTrajMaxCtas) grid = kTrajMaxCtas;
    count_launch();
    CY2P_TRAJ_DISPATCH(hausdorff_directed_kernel, <<<(int)grid, kTrajWarps * 32, 0, st>>>(C, len, d_series, (double *)d_ws));
    count_launch();
    hausdorff_symmetrise_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(C, (const double *)d_ws, d_out);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"
