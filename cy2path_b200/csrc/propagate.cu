// K1 — MSM state-probability propagation  X <- X . T  (reference sampling.py:38-45).
//
// Formulation: y[j, :] = sum_i T[i, j] * x[i, :] is a row-gather over T^T (destination-major CSR,
// built once by plan.cu), so every destination row is owned by one warp and written exactly once
// (no atomics on the iterate). X is cell-major [n][B]: one source row is B contiguous floats, so a
// gather of row i is a coalesced 128-bit-per-lane load.
//
// Kernels in this file
//   spmm_gather<VEC>      general B: warp = one destination row x (32*VEC) columns, (src,val) pairs
//                         fetched two at a time with warp-uniform 128-bit loads, fused fp64 epilogue
//                         for the cosine distance (x.pi and x.x partial sums).
//   spmv_b1_persistent    B == 1 (the reference's own case): cooperative persistent kernel, all
//                         iterations in one launch with a grid barrier per update; warp per row,
//                         lanes across the row's entries, shuffle reduction; writes the full
//                         [iters][n] history in place (iterate i+1 is produced directly in row i+1).
//   eps_finalize          eps = clip(1 - uv / sqrt(uu * vv), 0, 2)  (scipy cosine, sampling.py:44)
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cy2p {

constexpr int kWarps = 8;         // warps per CTA of the gather kernel
constexpr int kRowsPerCta = 64;   // destination rows per CTA (8 per warp): amortises the eps reduction

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T load(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void store(float *p, T v) { *reinterpret_cast<float4 *>(p) = v; }
    static __device__ __forceinline__ void fma(T &a, float t, T x) {
        a.x = fmaf(t, x.x, a.x);
        a.y = fmaf(t, x.y, a.y);
        a.z = fmaf(t, x.z, a.z);
        a.w = fmaf(t, x.w, a.w);
    }
    static __device__ __forceinline__ float get(const T &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
};
template <>
struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float *p, T v) { *p = v; }
    static __device__ __forceinline__ void fma(T &a, float t, T x) { a = fmaf(t, x, a); }
    static __device__ __forceinline__ float get(const T &v, int) { return v; }
};

// One update for destination rows [row_begin,row_end) and `ncols` columns starting at the column the
// base pointers address. eps_acc (may be null): [ncols][2] doubles, (x.pi, x.x) sums for this update.
template <int VEC, bool EPS>
__global__ void __launch_bounds__(kWarps * 32)
spmm_gather(int32_t row_begin, int32_t row_end, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs,
            const float *__restrict__ Xin, int64_t ld_in, float *__restrict__ Xout, int64_t ld_out, int32_t ncols,
            const double *__restrict__ stationary, double *__restrict__ eps_acc) {
    using V = Vec<VEC>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (blockIdx.y * 32 + lane) * VEC;
    const bool active = col < ncols;
    const int32_t row0 = row_begin + blockIdx.x * kRowsPerCta;
    const int32_t row1 = min(row0 + kRowsPerCta, row_end);
    double uv[VEC], uu[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) uv[k] = uu[k] = 0.0;

    if (active) {
        const float *xin = Xin + col;
        for (int32_t r = row0 + warp; r < row1; r += kWarps) {
            int32_t p = t_indptr[r];
            const int32_t end = t_indptr[r + 1];
            typename V::T acc = V::zero();
            if (p < end && (p & 1)) {  // align the entry cursor to 16 bytes
                const Pair e = pairs[p];
                V::fma(acc, e.val, V::load(xin + (int64_t)e.src * ld_in));
                ++p;
            }
#pragma unroll 2
            for (; p + 1 < end; p += 2) {  // two (src,val) entries per warp-uniform 128-bit load
                const int4 q = __ldg(reinterpret_cast<const int4 *>(pairs + p));
                const typename V::T x0 = V::load(xin + (int64_t)q.x * ld_in);
                const typename V::T x1 = V::load(xin + (int64_t)q.z * ld_in);
                V::fma(acc, __int_as_float(q.y), x0);
                V::fma(acc, __int_as_float(q.w), x1);
            }
            if (p < end) {
                const Pair e = pairs[p];
                V::fma(acc, e.val, V::load(xin + (int64_t)e.src * ld_in));
            }
            V::store(Xout + (int64_t)r * ld_out + col, acc);
            if (EPS) {
                const double pi = stationary[r];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const double y = (double)V::get(acc, k);
                    uv[k] = fma(y, pi, uv[k]);
                    uu[k] = fma(y, y, uu[k]);
                }
            }
        }
    }
    if (EPS) {
        __shared__ double red[kWarps][32 * VEC][2];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            red[warp][lane * VEC + k][0] = uv[k];
            red[warp][lane * VEC + k][1] = uu[k];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 32 * VEC * 2; t += kWarps * 32) {
            const int c = t >> 1, which = t & 1;
            const int gcol = blockIdx.y * 32 * VEC + c;
            if (gcol < ncols) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) s += red[w][c][which];
                atomicAdd(&eps_acc[(int64_t)gcol * 2 + which], s);
            }
        }
    }
}

// B == 1, all iterations in one cooperative launch.
//   hist != null: row it of hist is the source of update it, row it+1 (or `last` for the final
//                 update) its destination; hist[0] must already hold x0.
//   hist == null: x0 -> bufA -> bufB -> bufA ...; the final update writes `last`.
// eps_acc: [iters][2] doubles (zeroed), or null.
__global__ void __launch_bounds__(256)
spmv_b1_persistent(int32_t n, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs, const float *x0,
                   float *hist, float *bufA, float *bufB, float *last, int32_t iters,
                   const double *__restrict__ stationary, double *eps_acc) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    __shared__ double red[8][2];

    for (int32_t it = 0; it < iters; ++it) {
        const float *src;
        float *dst;
        if (hist) {
            src = hist + (int64_t)it * n;
            dst = (it + 1 < iters) ? hist + (int64_t)(it + 1) * n : last;
        } else {
            src = it == 0 ? x0 : ((it & 1) ? bufA : bufB);
            dst = (it + 1 == iters) ? last : ((it & 1) ? bufB : bufA);
        }
        double uv = 0.0, uu = 0.0;
        for (int32_t r = gwarp; r < n; r += nwarps) {
            const int32_t beg = t_indptr[r], end = t_indptr[r + 1];
            float acc = 0.f;
            for (int32_t p = beg + lane; p < end; p += 32) {
                const Pair e = pairs[p];
                acc = fmaf(e.val, __ldcg(src + e.src), acc);  // L2 load: src was written earlier in this launch
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                dst[r] = acc;
                if (eps_acc) {
                    const double y = (double)acc;
                    uv = fma(y, stationary[r], uv);
                    uu = fma(y, y, uu);
                }
            }
        }
        if (eps_acc) {
            if (lane == 0) {
                red[threadIdx.x >> 5][0] = uv;
                red[threadIdx.x >> 5][1] = uu;
            }
            __syncthreads();
            if (threadIdx.x < 2) {
                double s = 0.0;
                for (int w = 0; w < warps_per_cta; ++w) s += red[w][threadIdx.x];
                atomicAdd(&eps_acc[(int64_t)it * 2 + threadIdx.x], s);
            }
            __syncthreads();
        }
        grid.sync();
    }
}

__global__ void sumsq_kernel(int32_t n, const double *__restrict__ v, double *out) {
    double s = 0.0;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s = fma(v[i], v[i], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// eps[i] = clip(1 - uv/sqrt(uu*vv), 0, 2): scipy.spatial.distance.cosine as called at sampling.py:44
__global__ void eps_finalize(int64_t count, const double *__restrict__ acc, const double *__restrict__ vv,
                             double *__restrict__ eps) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double d = 1.0 - acc[2 * i] / sqrt(acc[2 * i + 1] * vv[0]);
    eps[i] = fmin(fmax(d, 0.0), 2.0);
}

// acc is [iters][B][2] in (iteration, column) order already; used when a column tile was accumulated
// into a [iters][tile][2] scratch: scatter into eps[iters][B].
__global__ void eps_finalize_tile(int32_t iters, int32_t tile_cols, int32_t col0, int32_t B,
                                  const double *__restrict__ acc, const double *__restrict__ vv,
                                  double *__restrict__ eps) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)iters * tile_cols) return;
    const int32_t it = (int32_t)(i / tile_cols), c = (int32_t)(i % tile_cols);
    const double d = 1.0 - acc[2 * i] / sqrt(acc[2 * i + 1] * vv[0]);
    eps[(int64_t)it * B + col0 + c] = fmin(fmax(d, 0.0), 2.0);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Column-tile width: two ping-pong tiles [n][bt] should stay L2-resident across the iterations.
static int32_t choose_tile(const cy2p_plan *plan, int32_t B) {
    const size_t budget = plan->l2_bytes ? plan->l2_bytes * 2 / 5 : (size_t)48 << 20;  // ~40% of L2
    int64_t bt = (int64_t)(budget / ((size_t)plan->n * 8));
    bt = bt / 128 * 128;
    if (bt < 128) bt = 128;
    if (bt > B) bt = B;
    return (int32_t)bt;
}

struct WsLayout {
    size_t eps_acc, vv, bufA, bufB, total;
    int32_t tile;
};

static WsLayout ws_layout(const cy2p_plan *plan, int32_t B, int32_t iters) {
    WsLayout w;
    w.tile = choose_tile(plan, B);
    size_t off = 0;
    w.eps_acc = off;
    off = align_up(off + sizeof(double) * 2 * (size_t)iters * (size_t)w.tile, 256);
    w.vv = off;
    off = align_up(off + sizeof(double), 256);
    w.bufA = off;
    off = align_up(off + sizeof(float) * (size_t)plan->n * (size_t)w.tile, 256);
    w.bufB = off;
    off = align_up(off + sizeof(float) * (size_t)plan->n * (size_t)w.tile, 256);
    w.total = off;
    return w;
}

template <int VEC>
static void launch_gather(const cy2p_plan *plan, int32_t row_begin, int32_t row_end, const float *in, int64_t ld_in,
                          float *out, int64_t ld_out, int32_t ncols, const double *stationary, double *eps_acc,
                          cudaStream_t st) {
    dim3 grid((unsigned)((row_end - row_begin + kRowsPerCta - 1) / kRowsPerCta),
              (unsigned)((ncols + 32 * VEC - 1) / (32 * VEC)));
    count_launch();
    if (eps_acc)
        spmm_gather<VEC, true><<<grid, kWarps * 32, 0, st>>>(row_begin, row_end, plan->d_t_indptr, plan->d_t_pairs, in,
                                                            ld_in, out, ld_out, ncols, stationary, eps_acc);
    else
        spmm_gather<VEC, false><<<grid, kWarps * 32, 0, st>>>(row_begin, row_end, plan->d_t_indptr, plan->d_t_pairs,
                                                             in, ld_in, out, ld_out, ncols, nullptr, nullptr);
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace cy2p

using namespace cy2p;

extern "C" {

size_t cy2p_propagate_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters) {
    if (!plan || B < 1 || iters < 0) return 0;
    return ws_layout(plan, B, iters).total;
}

int32_t cy2p_propagate_column_tile(const cy2p_plan *plan, int32_t B) {
    if (!plan || B < 1) return 0;
    return choose_tile(plan, B);
}

int32_t cy2p_propagate(cy2p_plan *plan, const float *d_X0, int32_t B, int32_t iters, const double *d_stationary,
                       float *d_Xfinal, float *d_history, int32_t history_stride, double *d_eps, void *d_ws,
                       size_t ws_bytes, void *stream) {
    CY2P_REQUIRE(plan && d_X0, "null plan / X0");
    CY2P_REQUIRE(B >= 1 && iters >= 0, "B >= 1 and iters >= 0 required");
    CY2P_REQUIRE(!d_history || history_stride >= 1, "history_stride must be >= 1");
    CY2P_REQUIRE(!d_eps || d_stationary, "eps requested without a stationary vector");
    CY2P_REQUIRE(d_Xfinal || d_history, "nothing to write: both Xfinal and history are null");
    const WsLayout w = ws_layout(plan, B, iters);
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    const int32_t n = plan->n;
    char *ws = (char *)d_ws;
    double *eps_acc = (double *)(ws + w.eps_acc);
    double *vv = (double *)(ws + w.vv);
    float *bufA = (float *)(ws + w.bufA), *bufB = (float *)(ws + w.bufB);

    if (iters == 0) {
        if (d_Xfinal && d_Xfinal != d_X0)
            CY2P_CUDA(cudaMemcpyAsync(d_Xfinal, d_X0, sizeof(float) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));
        return CY2P_OK;
    }
    if (d_eps) {
        CY2P_CUDA(cudaMemsetAsync(vv, 0, sizeof(double), st));
        count_launch(); sumsq_kernel<<<64, 256, 0, st>>>(n, d_stationary, vv);
    }
    if (d_history)  // row 0 of the history is the start distribution (sampling.py:39)
        CY2P_CUDA(cudaMemcpyAsync(d_history, d_X0, sizeof(float) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));

    // ---- B == 1: one cooperative launch for all iterations --------------------------------------
    if (B == 1 && (!d_history || history_stride == 1)) {
        if (d_eps) CY2P_CUDA(cudaMemsetAsync(eps_acc, 0, sizeof(double) * 2 * (size_t)iters, st));
        int per_sm = 0;
        CY2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_b1_persistent, 256, 0));
        int grid = plan->num_sms * (per_sm > 2 ? 2 : per_sm);
        const int need = (n + 7) / 8;
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        float *last = d_Xfinal ? d_Xfinal : bufB;
        double *acc = d_eps ? eps_acc : nullptr;
        int32_t n_ = n, iters_ = iters;
        const int32_t *tp = plan->d_t_indptr;
        const Pair *pp = plan->d_t_pairs;
        void *args[] = {&n_, &tp, &pp, (void *)&d_X0, &d_history, &bufA, &bufB, &last, &iters_, (void *)&d_stationary, &acc};
        count_launch();
        CY2P_CUDA(cudaLaunchCooperativeKernel((void *)spmv_b1_persistent, dim3(grid), dim3(256), args, 0, st));
        if (d_eps) {
            count_launch();
            eps_finalize<<<(iters + 255) / 256, 256, 0, st>>>((int64_t)iters, eps_acc, vv, d_eps);
        }
        CY2P_CUDA(cudaGetLastError());
        return CY2P_OK;
    }

    // ---- general B: column tiles, each carried through all iterations while it is L2-resident -----
    const bool vec4 = (B % 4 == 0) && aligned16(d_X0) && (!d_Xfinal || aligned16(d_Xfinal)) &&
                      (!d_history || aligned16(d_history));
    const int64_t slots = d_history ? ((int64_t)iters + history_stride - 1) / history_stride : 0;
    for (int32_t c0 = 0; c0 < B; c0 += w.tile) {
        const int32_t nc = (B - c0 < w.tile) ? B - c0 : w.tile;
        if (d_eps) CY2P_CUDA(cudaMemsetAsync(eps_acc, 0, sizeof(double) * 2 * (size_t)iters * (size_t)nc, st));
        const float *src = d_history ? d_history + c0 : d_X0 + c0;  // history row 0 holds x0
        int64_t ld_src = B;
        for (int32_t it = 0; it < iters; ++it) {
            float *dst;
            int64_t ld_dst;
            const int64_t nxt = (int64_t)it + 1;
            if (d_history && nxt % history_stride == 0 && nxt / history_stride < slots) {
                dst = d_history + (size_t)(nxt / history_stride) * (size_t)n * B + c0;
                ld_dst = B;
            } else if (nxt == iters && d_Xfinal) {
                dst = d_Xfinal + c0;
                ld_dst = B;
            } else {
                dst = (src == bufA) ? bufB : bufA;
                ld_dst = nc;
            }
            double *acc = d_eps ? eps_acc + 2 * (size_t)it * (size_t)nc : nullptr;
            const bool v4 = vec4 && (nc % 4 == 0);
            if (v4)
                launch_gather<4>(plan, 0, n, src, ld_src, dst, ld_dst, nc, d_stationary, acc, st);
            else
                launch_gather<1>(plan, 0, n, src, ld_src, dst, ld_dst, nc, d_stationary, acc, st);
            src = dst;
            ld_src = ld_dst;
        }
        if (d_Xfinal && src != d_Xfinal + c0)  // the last iterate landed in a history slot / scratch
            CY2P_CUDA(cudaMemcpy2DAsync(d_Xfinal + c0, sizeof(float) * (size_t)B, src, sizeof(float) * (size_t)ld_src,
                                        sizeof(float) * (size_t)nc, (size_t)n, cudaMemcpyDeviceToDevice, st));
        if (d_eps) {
            const int64_t cnt = (int64_t)iters * nc;
            count_launch(); eps_finalize_tile<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(iters, nc, c0, B, eps_acc, vv, d_eps);
        }
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_propagate_rows(cy2p_plan *plan, const float *d_X, float *d_Y, int32_t B, int32_t row_begin,
                            int32_t row_end, void *stream) {
    CY2P_REQUIRE(plan && d_X && d_Y, "null pointer");
    CY2P_REQUIRE(B >= 1 && row_begin >= 0 && row_end <= plan->n && row_begin <= row_end, "bad row range / B");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    if (row_begin == row_end) return CY2P_OK;
    if (B % 4 == 0 && aligned16(d_X) && aligned16(d_Y))
        launch_gather<4>(plan, row_begin, row_end, d_X, B, d_Y, B, B, nullptr, nullptr, st);
    else
        launch_gather<1>(plan, row_begin, row_end, d_X, B, d_Y, B, B, nullptr, nullptr, st);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"
