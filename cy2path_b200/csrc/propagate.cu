// K1 — MSM state-probability propagation  X <- X . T  (reference sampling.py:38-45).
//
// Formulation: y[j, :] = sum_i T[i, j] * x[i, :] is a row-gather over T^T (destination-major CSR,
// built once by plan.cu), so every destination row is owned by one warp and written exactly once
// (no atomics on the iterate). X is cell-major [n][B]: one source row is B contiguous values, so a
// gather of row i is a coalesced 128-bit-per-lane load.
//
// Kernels in this file (templated on the iterate type: float for the batched fp32 path the
// north_star names, double for the reference's own float64 single-start case and for long batched
// runs whose parity must hold over thousands of iterations — see DESIGN.md "precision")
//   spmm_gather<T,VEC>    general B: warp = one destination row x (32*VEC) columns, (src,val) pairs
//                         fetched two at a time with warp-uniform 128-bit loads, four row gathers in
//                         flight per warp, fused fp64 epilogue for the cosine distance (x.pi, x.x).
//   spmv_b1_persistent<T> B == 1 (the reference's own case): cooperative persistent kernel, all
//                         iterations in one launch with a grid barrier per update; warp per row,
//                         lanes across the row's entries, shuffle reduction; writes the full
//                         [iters][n] history in place (iterate i+1 is produced directly in row i+1).
//   eps_finalize_tile     eps = clip(1 - uv / sqrt(uu * vv), 0, 2)  (scipy cosine, sampling.py:44)
#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include <type_traits>

namespace cg = cooperative_groups;

namespace cy2p {

constexpr int kWarps = 8;       // warps per CTA of the gather kernel
constexpr int kEpsCopies = 4;    // eps partial sums are spread over this many copies (same-address atomics serialise in L2)

template <typename T, int VEC>
struct Vec;
template <>
struct Vec<float, 4> {
    using V = float4;
    static __device__ __forceinline__ V zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ V load(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void store(float *p, V v) { *reinterpret_cast<float4 *>(p) = v; }
    static __device__ __forceinline__ void fma(V &a, float t, V x) {
        a.x = fmaf(t, x.x, a.x);
        a.y = fmaf(t, x.y, a.y);
        a.z = fmaf(t, x.z, a.z);
        a.w = fmaf(t, x.w, a.w);
    }
    static __device__ __forceinline__ double get(const V &v, int k) {
        return (double)(k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w);
    }
};
template <>
struct Vec<float, 2> {
    using V = float2;
    static __device__ __forceinline__ V zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ V load(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
    static __device__ __forceinline__ void store(float *p, V v) { *reinterpret_cast<float2 *>(p) = v; }
    static __device__ __forceinline__ void fma(V &a, float t, V x) {
        a.x = fmaf(t, x.x, a.x);
        a.y = fmaf(t, x.y, a.y);
    }
    static __device__ __forceinline__ double get(const V &v, int k) { return (double)(k == 0 ? v.x : v.y); }
};
template <>
struct Vec<float, 1> {
    using V = float;
    static __device__ __forceinline__ V zero() { return 0.f; }
    static __device__ __forceinline__ V load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void store(float *p, V v) { *p = v; }
    static __device__ __forceinline__ void fma(V &a, float t, V x) { a = fmaf(t, x, a); }
    static __device__ __forceinline__ double get(const V &v, int) { return (double)v; }
};
template <>
struct Vec<double, 2> {
    using V = double2;
    static __device__ __forceinline__ V zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ V load(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
    static __device__ __forceinline__ void store(double *p, V v) { *reinterpret_cast<double2 *>(p) = v; }
    static __device__ __forceinline__ void fma(V &a, float t, V x) {
        a.x = ::fma((double)t, x.x, a.x);
        a.y = ::fma((double)t, x.y, a.y);
    }
    static __device__ __forceinline__ double get(const V &v, int k) { return k == 0 ? v.x : v.y; }
};
template <>
struct Vec<double, 1> {
    using V = double;
    static __device__ __forceinline__ V zero() { return 0.0; }
    static __device__ __forceinline__ V load(const double *p) { return __ldg(p); }
    static __device__ __forceinline__ void store(double *p, V v) { *p = v; }
    static __device__ __forceinline__ void fma(V &a, float t, V x) { a = ::fma((double)t, x, a); }
    static __device__ __forceinline__ double get(const V &v, int) { return v; }
};

// dest_ids (may be null): original cell id of destination row r when the rows are in the plan's locality order
// (reorder.cu); it indexes the stationary vector and, if out_orig is set, the output row (permuted -> original).
// One update for destination rows [row_begin,row_end) and `ncols` columns starting at the column the
// base pointers address. eps_acc (may be null): [kEpsCopies][ncols][2] doubles, (x.pi, x.x) partial sums of
// this update; CTA b adds into copy b % kEpsCopies and eps_finalize_tile sums the copies.
// The products are added in ascending source order — the order scipy's csc_matvec adds them.
template <typename T>
struct GatherArgs {
    const int32_t *t_indptr;
    const Pair *pairs;
    const int32_t *dest_ids;
    int32_t out_orig;
    const T *Xin;
    uint32_t ld_in;
    T *Xout;
    uint32_t ld_out;
    int32_t ncols;
    const double *stationary;
    const T *pi_rows;  // stationary value of destination row r in the order the rows are walked (iterate precision)
    double *eps_acc;
    T *const *peer_out;
    int32_t npeers;
    const uint32_t *peer_mask;  // per destination row: bit g set = peer g needs the row (null = every peer)
};

// Destination rows [row0, row1) x column slab `slab` (32*VEC columns) by one CTA; `copy_sel` picks the eps
// accumulator copy this CTA adds to.
template <typename T, int VEC, bool EPS>
__device__ __forceinline__ void gather_block(const GatherArgs<T> &a, int32_t row0, int32_t row1, int slab, int copy_sel) {
    using VT = Vec<T, VEC>;
    using V = typename VT::V;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (slab * 32 + lane) * VEC;
    const bool active = col < a.ncols;
    const uint32_t ld_in = a.ld_in, ld_out = a.ld_out;
    const Pair *__restrict__ pairs = a.pairs;
    // cosine partial sums: a thread covers only rows_per_cta / kWarps rows, so its own partials are kept in the
    // iterate's precision (registers decide occupancy here, and the kernel is latency-bound); every sum across
    // warps and CTAs is float64
    T uv[VEC], uu[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) uv[k] = uu[k] = (T)0;
    // CTA-wide accumulators of the epilogue: warps add their partials with shared-memory atomics as they finish and the
    // LAST warp to finish forwards the sums — no barrier at the end of the CTA (a barrier + block reduction there cost
    // 10% of the kernel: 1.34e11 -> 1.48e11 cell-it/s without it, profiles/README_r01.md §8)
    __shared__ T s_acc[2][VEC][32];
    __shared__ unsigned s_done;
    if (EPS) {
        for (int t = threadIdx.x; t < 2 * VEC * 32; t += kWarps * 32) (&s_acc[0][0][0])[t] = (T)0;
        if (threadIdx.x == 0) s_done = 0u;
        __syncthreads();  // at the start of the CTA, where all warps arrive together
    }

    if (active) {
        const T *__restrict__ xin = a.Xin + col;
        for (int32_t r = row0 + warp; r < row1; r += kWarps) {
            int32_t p = __ldg(a.t_indptr + r);
            const int32_t end = __ldg(a.t_indptr + r + 1);
            // issued before the gathers so that their latency hides behind them (fetched after the row, the chain
            // dest id -> stationary value cost 17% of the kernel)
            const T pi = EPS ? __ldg(a.pi_rows + r) : (T)0;
            const int32_t orow = a.out_orig ? __ldg(a.dest_ids + r) : r;
            V acc = VT::zero();
            if (p < end && (p & 1)) {  // align the entry cursor to 16 bytes
                const Pair e = pairs[p];
                VT::fma(acc, e.val, VT::load(xin + (uint64_t)(uint32_t)e.src * ld_in));
                ++p;
            }
            for (; p + 3 < end; p += 4) {  // four row gathers in flight, (src,val) pairs by 128-bit loads
                const int4 q0 = __ldg(reinterpret_cast<const int4 *>(pairs + p));
                const int4 q1 = __ldg(reinterpret_cast<const int4 *>(pairs + p + 2));
                const V x0 = VT::load(xin + (uint64_t)(uint32_t)q0.x * ld_in);
                const V x1 = VT::load(xin + (uint64_t)(uint32_t)q0.z * ld_in);
                const V x2 = VT::load(xin + (uint64_t)(uint32_t)q1.x * ld_in);
                const V x3 = VT::load(xin + (uint64_t)(uint32_t)q1.z * ld_in);
                VT::fma(acc, __int_as_float(q0.y), x0);
                VT::fma(acc, __int_as_float(q0.w), x1);
                VT::fma(acc, __int_as_float(q1.y), x2);
                VT::fma(acc, __int_as_float(q1.w), x3);
            }
            if (p + 1 < end) {
                const int4 q = __ldg(reinterpret_cast<const int4 *>(pairs + p));
                const V x0 = VT::load(xin + (uint64_t)(uint32_t)q.x * ld_in);
                const V x1 = VT::load(xin + (uint64_t)(uint32_t)q.z * ld_in);
                VT::fma(acc, __int_as_float(q.y), x0);
                VT::fma(acc, __int_as_float(q.w), x1);
                p += 2;
            }
            if (p < end) {
                const Pair e = pairs[p];
                VT::fma(acc, e.val, VT::load(xin + (uint64_t)(uint32_t)e.src * ld_in));
            }
            const uint64_t off = (uint64_t)(uint32_t)orow * ld_out + col;
            if (a.peer_out == nullptr) {
                VT::store(a.Xout + off, acc);
            } else {
                // fused all-gather (row-partitioned propagation): the finished row goes straight into every rank's
                // next-iterate buffer — local HBM for this rank, NVLink peer stores for the others
                // (halo push: only the ranks whose rows gather this row, `peer_mask`)
                const uint32_t want = a.peer_mask ? __ldg(a.peer_mask + r) : 0xffffffffu;
                for (int32_t g = 0; g < a.npeers; ++g)
                    if ((want >> (g & 31)) & 1u) VT::store(a.peer_out[g] + off, acc);
            }
            if (EPS) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const T y = (T)VT::get(acc, k);
                    uv[k] = fma(y, pi, uv[k]);
                    uu[k] = fma(y, y, uu[k]);
                }
            }
        }
    }
    if (EPS) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            atomicAdd(&s_acc[0][k][lane], uv[k]);
            atomicAdd(&s_acc[1][k][lane], uu[k]);
        }
        __syncwarp();  // every lane's shared-memory atomics are issued before lane 0 takes the ticket
        __threadfence_block();
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(&s_done, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == kWarps - 1) {  // every warp's partials are in: one float64 atomic per (column, which) and CTA
            __threadfence_block();
            for (int t = lane; t < 2 * VEC * 32; t += 32) {
                const int which = t / (VEC * 32), k = (t / 32) % VEC, ln = t & 31;
                const int gcol = (slab * 32 + ln) * VEC + k;
                if (gcol < a.ncols)
                    atomicAdd(&a.eps_acc[((int64_t)copy_sel * a.ncols + gcol) * 2 + which],
                              (double)*(volatile T *)&s_acc[which][k][ln]);
            }
        }
    }
}

// Grid version: blockIdx.x = row group, blockIdx.y = column slab.
template <typename T, int VEC, bool EPS>
__global__ void __launch_bounds__(kWarps * 32, sizeof(T) == 4 ? 6 : 3)
spmm_gather(int32_t row_begin, int32_t row_end, int32_t rows_per_cta, GatherArgs<T> a) {
    const int32_t row0 = row_begin + blockIdx.x * rows_per_cta;
    gather_block<T, VEC, EPS>(a, row0, min(row0 + rows_per_cta, row_end), blockIdx.y, blockIdx.x % kEpsCopies);
}

// Persistent, SM-affine version for the long batched runs: SM s owns the contiguous row groups
// [s*R, (s+1)*R) for every column slab and its resident CTAs draw them from a per-SM counter, slab by slab, so the
// ~6 CTAs of an SM work on ADJACENT row groups of the SAME slab at the same time. In the plan's locality order
// adjacent rows share sources, which turns part of the L2->SM fill traffic into L1 hits. A CTA whose SM has run dry
// steals from the other SMs' queues. counters: [num_sms] ints, zero at launch.
template <typename T, int VEC, bool EPS>
__global__ void __launch_bounds__(kWarps * 32, sizeof(T) == 4 ? 6 : 3)
spmm_gather_sm(int32_t n_rows, int32_t rows_per_cta, int32_t nslabs, int32_t num_sms, int32_t *__restrict__ counters,
               GatherArgs<T> a) {
    __shared__ int32_t s_item;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int32_t nrg = (n_rows + rows_per_cta - 1) / rows_per_cta;  // row groups
    const int32_t R = (nrg + num_sms - 1) / num_sms;                 // row groups per SM
    for (int32_t v = 0; v < num_sms; ++v) {
        const int32_t owner = (int32_t)((smid + v) % (uint32_t)num_sms);
        const int32_t rg0 = owner * R, cnt = max(0, min(nrg, rg0 + R) - rg0);
        const int32_t total = cnt * nslabs;
        for (;;) {
            if (threadIdx.x == 0) s_item = total > 0 ? atomicAdd(&counters[owner], 1) : total;
            __syncthreads();
            const int32_t q = s_item;
            __syncthreads();
            if (q >= total) break;
            const int32_t slab = q / cnt, rg = rg0 + q % cnt;
            const int32_t row0 = rg * rows_per_cta;
            gather_block<T, VEC, EPS>(a, row0, min(row0 + rows_per_cta, n_rows), slab, rg % kEpsCopies);
        }
    }
}

// Staged gather for the permuted -> permuted updates (the bulk of a long batched run). The plan cuts the permuted
// rows into blocks whose distinct sources fit shared memory (reorder.cu). A CTA takes one block and walks a run of
// 256-byte column slabs with a two-stage pipeline: while the rows of slab s are formed out of one shared-memory
// buffer, cp.async copies every distinct source segment of slab s+1 from L2 into the other buffer (all copies
// independent, no registers held). A source that several rows of the block share costs one L2 -> SM transfer instead
// of one per row — the L2 -> SM fill path is what bounds spmm_gather (profiles/README_r01.md §2) — and the block's
// (index, value) pairs and row pointers are read once per run of slabs instead of once per slab.
constexpr int kSegBytes = 256;    // bytes of one staged source segment = 32 lanes x VEC x sizeof(T)
constexpr int kStagedWarps = 16;  // warps per CTA of the staged kernel
constexpr int kStagedMaxRows = 64, kStagedMaxPairs = 2048;

template <typename T>
struct StagedArgs {
    const int32_t *blk_row, *blk_uptr, *blk_usrc;
    const int32_t *r_indptr;
    const Pair *lpairs;
    const int32_t *dest_ids;
    const T *Xin;
    T *Xout;
    uint32_t ld;  // both iterates are scratch tiles with the same pitch
    int32_t ncols, nslabs, slabs_per_cta, ucap;
    const double *stationary;
    const T *pi_rows;
    double *eps_acc;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <typename T, int VEC, bool EPS>
__global__ void __launch_bounds__(kStagedWarps * 32, 1) spmm_staged(StagedArgs<T> a) {
    static_assert(32 * VEC * sizeof(T) == kSegBytes, "one warp spans one staged segment");
    using VT = Vec<T, VEC>;
    using V = typename VT::V;
    constexpr int kSegElems = kSegBytes / sizeof(T);
    extern __shared__ __align__(16) unsigned char s_dyn[];
    // layout: buf[2][ucap][kSegBytes] | pairs[kStagedMaxPairs] | rowptr[kStagedMaxRows + 1] | red
    unsigned char *s_buf = s_dyn;
    Pair *s_pairs = reinterpret_cast<Pair *>(s_dyn + 2 * (size_t)a.ucap * kSegBytes);
    int32_t *s_rp = reinterpret_cast<int32_t *>(s_pairs + kStagedMaxPairs);
    T *s_red = reinterpret_cast<T *>(s_rp + kStagedMaxRows + 4);  // [2][VEC][kStagedWarps][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int32_t u0 = __ldg(a.blk_uptr + b), U = __ldg(a.blk_uptr + b + 1) - u0;
    const int32_t row0 = __ldg(a.blk_row + b), row1 = __ldg(a.blk_row + b + 1);
    const int nrows = row1 - row0;
    const int slab0 = blockIdx.y * a.slabs_per_cta, slab1 = min(a.nslabs, slab0 + a.slabs_per_cta);
    const uint32_t ld = a.ld;
    // ---- prologue: this block's pairs and row pointers into shared memory; this warp's run of source ids into registers
    const int32_t p0 = __ldg(a.r_indptr + row0);
    const int32_t np = __ldg(a.r_indptr + row1) - p0;
    for (int i = threadIdx.x; i < np; i += kStagedWarps * 32) s_pairs[i] = a.lpairs[p0 + i];
    for (int i = threadIdx.x; i <= nrows; i += kStagedWarps * 32) s_rp[i] = __ldg(a.r_indptr + row0 + i) - p0;
    const int per_warp = (U + kStagedWarps - 1) / kStagedWarps;  // <= 32 (ucap <= 512)
    const int w0 = min(U, warp * per_warp), cnt = min(U, w0 + per_warp) - w0;
    const int32_t mine = lane < cnt ? __ldg(a.blk_usrc + u0 + w0 + lane) : 0;
    const int hw = lane >> 4, hl = lane & 15;
    const uint64_t pitch = (uint64_t)ld * sizeof(T);
    auto issue = [&](int slab, int stage) {  // half a warp per segment, 16 bytes per lane
        const unsigned char *xin = reinterpret_cast<const unsigned char *>(a.Xin + (size_t)slab * 32 * VEC) + hl * 16;
        unsigned char *dst = s_buf + (size_t)stage * a.ucap * kSegBytes + (size_t)w0 * kSegBytes + hl * 16;
#pragma unroll 4
        for (int i = 0; i < cnt; i += 2) {
            const int32_t src = __shfl_sync(0xffffffffu, mine, min(i + hw, cnt - 1));
            if (i + hw < cnt) cp_async16(dst + (size_t)(i + hw) * kSegBytes, xin + (uint64_t)(uint32_t)src * pitch);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(slab0, 0);
    for (int slab = slab0; slab < slab1; ++slab) {
        const int stage = (slab - slab0) & 1;
        if (slab + 1 < slab1) {
            issue(slab + 1, stage ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // slab's segments (and, the first time, the pairs) are visible to every warp
        const T *sx = reinterpret_cast<const T *>(s_buf + (size_t)stage * a.ucap * kSegBytes) + lane * VEC;
        const int col = (slab * 32 + lane) * VEC;
        T uv[VEC], uu[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) uv[k] = uu[k] = (T)0;
        for (int rr = warp; rr < nrows; rr += kStagedWarps) {
            int p = s_rp[rr];
            const int end = s_rp[rr + 1];
            V acc = VT::zero();
            for (; p + 3 < end; p += 4) {  // same add order as spmm_gather: ascending permuted source
                const Pair e0 = s_pairs[p], e1 = s_pairs[p + 1], e2 = s_pairs[p + 2], e3 = s_pairs[p + 3];
                const V x0 = *reinterpret_cast<const V *>(sx + e0.src * kSegElems);
                const V x1 = *reinterpret_cast<const V *>(sx + e1.src * kSegElems);
                const V x2 = *reinterpret_cast<const V *>(sx + e2.src * kSegElems);
                const V x3 = *reinterpret_cast<const V *>(sx + e3.src * kSegElems);
                VT::fma(acc, e0.val, x0);
                VT::fma(acc, e1.val, x1);
                VT::fma(acc, e2.val, x2);
                VT::fma(acc, e3.val, x3);
            }
            for (; p < end; ++p) {
                const Pair e = s_pairs[p];
                VT::fma(acc, e.val, *reinterpret_cast<const V *>(sx + e.src * kSegElems));
            }
            const int32_t r = row0 + rr;
            VT::store(a.Xout + (uint64_t)(uint32_t)r * ld + col, acc);
            if (EPS) {
                const T pi = __ldg(a.pi_rows + r);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const T y = (T)VT::get(acc, k);
                    uv[k] = fma(y, pi, uv[k]);
                    uu[k] = fma(y, y, uu[k]);
                }
            }
        }
        if (EPS) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                s_red[((0 * VEC + k) * kStagedWarps + warp) * 32 + lane] = uv[k];
                s_red[((1 * VEC + k) * kStagedWarps + warp) * 32 + lane] = uu[k];
            }
        }
        __syncthreads();  // all reads of this stage's buffer are done: the copy issued next iteration may overwrite it
        if (EPS) {
            for (int t = threadIdx.x; t < 2 * VEC * 32; t += kStagedWarps * 32) {
                const int which = t / (VEC * 32), k = (t / 32) % VEC, ln = t & 31;
                const int gcol = (slab * 32 + ln) * VEC + k;
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < kStagedWarps; ++w) sum += (double)s_red[((which * VEC + k) * kStagedWarps + w) * 32 + ln];
                atomicAdd(&a.eps_acc[((int64_t)(b % kEpsCopies) * a.ncols + gcol) * 2 + which], sum);
            }
            // s_red is rewritten only after the next iteration's first barrier
        }
    }
}

// B == 1, all iterations in one cooperative launch.
//   hist != null: row it of hist is the source of update it, row it+1 (or `last` for the final
//                 update) its destination; hist[0] must already hold x0.
//   hist == null: x0 -> bufA -> bufB -> bufA ...; the final update writes `last`.
// eps_acc: [iters][2] doubles (zeroed), or null.
template <typename T>
__global__ void __launch_bounds__(256)
spmv_b1_persistent(int32_t n, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs, const T *x0,
                   T *hist, T *bufA, T *bufB, T *last, int32_t iters, const double *__restrict__ stationary,
                   double *eps_acc) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    __shared__ double red[8][2];

    for (int32_t it = 0; it < iters; ++it) {
        const T *src;
        T *dst;
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
            T acc = (T)0;
            for (int32_t p = beg + lane; p < end; p += 32) {
                const Pair e = pairs[p];
                // L2 load: src was written earlier in this launch (L1 is not coherent across SMs)
                acc = fma((T)e.val, __ldcg(src + e.src), acc);
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

// B == 1 for small graphs (the reference's typical 3-10k cell data sets): ONE thread-block cluster runs every
// iteration; the per-update barrier is the hardware cluster barrier (~0.2 us) instead of a grid-wide one (~5 us), each
// CTA keeps the (src,val) pairs of its destination rows in shared memory, a warp gathers four rows at a time so the
// four L2 loads overlap, and the iterate stays in L2 (loads bypass L1, which is not coherent across SMs).
//   eps_part: [iters][nctas * 32 warps][2] doubles written with plain stores (no atomics), or null
template <typename T>
__global__ void __launch_bounds__(1024)
spmv_b1_cluster(int32_t n, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs, const T *x0, T *hist,
                T *bufA, T *bufB, T *last, int32_t iters, const double *__restrict__ stationary, double *eps_part,
                int32_t pairs_in_smem) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nctas = (int)cluster.num_blocks();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int m = (n + nctas - 1) / nctas;
    const int r0 = min(n, rank * m), r1 = min(n, r0 + m);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *s_ip = reinterpret_cast<int32_t *>(smem_raw);                        // [m + 1] row pointers of my slice
    Pair *s_pairs = reinterpret_cast<Pair *>(smem_raw + (((size_t)m + 2) / 2) * 8);  // 8-byte aligned
    const int32_t p_base = t_indptr[r0];
    for (int32_t i = threadIdx.x; i <= r1 - r0; i += blockDim.x) s_ip[i] = t_indptr[r0 + i];
    if (pairs_in_smem) {
        const int32_t cnt = t_indptr[r1] - p_base;
        for (int32_t i = threadIdx.x; i < cnt; i += blockDim.x) s_pairs[i] = pairs[p_base + i];
    }
    __syncthreads();
    const Pair *pp = pairs_in_smem ? s_pairs - p_base : pairs;  // indexable by the global entry position
    // 8 lanes per destination row (kNN rows hold ~30 entries: 4 per lane), 4 rows per warp pass, 3 shuffle steps
    constexpr int G = 8, RPW = 32 / G;
    const int gl = lane % G, grp = lane / G;
    for (int32_t it = 0; it < iters; ++it) {
        const T *src;
        T *dst;
        if (hist) {
            src = hist + (int64_t)it * n;
            dst = (it + 1 < iters) ? hist + (int64_t)(it + 1) * n : last;
        } else {
            src = it == 0 ? x0 : ((it & 1) ? bufA : bufB);
            dst = (it + 1 == iters) ? last : ((it & 1) ? bufB : bufA);
        }
        double uv = 0.0, uu = 0.0;
        for (int32_t rb = r0 + warp * RPW; rb < r1; rb += nw * RPW) {
            const int32_t r = rb + grp;
            const bool ok = r < r1;
            const int32_t beg = ok ? s_ip[r - r0] : 0, end = ok ? s_ip[r - r0 + 1] : 0;
            T acc = (T)0;
            for (int32_t p0 = beg + gl; p0 < end; p0 += 4 * G) {  // four gathers in flight per lane
                Pair e[4];
                T xv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int32_t p = p0 + k * G;
                    e[k] = p < end ? pp[p] : Pair{0, 0.f};
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) xv[k] = (p0 + k * G < end) ? __ldcg(src + e[k].src) : (T)0;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fma((T)e[k].val, xv[k], acc);
            }
#pragma unroll
            for (int o = G / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (ok && gl == 0) {
                dst[r] = acc;
                if (eps_part) {
                    const double yd = (double)acc;
                    uv = fma(yd, stationary[r], uv);
                    uu = fma(yd, yd, uu);
                }
            }
        }
        if (eps_part) {  // per-warp partial sums, plain stores; eps_finalize_tile adds the nctas * nw partials
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) {
                uv += __shfl_xor_sync(0xffffffffu, uv, o);
                uu += __shfl_xor_sync(0xffffffffu, uu, o);
            }
            if (lane == 0) {
                double *dstp = eps_part + ((int64_t)it * nctas * nw + (int64_t)rank * nw + warp) * 2;
                dstp[0] = uv;
                dstp[1] = uu;
            }
        }
        cluster.sync();  // release/acquire at cluster scope: the new iterate is visible to every CTA of the cluster
    }
}

// Same cluster, but the iterate itself lives in (distributed) shared memory: every CTA keeps a full double-buffered
// copy of x, reads its gathers from its own copy (LDS instead of an L2 round trip) and pushes each new value into
// all CTAs' copies with st.shared::cluster. The per-update barrier then only has to order shared-memory traffic: the
// global-memory release/acquire of the variant above costs ~5 us per update on B200, this one ~1 us. The history
// row (global) is written with plain stores; nobody reads it inside the kernel.
template <typename T>
__global__ void __launch_bounds__(1024)
spmv_b1_cluster_dsmem(int32_t n, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs, const T *x0,
                      T *hist, T *last, int32_t iters, const double *__restrict__ stationary, double *eps_part) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nctas = (int)cluster.num_blocks();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int m = (n + nctas - 1) / nctas;
    const int r0 = min(n, rank * m), r1 = min(n, r0 + m);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_x = reinterpret_cast<T *>(smem_raw);                                 // [2][n]
    int32_t *s_ip = reinterpret_cast<int32_t *>(s_x + 2 * (size_t)n);         // [m + 1]
    Pair *s_pairs = reinterpret_cast<Pair *>(reinterpret_cast<unsigned char *>(s_ip) + (((size_t)m + 2) / 2) * 8);
    const int32_t p_base = t_indptr[r0];
    for (int32_t i = threadIdx.x; i <= r1 - r0; i += blockDim.x) s_ip[i] = t_indptr[r0 + i];
    const int32_t cnt = t_indptr[r1] - p_base;
    for (int32_t i = threadIdx.x; i < cnt; i += blockDim.x) s_pairs[i] = pairs[p_base + i];
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) s_x[i] = x0[i];
    cluster.sync();  // every CTA's shared memory is initialised before anyone stores into it remotely
    const Pair *pp = s_pairs - p_base;
    constexpr int G = 8, RPW = 32 / G;
    const int gl = lane % G, grp = lane / G;
    for (int32_t it = 0; it < iters; ++it) {
        const T *cur = s_x + (size_t)(it & 1) * n;
        T *nxt_local = s_x + (size_t)((it + 1) & 1) * n;
        T *gdst = hist ? ((it + 1 < iters) ? hist + (int64_t)(it + 1) * n : last) : (it + 1 == iters ? last : nullptr);
        double uv = 0.0, uu = 0.0;
        for (int32_t rb = r0 + warp * RPW; rb < r1; rb += nw * RPW) {
            const int32_t r = rb + grp;
            const bool ok = r < r1;
            const int32_t beg = ok ? s_ip[r - r0] : 0, end = ok ? s_ip[r - r0 + 1] : 0;
            T acc = (T)0;
            for (int32_t p = beg + gl; p < end; p += G) {
                const Pair e = pp[p];
                acc = fma((T)e.val, cur[e.src], acc);
            }
#pragma unroll
            for (int o = G / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (ok) {
                // every lane of the group holds the row sum: lane gl pushes it to CTAs gl, gl + G, ...
                for (int dstc = gl; dstc < nctas; dstc += G) cluster.map_shared_rank(nxt_local, dstc)[r] = acc;
                if (gl == 0) {
                    if (gdst) gdst[r] = acc;
                    if (eps_part) {
                        const double yd = (double)acc;
                        uv = fma(yd, stationary[r], uv);
                        uu = fma(yd, yd, uu);
                    }
                }
            }
        }
        if (eps_part) {
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) {
                uv += __shfl_xor_sync(0xffffffffu, uv, o);
                uu += __shfl_xor_sync(0xffffffffu, uu, o);
            }
            if (lane == 0) {
                double *dstp = eps_part + ((int64_t)it * nctas * nw + (int64_t)rank * nw + warp) * 2;
                dstp[0] = uv;
                dstp[1] = uu;
            }
        }
        cluster.sync();
    }
}

// B == 1 update of a row range (row-partitioned propagation, config 5): 8 lanes per destination row, four gathers in
// flight per lane, 3 shuffle steps. The general kernel would leave 31 of 32 lanes idle at B == 1.
template <typename T>
__global__ void __launch_bounds__(256)
spmv_rows(int32_t row_begin, int32_t row_end, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs,
          const T *__restrict__ x, T *__restrict__ y, T *const *__restrict__ peer_out, int32_t npeers,
          const uint32_t *__restrict__ peer_mask, const float *__restrict__ lo = nullptr /* float64 matrix: value = val + lo */) {
    constexpr int G = 8;
    const int lane = threadIdx.x & 31, gl = lane % G;
    const int32_t r = row_begin + (int32_t)((blockIdx.x * (blockDim.x / G)) + threadIdx.x / G);
    const bool ok = r < row_end;
    const int32_t beg = ok ? t_indptr[r] : 0, end = ok ? t_indptr[r + 1] : 0;
    T acc = (T)0;
    for (int32_t p0 = beg + gl; p0 < end; p0 += 4 * G) {
        Pair e[4];
        T xv[4], tv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int32_t p = p0 + k * G;
            e[k] = p < end ? pairs[p] : Pair{0, 0.f};
            tv[k] = (T)e[k].val;
            if (lo != nullptr && p < end) tv[k] += (T)lo[p];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) xv[k] = (p0 + k * G < end) ? __ldg(x + e[k].src) : (T)0;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma(tv[k], xv[k], acc);
    }
#pragma unroll
    for (int o = G / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (ok) {
        if (peer_out == nullptr) {
            if (gl == 0) y[r] = acc;
        } else {  // every lane of the group holds the sum: lane gl serves peers gl, gl + G, ...
            const uint32_t want = peer_mask ? __ldg(peer_mask + r) : 0xffffffffu;
            for (int32_t g = gl; g < npeers; g += G)
                if ((want >> (g & 31)) & 1u) peer_out[g][r] = acc;
        }
    }
}

// 2 <= B <= 16 columns (the models' node-level products with T: B = S + L = 14 at config 3; the subspace iteration of
// stationary.py: B = 10): a half-warp per destination row, lane = column, eight entries and eight row gathers in
// flight. The general kernel gives such a row a whole warp with most lanes idle and one dependent (entry, gather) pair
// at a time: 47 us for 50,000 rows x 30 entries x 14 columns (ncu: 27 cycles of latency per issued instruction).
template <typename T>
__global__ void __launch_bounds__(256)
spmm_rows_narrow(int32_t row_begin, int32_t row_end, const int32_t *__restrict__ t_indptr, const Pair *__restrict__ pairs,
                 const T *__restrict__ x, int32_t B, T *__restrict__ y) {
    const int lane = threadIdx.x & 31, half = lane >> 4, col = lane & 15;
    const int32_t r = row_begin + (int32_t)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + half);
    const bool ok = r < row_end;
    const int32_t beg = ok ? t_indptr[r] : 0, end = ok ? t_indptr[r + 1] : 0;
    const bool live = col < B;
    T acc = (T)0;
    for (int32_t p0 = beg; p0 < end; p0 += 8) {
        Pair e[8];
        T xv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = p0 + k < end ? pairs[p0 + k] : Pair{0, 0.f};
#pragma unroll
        for (int k = 0; k < 8; ++k) xv[k] = (live && p0 + k < end) ? __ldg(x + (size_t)e[k].src * B + col) : (T)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fma((T)e[k].val, xv[k], acc);
    }
    if (ok && live) y[(size_t)r * B + col] = acc;
}

// Persistent single-start update of a row range for the row-partitioned run (config 5, B == 1): ALL updates in one
// cooperative launch. Per update: the rank's rows (8 lanes per row, as spmv_rows) are formed from its full-size
// iterate and stored into the next-iterate buffers of the ranks that gather them (own HBM + NVLink peer stores, per-row
// masks); then grid barrier -> cross-GPU flag exchange by one CTA (every rank writes its epoch into its slot of every
// rank's flag array and waits for all slots of its own: the protocol of cy2p_xbarrier, propagate_x.cu) -> grid barrier.
// No kernel launch and no host round trip per update: at 10^6 cells on 8 GPUs an update is ~6 us of gathers.
// The iterate is read with ordinary L1-cached loads (ld.global.ca — in the locality order neighbouring rows share
// sectors, and the L2-only loads of a first version ran the gather 2x slower): rows written by peers and by this rank one
// update ago cannot be served from a stale L1 line because every CTA passes the gpu-scope fences of grid.sync() after
// the flag exchange (a gpu-scope fence invalidates the SM's L1), and the peers' rows are in this GPU's L2 before their flag
// (fence.sys on the writer's side). A barrier that is not reached within ~2 s sets *d_error and stops waiting for good.
__device__ __forceinline__ float ld_ca(const float *p) {
    float v;
    asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_ca(const double *p) {
    double v;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
template <typename T>
__global__ void __launch_bounds__(256)
spmv_rows_persist(int32_t row_begin, int32_t row_end, const int32_t *__restrict__ t_indptr,
                  const Pair *__restrict__ pairs, T *const *__restrict__ peer_buf0, T *const *__restrict__ peer_buf1,
                  const uint32_t *__restrict__ peer_mask, int32_t rank, int32_t world, int32_t iters,
                  int32_t *const *__restrict__ peer_flags, volatile int32_t *my_flags, int32_t *epoch_counter,
                  int32_t *d_error) {
    cg::grid_group grid = cg::this_grid();
    constexpr int G = 8;
    const int lane = threadIdx.x & 31, gl = lane % G;
    const int32_t epoch0 = *(volatile int32_t *)epoch_counter;
    const int32_t rows_per_pass = gridDim.x * (blockDim.x / G);
    bool dead = false;
    for (int32_t it = 0; it < iters; ++it) {
        T *const *src_tab = (it & 1) ? peer_buf1 : peer_buf0;
        T *const *dst_tab = (it & 1) ? peer_buf0 : peer_buf1;
        const T *x = src_tab[rank];
        // the loop bound is WARP-uniform (a warp covers 32 / G consecutive rows): all 32 lanes reach the shuffles
        for (int32_t rw = row_begin + (int32_t)(blockIdx.x * (blockDim.x / G) + (threadIdx.x >> 5) * (32 / G)); rw < row_end;
             rw += rows_per_pass) {
            const int32_t r = rw + lane / G;
            const bool ok = r < row_end;
            const int32_t beg = ok ? __ldg(t_indptr + r) : 0, end = ok ? __ldg(t_indptr + r + 1) : 0;
            T acc = (T)0;
            for (int32_t p0 = beg + gl; p0 < end; p0 += 4 * G) {
                Pair e[4];
                T xv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int32_t p = p0 + k * G;
                    e[k] = p < end ? pairs[p] : Pair{0, 0.f};
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) xv[k] = (p0 + k * G < end) ? ld_ca(x + e[k].src) : (T)0;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = fma((T)e[k].val, xv[k], acc);
            }
#pragma unroll
            for (int o = G / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (ok) {
                const uint32_t want = peer_mask ? __ldg(peer_mask + r) : 0xffffffffu;
                for (int32_t g = gl; g < world; g += G)
                    if ((want >> (g & 31)) & 1u) dst_tab[g][r] = acc;
            }
        }
        __threadfence_system();
        grid.sync();
        if (blockIdx.x == 0 && (int)threadIdx.x < world && !dead) {
            const int32_t epoch = epoch0 + it + 1;
            *(volatile int32_t *)(peer_flags[threadIdx.x] + rank) = epoch;
            const long long t0 = clock64();
            while (my_flags[threadIdx.x] - epoch < 0) {
                if (clock64() - t0 > 4000000000LL || *(volatile int32_t *)d_error) {
                    atomicExch(d_error, 1);
                    dead = true;
                    break;
                }
            }
            __threadfence_system();
        }
        grid.sync();
        __threadfence();  // (explicit: the L1 of this SM holds nothing older than the exchange)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *epoch_counter = epoch0 + iters;
}

// pi_rows[r] = (T) stationary[dest_ids ? dest_ids[r] : r]: the stationary vector in the order the rows are walked
template <typename T>
__global__ void pi_rows_kernel(int32_t n, const double *__restrict__ stationary, const int32_t *__restrict__ dest_ids,
                               T *__restrict__ out) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r] = (T)stationary[dest_ids ? dest_ids[r] : r];
}

__global__ void sumsq_kernel(int32_t n, const double *__restrict__ v, double *out) {
    double s = 0.0;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s = fma(v[i], v[i], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// One update's eps of a single float64 row, one CTA, fixed summation order (bit-reproducible): used by the float64-
// matrix single-start path below.
__global__ void __launch_bounds__(1024)
eps_row_kernel(int32_t n, const double *__restrict__ y, const double *__restrict__ pi, const double *__restrict__ vv,
               double *__restrict__ eps_out) {
    __shared__ double s_uv[32], s_uu[32];
    double uv = 0.0, uu = 0.0;
    for (int32_t i = threadIdx.x; i < n; i += 1024) {
        const double v = y[i];
        uv = fma(v, pi[i], uv);
        uu = fma(v, v, uu);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        uv += __shfl_xor_sync(0xffffffffu, uv, o);
        uu += __shfl_xor_sync(0xffffffffu, uu, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_uv[threadIdx.x >> 5] = uv;
        s_uu[threadIdx.x >> 5] = uu;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uv = 0.0;
        uu = 0.0;
        for (int k = 0; k < 32; ++k) {
            uv += s_uv[k];
            uu += s_uu[k];
        }
        const double d = 1.0 - uv / sqrt(uu * vv[0]);
        eps_out[0] = fmin(fmax(d, 0.0), 2.0);
    }
}

// eps[i] = clip(1 - uv/sqrt(uu*vv), 0, 2): scipy.spatial.distance.cosine as called at sampling.py:44.
// acc is [iters][copies][tile_cols][2] for the column tile starting at col0; eps is [iters][B].
__global__ void eps_finalize_tile(int32_t iters, int32_t copies, int32_t tile_cols, int32_t col0, int32_t B,
                                  const double *__restrict__ acc, const double *__restrict__ vv,
                                  double *__restrict__ eps) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)iters * tile_cols) return;
    const int32_t it = (int32_t)(i / tile_cols), c = (int32_t)(i % tile_cols);
    const double *a = acc + ((int64_t)it * copies * tile_cols + c) * 2;
    double uv = 0.0, uu = 0.0;
    for (int32_t k = 0; k < copies; ++k) {
        uv += a[(int64_t)k * tile_cols * 2];
        uu += a[(int64_t)k * tile_cols * 2 + 1];
    }
    const double d = 1.0 - uv / sqrt(uu * vv[0]);
    eps[(int64_t)it * B + col0 + c] = fmin(fmax(d, 0.0), 2.0);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// Column-tile width of the ping-pong scratch. CY2P_TILE overrides (tuning knob).
static int32_t choose_tile(const cy2p_plan *plan, int32_t B, size_t elem) {
    int64_t bt = env_int("CY2P_TILE", 0);
    if (bt <= 0) {
        // Measured on B200 (profiles/k1_sweep_r01.md): wider tiles win (more CTAs per launch, fewer launches)
        // even when the two ping-pong tiles exceed L2; cap the scratch at ~1 GiB.
        bt = (int64_t)(((size_t)1 << 30) / ((size_t)plan->n * 2 * elem));
        if (bt > 2048) bt = 2048;
    }
    bt = bt / 128 * 128;
    if (bt < 128) bt = 128;
    if (bt > B) bt = B;
    return (int32_t)bt;
}

struct WsLayout {
    size_t eps_acc, vv, bufA, bufB, counters, pi_id, pi_perm, total;
    int32_t tile;
};
constexpr int kCounterSlots = 8192;  // per-launch SM work counters, zeroed in blocks of this many launches
constexpr int kMaxSms = 256;

static WsLayout ws_layout(const cy2p_plan *plan, int32_t B, int32_t iters, size_t elem) {
    WsLayout w;
    w.tile = choose_tile(plan, B, elem);
    size_t off = 0;
    w.eps_acc = off;
    off = align_up(off + sizeof(double) * 2 * (size_t)iters * ((size_t)w.tile * kEpsCopies > 512 ? (size_t)w.tile * kEpsCopies : 512), 256);
    w.vv = off;
    off = align_up(off + sizeof(double), 256);
    w.bufA = off;
    off = align_up(off + elem * (size_t)plan->n * (size_t)w.tile, 256);
    w.bufB = off;
    off = align_up(off + elem * (size_t)plan->n * (size_t)w.tile, 256);
    w.counters = off;
    off = align_up(off + sizeof(int32_t) * (size_t)kCounterSlots * kMaxSms, 256);
    w.pi_id = off;
    off = align_up(off + elem * (size_t)plan->n, 256);
    w.pi_perm = off;
    off = align_up(off + elem * (size_t)plan->n, 256);
    w.total = off;
    return w;
}

// Which T^T the update walks: the caller's order (identity) or the plan's locality order, with sources addressed in
// the original or the permuted iterate.
struct GatherOperand {
    const int32_t *indptr;
    const Pair *pairs;
    const int32_t *dest_ids;  // null for the identity order
    int32_t out_orig;
    void *const *peer_out = nullptr;  // device array of `npeers` output base pointers (fused all-gather), or null
    int32_t npeers = 0;
    const uint32_t *peer_mask = nullptr;
};

template <typename T, int VEC>
static void launch_gather(const GatherOperand &g, int32_t row_begin, int32_t row_end, const T *in, uint32_t ld_in,
                          T *out, uint32_t ld_out, int32_t ncols, const double *stationary, double *eps_acc,
                          cudaStream_t st, int32_t *sm_counters = nullptr, int32_t num_sms = 0,
                          const T *pi_rows = nullptr) {
    // rows per CTA: 64 for wide tiles (each warp forms 8 rows; +3% over 32 with the grouped order, 1.31e11 vs 1.27e11
    // cell-it/s at 50k cells, profiles/k1_sweep7_r01.jsonl), 32 for narrow ones where 64 leaves too few CTAs
    static const int rows_env = env_int("CY2P_ROWS_PER_CTA", 0);
    const int rows_per_cta = rows_env > 0 ? std::max(rows_env, kWarps) : (ncols >= 512 ? 64 : 32);
    GatherArgs<T> a;
    a.t_indptr = g.indptr;
    a.pairs = g.pairs;
    a.dest_ids = g.dest_ids;
    a.out_orig = g.out_orig;
    a.Xin = in;
    a.ld_in = ld_in;
    a.Xout = out;
    a.ld_out = ld_out;
    a.ncols = ncols;
    a.stationary = stationary;
    a.pi_rows = pi_rows;
    a.eps_acc = eps_acc;
    a.peer_out = (T *const *)g.peer_out;
    a.npeers = g.npeers;
    a.peer_mask = g.peer_mask;
    const int nslabs = (ncols + 32 * VEC - 1) / (32 * VEC);
    count_launch();
    if (sm_counters != nullptr && row_begin == 0) {  // persistent SM-affine schedule (counters zeroed by the caller)
        const int ctas = num_sms * (sizeof(T) == 4 ? 6 : 3);
        if (eps_acc)
            spmm_gather_sm<T, VEC, true><<<ctas, kWarps * 32, 0, st>>>(row_end, rows_per_cta, nslabs, num_sms, sm_counters, a);
        else
            spmm_gather_sm<T, VEC, false><<<ctas, kWarps * 32, 0, st>>>(row_end, rows_per_cta, nslabs, num_sms, sm_counters, a);
        return;
    }
    dim3 grid((unsigned)((row_end - row_begin + rows_per_cta - 1) / rows_per_cta), (unsigned)nslabs);
    if (eps_acc)
        spmm_gather<T, VEC, true><<<grid, kWarps * 32, 0, st>>>(row_begin, row_end, rows_per_cta, a);
    else
        spmm_gather<T, VEC, false><<<grid, kWarps * 32, 0, st>>>(row_begin, row_end, rows_per_cta, a);
}

template <typename T>
struct Wide {
    static constexpr int VEC = 16 / sizeof(T);  // 128-bit per lane
};

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

template <typename T>
static int32_t propagate_impl(cy2p_plan *plan, const T *d_X0, int32_t B, int32_t iters, const double *d_stationary,
                              T *d_Xfinal, T *d_history, int32_t history_stride, double *d_eps, void *d_ws,
                              size_t ws_bytes, void *stream) {
    constexpr int VW = Wide<T>::VEC;
    CY2P_REQUIRE(plan && d_X0, "null plan / X0");
    CY2P_REQUIRE(B >= 1 && iters >= 0, "B >= 1 and iters >= 0 required");
    CY2P_REQUIRE(!d_history || history_stride >= 1, "history_stride must be >= 1");
    CY2P_REQUIRE(!d_eps || d_stationary, "eps requested without a stationary vector");
    CY2P_REQUIRE(d_Xfinal || d_history, "nothing to write: both Xfinal and history are null");
    const WsLayout w = ws_layout(plan, B, iters, sizeof(T));
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
    T *bufA = (T *)(ws + w.bufA), *bufB = (T *)(ws + w.bufB);

    if (iters == 0) {
        if (d_Xfinal && d_Xfinal != d_X0)
            CY2P_CUDA(cudaMemcpyAsync(d_Xfinal, d_X0, sizeof(T) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));
        return CY2P_OK;
    }
    if (d_eps) {
        CY2P_CUDA(cudaMemsetAsync(vv, 0, sizeof(double), st));
        count_launch();
        sumsq_kernel<<<64, 256, 0, st>>>(n, d_stationary, vv);
    }
    if (d_history)  // row 0 of the history is the start distribution (sampling.py:39)
        CY2P_CUDA(cudaMemcpyAsync(d_history, d_X0, sizeof(T) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));

    // ---- B == 1 on a float64 matrix (cy2p_plan_create_host_f64): one launch per update with value = hi + lo. scvelo's
    // matrices are float32 and never come here; this path trades the single-launch kernels below for the exact product
    // the reference forms with a user-supplied float64 T (sampling.py:40-42). -------------------------------------------
    if constexpr (std::is_same<T, double>::value) {
        if (B == 1 && plan->d_t_lo && (!d_history || history_stride == 1)) {
            const double *src = d_X0;
            const unsigned grid = (unsigned)((n + 31) / 32);
            for (int32_t it = 0; it < iters; ++it) {
                const bool last = it + 1 == iters;
                // history row h is the iterate BEFORE update h (row 0 = the start, copied above)
                double *dst = (d_history && !last) ? d_history + (size_t)(it + 1) * n
                                                   : (last && d_Xfinal ? d_Xfinal : ((it & 1) ? bufA : bufB));
                count_launch();
                spmv_rows<double><<<grid, 256, 0, st>>>(0, n, plan->d_t_indptr, plan->d_t_pairs, src, dst, nullptr, 0,
                                                        nullptr, plan->d_t_lo);
                if (d_eps) {
                    count_launch();
                    eps_row_kernel<<<1, 1024, 0, st>>>(n, dst, d_stationary, vv, d_eps + it);
                }
                src = dst;
            }
            CY2P_CUDA(cudaGetLastError());
            return CY2P_OK;
        }
    }

    // ---- B == 1: all iterations in one launch --------------------------------------------------
    if (B == 1 && (!d_history || history_stride == 1)) {
        T *last = d_Xfinal ? d_Xfinal : bufB;
        int32_t n_ = n, iters_ = iters;
        const int32_t *tp = plan->d_t_indptr;
        const Pair *pp = plan->d_t_pairs;
        // (a) small graphs: one thread-block cluster (hardware barrier per update)
        static const int cluster_max_nnz = env_int("CY2P_CLUSTER_MAX_NNZ", 600000);
        if (plan->nnz <= cluster_max_nnz && plan->cluster_ctas >= 0) {
            if (plan->cluster_ctas == 0) {  // probe once per plan: largest cluster (16, then 8) that can be resident
                plan->cluster_ctas = -1;
                for (int c : {16, 8}) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(c);
                    cfg.blockDim = dim3(1024);
                    cfg.dynamicSmemBytes = 0;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = c;
                    at[0].val.clusterDim.y = 1;
                    at[0].val.clusterDim.z = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    if (c > 8 &&
                        cudaFuncSetAttribute(spmv_b1_cluster<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
                            cudaSuccess) {
                        cudaGetLastError();
                        continue;
                    }
                    int nclusters = 0;
                    if (cudaOccupancyMaxActiveClusters(&nclusters, spmv_b1_cluster<T>, &cfg) == cudaSuccess &&
                        nclusters >= 1) {
                        plan->cluster_ctas = c;
                        break;
                    }
                    cudaGetLastError();
                }
            }
            const int c = plan->cluster_ctas;
            if (c > 0) {
                // shared memory: the row pointers of a slice + the (src,val) pairs of the fattest slice
                const int m = (n + c - 1) / c;
                if (plan->h_t_indptr.empty()) {
                    plan->h_t_indptr.resize((size_t)n + 1);
                    CY2P_CUDA(cudaMemcpyAsync(plan->h_t_indptr.data(), plan->d_t_indptr, sizeof(int32_t) * ((size_t)n + 1),
                                              cudaMemcpyDeviceToHost, st));
                    CY2P_CUDA(cudaStreamSynchronize(st));
                }
                size_t worst = 0;
                for (int r = 0; r < c; ++r) {
                    const int a = std::min(n, r * m), b = std::min(n, a + m);
                    worst = std::max(worst, (size_t)(plan->h_t_indptr[b] - plan->h_t_indptr[a]));
                }
                const size_t ip_bytes = (((size_t)m + 2) / 2) * 8;
                size_t smem = ip_bytes + worst * sizeof(Pair);
                int32_t in_smem = smem <= 200 * 1024 ? 1 : 0;
                if (!in_smem) smem = ip_bytes;
                // (a1) iterate in distributed shared memory when two copies of x also fit
                const size_t smem_ds = 2 * (size_t)n * sizeof(T) + ip_bytes + worst * sizeof(Pair);
                static const int use_dsmem = env_int("CY2P_B1_DSMEM", 1);
                if (use_dsmem && smem_ds <= 200 * 1024 && (d_history || d_Xfinal)) {
                    if (c > 8)
                        CY2P_CUDA(cudaFuncSetAttribute(spmv_b1_cluster_dsmem<T>,
                                                       cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                    CY2P_CUDA(cudaFuncSetAttribute(spmv_b1_cluster_dsmem<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)smem_ds));
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(c);
                    cfg.blockDim = dim3(1024);
                    cfg.dynamicSmemBytes = smem_ds;
                    cfg.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = c;
                    at[0].val.clusterDim.y = 1;
                    at[0].val.clusterDim.z = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    double *part = d_eps ? eps_acc : nullptr;
                    count_launch();
                    cudaError_t le = cudaLaunchKernelEx(&cfg, spmv_b1_cluster_dsmem<T>, n_, tp, pp, d_X0, d_history, last,
                                                        iters_, d_stationary, part);
                    if (le == cudaSuccess) {
                        if (d_eps) {
                            count_launch();
                            eps_finalize_tile<<<(iters + 255) / 256, 256, 0, st>>>(iters, c * 32, 1, 0, 1, eps_acc, vv, d_eps);
                        }
                        CY2P_CUDA(cudaGetLastError());
                        return CY2P_OK;
                    }
                    cudaGetLastError();  // not launchable with this much shared memory per cluster: fall through
                }
                if (c > 8)
                    CY2P_CUDA(cudaFuncSetAttribute(spmv_b1_cluster<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                CY2P_CUDA(cudaFuncSetAttribute(spmv_b1_cluster<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(c);
                cfg.blockDim = dim3(1024);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = c;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                double *part = d_eps ? eps_acc : nullptr;  // [iters][c * 32][2], written with plain stores
                count_launch();
                CY2P_CUDA(cudaLaunchKernelEx(&cfg, spmv_b1_cluster<T>, n_, tp, pp, d_X0, d_history, bufA, bufB, last, iters_,
                                             d_stationary, part, in_smem));
                if (d_eps) {
                    count_launch();
                    eps_finalize_tile<<<(iters + 255) / 256, 256, 0, st>>>(iters, c * 32, 1, 0, 1, eps_acc, vv, d_eps);
                }
                CY2P_CUDA(cudaGetLastError());
                return CY2P_OK;
            }
        }
        // (b) one cooperative launch, grid barrier per update
        if (d_eps) CY2P_CUDA(cudaMemsetAsync(eps_acc, 0, sizeof(double) * 2 * (size_t)iters, st));
        int per_sm = 0;
        CY2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_b1_persistent<T>, 256, 0));
        int grid = plan->num_sms * (per_sm > 2 ? 2 : per_sm);
        const int need = (n + 7) / 8;
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        double *acc = d_eps ? eps_acc : nullptr;
        void *args[] = {&n_, &tp, &pp, (void *)&d_X0, &d_history, &bufA, &bufB, &last, &iters_, (void *)&d_stationary, &acc};
        count_launch();
        CY2P_CUDA(cudaLaunchCooperativeKernel((void *)spmv_b1_persistent<T>, dim3(grid), dim3(256), args, 0, st));
        if (d_eps) {
            count_launch();
            eps_finalize_tile<<<(iters + 255) / 256, 256, 0, st>>>(iters, 1, 1, 0, 1, eps_acc, vv, d_eps);
        }
        CY2P_CUDA(cudaGetLastError());
        return CY2P_OK;
    }

    // ---- general B: column tiles, each carried through all iterations while it is L2-resident -----
    const bool wide = (B % VW == 0) && aligned16(d_X0) && (!d_Xfinal || aligned16(d_Xfinal)) &&
                      (!d_history || aligned16(d_history));
    const int64_t slots = d_history ? ((int64_t)iters + history_stride - 1) / history_stride : 0;
    // locality ordering: worth building when the work amortises the one-off host pass
    if (plan->reorder_state == 0 && (int64_t)B * iters >= 4096 && B >= 32) {
        const int32_t rc = ensure_reordered(plan, st);
        if (rc != CY2P_OK) return rc;
    }
    const bool reord = plan->reorder_state == 1;
    T *pi_rows = nullptr;  // stationary values in the order the destination rows are walked by every launch below
    if (d_eps) {
        pi_rows = (T *)(ws + (reord ? w.pi_perm : w.pi_id));
        count_launch();
        pi_rows_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(n, d_stationary, reord ? plan->d_order : nullptr, pi_rows);
    }
    // Opt-in experiment (CY2P_STAGED=1, read when the plan is ordered and at every call): measured SLOWER than
    // spmm_gather on B200 — 0.64-0.76e11 vs 1.26e11 cell-it/s at 50k cells (profiles/k1_sweep6_staged_r01.jsonl).
    // It does cut the L2 -> SM sectors by 45% (197 M vs 357 M per launch), but forming the rows out of shared memory
    // costs three shared-memory wavefronts and ~6 instructions per 256-byte segment with one CTA of 16 warps per SM,
    // against one 512-byte LDG.128 and ~10 instructions per segment at 48 warps per SM in the plain kernel.
    const int staged_env = env_int("CY2P_STAGED", 0);
    static const int staged_spc = std::max(1, env_int("CY2P_STAGE_SLABS", 8));  // column slabs a CTA walks
    bool staged = staged_env && reord && plan->nblk > 0;
    size_t staged_smem = 0;
    if (staged) {
        constexpr int SV = kSegBytes / 32 / (int)sizeof(T);
        staged_smem = 2 * (size_t)plan->blk_ucap * kSegBytes + sizeof(Pair) * kStagedMaxPairs +
                      sizeof(int32_t) * (kStagedMaxRows + 4) + sizeof(T) * 2 * SV * kStagedWarps * 32;
        cudaFuncSetAttribute(spmm_staged<T, SV, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(spmm_staged<T, SV, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (staged_smem > 226 * 1024 || plan->blk_ucap > 32 * kStagedWarps ||
            cudaFuncSetAttribute(spmm_staged<T, SV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem) != cudaSuccess ||
            cudaFuncSetAttribute(spmm_staged<T, SV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem) != cudaSuccess) {
            cudaGetLastError();
            staged = false;
        }
    }
    const GatherOperand ident = {plan->d_t_indptr, plan->d_t_pairs, nullptr, 0};
    // SM-affine persistent schedule (experiment, off by default): measured SLOWER than the plain grid on B200
    // (1.01e11 vs 1.23e11 cell-it/s at 50k cells, 0.81e11 vs 1.21e11 at 250k, profiles/k1_sweep5_sm_affine_r01.jsonl):
    // the plain grid's wave front covers one contiguous band of the locality-ordered rows, which is what keeps the
    // gathers in L2; 148 SMs working 148 distant bands lose that. CY2P_SM_AFFINE=1 enables it.
    static const int sm_affine_env = env_int("CY2P_SM_AFFINE", 0);
    const bool sm_affine = sm_affine_env && reord && n >= 8192 && plan->num_sms <= kMaxSms;
    int32_t *counters = (int32_t *)(ws + w.counters);
    int64_t launch_no = 0;
    for (int32_t c0 = 0; c0 < B; c0 += w.tile) {
        const int32_t nc = (B - c0 < w.tile) ? B - c0 : w.tile;
        if (d_eps)
            CY2P_CUDA(cudaMemsetAsync(eps_acc, 0, sizeof(double) * 2 * (size_t)iters * (size_t)nc * kEpsCopies, st));
        const T *src = d_history ? d_history + c0 : d_X0 + c0;  // history row 0 holds x0
        uint32_t ld_src = (uint32_t)B;
        bool src_perm = false;  // caller-visible buffers are in the caller's cell order, the scratch in plan order
        for (int32_t it = 0; it < iters; ++it) {
            T *dst;
            uint32_t ld_dst;
            bool dst_perm = false;
            const int64_t nxt = (int64_t)it + 1;
            if (d_history && nxt % history_stride == 0 && nxt / history_stride < slots) {
                dst = d_history + (size_t)(nxt / history_stride) * (size_t)n * B + c0;
                ld_dst = (uint32_t)B;
            } else if (nxt == iters && d_Xfinal) {
                dst = d_Xfinal + c0;
                ld_dst = (uint32_t)B;
            } else {
                dst = (src == bufA) ? bufB : bufA;
                ld_dst = (uint32_t)nc;
                dst_perm = reord;
            }
            double *acc = d_eps ? eps_acc + 2 * (size_t)it * (size_t)nc * kEpsCopies : nullptr;
            GatherOperand g = ident;
            if (reord) {
                g.indptr = plan->d_r_indptr;
                g.pairs = src_perm ? plan->d_r_pairs_pp : plan->d_r_pairs_op;
                g.dest_ids = plan->d_order;
                g.out_orig = dst_perm ? 0 : 1;
            }
            int32_t *cnt = nullptr;
            if (sm_affine && src_perm) {  // permuted -> permuted / permuted -> original updates
                const int64_t slot = launch_no % kCounterSlots;
                if (slot == 0)
                    CY2P_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * (size_t)kCounterSlots * kMaxSms, st));
                cnt = counters + slot * kMaxSms;
                ++launch_no;
            }
            constexpr int SV = kSegBytes / 32 / (int)sizeof(T);  // elements per lane of the staged kernel
            if (staged && src_perm && dst_perm && nc % (32 * SV) == 0 && ld_src == ld_dst) {
                StagedArgs<T> sa;
                sa.blk_row = plan->d_blk_row;
                sa.blk_uptr = plan->d_blk_uptr;
                sa.blk_usrc = plan->d_blk_usrc;
                sa.r_indptr = plan->d_r_indptr;
                sa.lpairs = plan->d_r_pairs_lp;
                sa.dest_ids = plan->d_order;
                sa.Xin = src;
                sa.Xout = dst;
                sa.ld = ld_src;
                sa.ncols = nc;
                sa.nslabs = nc / (32 * SV);
                sa.slabs_per_cta = staged_spc;
                sa.ucap = plan->blk_ucap;
                sa.stationary = d_stationary;
                sa.pi_rows = pi_rows;
                sa.eps_acc = acc;
                const dim3 grid((unsigned)plan->nblk, (unsigned)((sa.nslabs + staged_spc - 1) / staged_spc));
                count_launch();
                if (acc)
                    spmm_staged<T, SV, true><<<grid, kStagedWarps * 32, staged_smem, st>>>(sa);
                else
                    spmm_staged<T, SV, false><<<grid, kStagedWarps * 32, staged_smem, st>>>(sa);
            } else if (wide && (nc % VW == 0)) {
                // CY2P_GATHER_VEC = 2 / 1 (experiment): a warp takes 256- / 128-byte row segments instead of 512-byte
                // ones, so the SM's L1 holds 2x / 4x as many segments (LRU model of tools/reorder_lab.cpp: hit rate
                // 22% -> 36% -> 50%) at the price of more (src,val) loads per operand byte
                const int vec_env = env_int("CY2P_GATHER_VEC", 0);
                if (sizeof(T) == 4 && vec_env == 2)
                    launch_gather<T, 2>(g, 0, n, src, ld_src, dst, ld_dst, nc, d_stationary, acc, st, cnt, plan->num_sms,
                                        pi_rows);
                else if (vec_env == 1)
                    launch_gather<T, 1>(g, 0, n, src, ld_src, dst, ld_dst, nc, d_stationary, acc, st, cnt, plan->num_sms,
                                        pi_rows);
                else
                    launch_gather<T, VW>(g, 0, n, src, ld_src, dst, ld_dst, nc, d_stationary, acc, st, cnt,
                                         plan->num_sms, pi_rows);
            } else
                launch_gather<T, 1>(g, 0, n, src, ld_src, dst, ld_dst, nc, d_stationary, acc, st, cnt, plan->num_sms,
                                    pi_rows);
            src = dst;
            ld_src = ld_dst;
            src_perm = dst_perm;
        }
        if (d_Xfinal && src != d_Xfinal + c0)  // the last iterate landed in a history slot / scratch
            CY2P_CUDA(cudaMemcpy2DAsync(d_Xfinal + c0, sizeof(T) * (size_t)B, src, sizeof(T) * (size_t)ld_src,
                                        sizeof(T) * (size_t)nc, (size_t)n, cudaMemcpyDeviceToDevice, st));
        if (d_eps) {
            const int64_t cnt = (int64_t)iters * nc;
            count_launch();
            eps_finalize_tile<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(iters, kEpsCopies, nc, c0, B, eps_acc, vv,
                                                                             d_eps);
        }
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

template <typename T>
static int32_t propagate_rows_impl(cy2p_plan *plan, const T *d_X, T *d_Y, int32_t B, int32_t row_begin,
                                   int32_t row_end, void *stream) {
    constexpr int VW = Wide<T>::VEC;
    CY2P_REQUIRE(plan && d_X && d_Y, "null pointer");
    CY2P_REQUIRE(B >= 1 && row_begin >= 0 && row_end <= plan->n && row_begin <= row_end, "bad row range / B");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    if (row_begin == row_end) return CY2P_OK;
    const GatherOperand ident = {plan->d_t_indptr, plan->d_t_pairs, nullptr, 0};
    if (B == 1) {
        const int rows_per_cta = 256 / 8;
        count_launch();
        spmv_rows<T><<<(unsigned)((row_end - row_begin + rows_per_cta - 1) / rows_per_cta), 256, 0, st>>>(
            row_begin, row_end, plan->d_t_indptr, plan->d_t_pairs, d_X, d_Y, nullptr, 0, nullptr);
    } else if (B <= 16 && env_int("CY2P_ROWS_NARROW", 1)) {
        const int rows_per_cta = 16;  // 8 warps x 2 rows
        count_launch();
        spmm_rows_narrow<T><<<(unsigned)((row_end - row_begin + rows_per_cta - 1) / rows_per_cta), 256, 0, st>>>(
            row_begin, row_end, plan->d_t_indptr, plan->d_t_pairs, d_X, B, d_Y);
    } else if (B % VW == 0 && aligned16(d_X) && aligned16(d_Y))
        launch_gather<T, VW>(ident, row_begin, row_end, d_X, B, d_Y, B, B, nullptr, nullptr, st);
    else
        launch_gather<T, 1>(ident, row_begin, row_end, d_X, B, d_Y, B, B, nullptr, nullptr, st);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

// Row range of one update entirely in the plan's locality order (X and Y permuted): the row-partitioned driver
// permutes the start distribution once, iterates here, and un-permutes the result once.
template <typename T>
static int32_t propagate_rows_ordered_impl(cy2p_plan *plan, const T *d_Xp, T *d_Yp, int32_t B, int32_t q_begin,
                                           int32_t q_end, void *stream, void *const *d_peer_out = nullptr,
                                           int32_t npeers = 0, const uint32_t *d_peer_mask = nullptr) {
    constexpr int VW = Wide<T>::VEC;
    CY2P_REQUIRE(plan && d_Xp && (d_Yp || d_peer_out), "null pointer");
    CY2P_REQUIRE(!d_peer_out || (npeers >= 1 && npeers <= 64), "1 <= npeers <= 64 required with peer outputs");
    CY2P_REQUIRE(!d_peer_mask || (d_peer_out && npeers <= 32), "a peer mask needs peer outputs and at most 32 peers");
    CY2P_REQUIRE(B >= 1 && q_begin >= 0 && q_end <= plan->n && q_begin <= q_end, "bad row range / B");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    if (q_begin == q_end) return CY2P_OK;
    const bool reord = plan->reorder_state == 1;
    const int32_t *ip = reord ? plan->d_r_indptr : plan->d_t_indptr;
    const Pair *pr = reord ? plan->d_r_pairs_pp : plan->d_t_pairs;
    if (B == 1) {
        const int rows_per_cta = 256 / 8;
        count_launch();
        spmv_rows<T><<<(unsigned)((q_end - q_begin + rows_per_cta - 1) / rows_per_cta), 256, 0, st>>>(
            q_begin, q_end, ip, pr, d_Xp, d_Yp, (T *const *)d_peer_out, npeers, d_peer_mask);
    } else {
        GatherOperand g = {ip, pr, nullptr, 0};  // rows stay in plan order on both sides
        g.peer_out = d_peer_out;
        g.npeers = npeers;
        g.peer_mask = d_peer_mask;
        if (B % VW == 0 && aligned16(d_Xp) && aligned16(d_Yp))
            launch_gather<T, VW>(g, q_begin, q_end, d_Xp, B, d_Yp, B, B, nullptr, nullptr, st);
        else
            launch_gather<T, 1>(g, q_begin, q_end, d_Xp, B, d_Yp, B, B, nullptr, nullptr, st);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

template <typename T>
static int32_t rows_persist_impl(cy2p_plan *plan, void *const *d_peer_buf0, void *const *d_peer_buf1, int32_t world,
                                 int32_t rank, const uint32_t *d_row_peers, int32_t q_begin, int32_t q_end, int32_t iters,
                                 void *const *d_peer_flags, int32_t *d_my_flags, int32_t *d_epoch_counter,
                                 int32_t *d_error, void *stream) {
    CY2P_REQUIRE(plan && d_peer_buf0 && d_peer_buf1 && d_peer_flags && d_my_flags && d_epoch_counter && d_error, "null pointer");
    CY2P_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "1 <= world <= 32 and 0 <= rank < world required");
    CY2P_REQUIRE(q_begin >= 0 && q_end <= plan->n && q_begin <= q_end && iters >= 0, "bad row range / iters");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    if (iters == 0) return CY2P_OK;
    const bool reord = plan->reorder_state == 1;
    const int32_t *ip = reord ? plan->d_r_indptr : plan->d_t_indptr;
    const Pair *pr = reord ? plan->d_r_pairs_pp : plan->d_t_pairs;
    int per_sm = 0;
    CY2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_rows_persist<T>, 256, 0));
    int grid = plan->num_sms * (per_sm > 4 ? 4 : per_sm);
    const int need = (q_end - q_begin + 31) / 32;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    void *args[] = {&q_begin, &q_end, &ip, &pr, &d_peer_buf0, &d_peer_buf1, &d_row_peers, &rank, &world, &iters,
                    &d_peer_flags, &d_my_flags, &d_epoch_counter, &d_error};
    count_launch();
    CY2P_CUDA(cudaLaunchCooperativeKernel((void *)spmv_rows_persist<T>, dim3(grid), dim3(256), args, 0, st));
    return CY2P_OK;
}

__global__ void iota_kernel(int32_t n, int32_t *out) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

}  // namespace cy2p

using namespace cy2p;

extern "C" {

int32_t cy2p_plan_order(cy2p_plan *plan, int32_t *d_order, void *stream) {
    CY2P_REQUIRE(plan && d_order, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    const int32_t rc = ensure_reordered(plan, st);
    if (rc != CY2P_OK) return rc;
    if (plan->reorder_state == 1) {
        CY2P_CUDA(cudaMemcpyAsync(d_order, plan->d_order, sizeof(int32_t) * (size_t)plan->n, cudaMemcpyDeviceToDevice, st));
    } else {
        count_launch();
        iota_kernel<<<(plan->n + 255) / 256, 256, 0, st>>>(plan->n, d_order);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_propagate_rows_ordered(cy2p_plan *plan, const float *d_Xp, float *d_Yp, int32_t B, int32_t q_begin,
                                    int32_t q_end, void *stream) {
    return propagate_rows_ordered_impl<float>(plan, d_Xp, d_Yp, B, q_begin, q_end, stream);
}

int32_t cy2p_propagate_rows_ordered_f64(cy2p_plan *plan, const double *d_Xp, double *d_Yp, int32_t B, int32_t q_begin,
                                        int32_t q_end, void *stream) {
    return propagate_rows_ordered_impl<double>(plan, d_Xp, d_Yp, B, q_begin, q_end, stream);
}

int32_t cy2p_propagate_rows_allgather(cy2p_plan *plan, const float *d_Xp, void *const *d_peer_out, int32_t npeers,
                                      int32_t B, int32_t q_begin, int32_t q_end, void *stream) {
    return propagate_rows_ordered_impl<float>(plan, d_Xp, nullptr, B, q_begin, q_end, stream, d_peer_out, npeers);
}

int32_t cy2p_propagate_rows_allgather_f64(cy2p_plan *plan, const double *d_Xp, void *const *d_peer_out,
                                          int32_t npeers, int32_t B, int32_t q_begin, int32_t q_end, void *stream) {
    return propagate_rows_ordered_impl<double>(plan, d_Xp, nullptr, B, q_begin, q_end, stream, d_peer_out, npeers);
}

int32_t cy2p_propagate_rows_push(cy2p_plan *plan, const float *d_Xp, void *const *d_peer_out, int32_t npeers,
                                 const uint32_t *d_row_peers, int32_t B, int32_t q_begin, int32_t q_end, void *stream) {
    return propagate_rows_ordered_impl<float>(plan, d_Xp, nullptr, B, q_begin, q_end, stream, d_peer_out, npeers,
                                              d_row_peers);
}

int32_t cy2p_propagate_rows_push_f64(cy2p_plan *plan, const double *d_Xp, void *const *d_peer_out, int32_t npeers,
                                     const uint32_t *d_row_peers, int32_t B, int32_t q_begin, int32_t q_end,
                                     void *stream) {
    return propagate_rows_ordered_impl<double>(plan, d_Xp, nullptr, B, q_begin, q_end, stream, d_peer_out, npeers,
                                               d_row_peers);
}

int32_t cy2p_propagate_rows_persist(cy2p_plan *plan, void *const *d_peer_buf0, void *const *d_peer_buf1, int32_t world,
                                    int32_t rank, const uint32_t *d_row_peers, int32_t q_begin, int32_t q_end,
                                    int32_t iters, void *const *d_peer_flags, int32_t *d_my_flags,
                                    int32_t *d_epoch_counter, int32_t *d_error, void *stream) {
    return rows_persist_impl<float>(plan, d_peer_buf0, d_peer_buf1, world, rank, d_row_peers, q_begin, q_end, iters,
                                    d_peer_flags, d_my_flags, d_epoch_counter, d_error, stream);
}

int32_t cy2p_propagate_rows_persist_f64(cy2p_plan *plan, void *const *d_peer_buf0, void *const *d_peer_buf1,
                                        int32_t world, int32_t rank, const uint32_t *d_row_peers, int32_t q_begin,
                                        int32_t q_end, int32_t iters, void *const *d_peer_flags, int32_t *d_my_flags,
                                        int32_t *d_epoch_counter, int32_t *d_error, void *stream) {
    return rows_persist_impl<double>(plan, d_peer_buf0, d_peer_buf1, world, rank, d_row_peers, q_begin, q_end, iters,
                                     d_peer_flags, d_my_flags, d_epoch_counter, d_error, stream);
}

int32_t cy2p_enable_peer_access(int32_t device, int32_t peer) {
    CY2P_CUDA(cudaSetDevice(device));
    if (device == peer) return CY2P_OK;
    int can = 0;
    CY2P_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) {
        set_error("device %d cannot access peer %d", device, peer);
        return CY2P_ERR_CUDA;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer, cudaGetErrorString(e));
        return CY2P_ERR_CUDA;
    }
    cudaGetLastError();
    return CY2P_OK;
}

size_t cy2p_propagate_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters) {
    if (!plan || B < 1 || iters < 0) return 0;
    return ws_layout(plan, B, iters, sizeof(float)).total;
}
size_t cy2p_propagate_f64_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters) {
    if (!plan || B < 1 || iters < 0) return 0;
    return ws_layout(plan, B, iters, sizeof(double)).total;
}

int32_t cy2p_propagate_column_tile(const cy2p_plan *plan, int32_t B) {
    if (!plan || B < 1) return 0;
    return choose_tile(plan, B, sizeof(float));
}

int32_t cy2p_propagate(cy2p_plan *plan, const float *d_X0, int32_t B, int32_t iters, const double *d_stationary,
                       float *d_Xfinal, float *d_history, int32_t history_stride, double *d_eps, void *d_ws,
                       size_t ws_bytes, void *stream) {
    return propagate_impl<float>(plan, d_X0, B, iters, d_stationary, d_Xfinal, d_history, history_stride, d_eps, d_ws,
                                 ws_bytes, stream);
}

int32_t cy2p_propagate_f64(cy2p_plan *plan, const double *d_X0, int32_t B, int32_t iters, const double *d_stationary,
                           double *d_Xfinal, double *d_history, int32_t history_stride, double *d_eps, void *d_ws,
                           size_t ws_bytes, void *stream) {
    return propagate_impl<double>(plan, d_X0, B, iters, d_stationary, d_Xfinal, d_history, history_stride, d_eps,
                                  d_ws, ws_bytes, stream);
}

int32_t cy2p_propagate_rows(cy2p_plan *plan, const float *d_X, float *d_Y, int32_t B, int32_t row_begin,
                            int32_t row_end, void *stream) {
    return propagate_rows_impl<float>(plan, d_X, d_Y, B, row_begin, row_end, stream);
}

int32_t cy2p_propagate_rows_f64(cy2p_plan *plan, const double *d_X, double *d_Y, int32_t B, int32_t row_begin,
                                int32_t row_end, void *stream) {
    return propagate_rows_impl<double>(plan, d_X, d_Y, B, row_begin, row_end, stream);
}

}  // extern "C"
