// K1x — batched MSM propagation  X <- X . T  with an EXTENDED-PRECISION iterate (reference sampling.py:33,38-45).
//
// Why: the reference multiplies a float64 iterate by the float32 T in float64 (scipy upcasts). The float32-rounded
// rows of T do not sum to exactly 1, so on absorbing regions the reference's iterate grows by ~1e-8 per step; a
// float32 iterate cannot represent that growth (it rounds away every step) and drifts linearly from the reference:
// relative L1 3.9e-5 after 2,000 iterations, and the float32 FMA chain of the row sum does the same even when the
// iterate is stored wider (tools/precision_lab.py). The north_star's bar is 1e-6. What holds it is (i) products and
// row sums in float64 and (ii) an iterate with >= 28 mantissa bits.
//
// Format "f48" / "f40" (truncated float64): an element is the top 32 bits of its float64 pattern (sign, 11-bit
// exponent, 20 mantissa bits: the HEAD) plus the next 16 / 8 mantissa bits (the TAIL), rounded to nearest: 36 / 28
// mantissa bits in 6 / 5 bytes instead of 8. The head IS the high word of a double and the tail, shifted to the top of
// the low word, is the rest, so a gathered operand becomes a DFMA operand with ONE integer instruction per element and
// no conversion (F2F runs at a quarter of the DFMA rate). The kernel is bound by the bytes it moves from L2 into the
// SMs (profiles/README_r02.md), so 6 bytes instead of 8 is 25% less of the binding traffic.
//
// Layout of an iterate tile in the work space: row r, column slab s (32*CPL columns) is one contiguous record
//   [ heads: 32*CPL x uint32 | tails: 32*CPL x uint16 (uint8) ]      768 (640) bytes at CPL = 4
// so a warp's gather of one source is one LDG.128 + one LDG.64 (LDG.32) per lane from one contiguous span.
//
// eps (sampling.py:44-45) is deterministic here: per-thread float64 partial sums over the thread's rows in a fixed
// order, then 64-bit FIXED-POINT integer atomics (integer addition is associative, so the order in which warps and
// CTAs arrive cannot change a bit of the result) — see eps_scale_kernel for the scale.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace cy2p {

namespace cgx = cooperative_groups;

constexpr int kXEpsCopies = 4;  // spread of the global integer accumulators (same-address atomics serialise in L2)

template <int LOB, int CPL>
struct XFmt {
    static constexpr int kHiBytes = 32 * CPL * 4;
    static constexpr int kLoBytes = 32 * CPL * LOB / 8;
    static constexpr int kRec = kHiBytes + kLoBytes;  // bytes of one (row, slab) record
    static constexpr int kCols = 32 * CPL;
};

struct XArgs {
    const int32_t *indptr;
    const PairD *pairs;
    const int32_t *dest_ids;  // original cell id of walked row r (null: identity)
    int32_t out_orig;         // float outputs: write row dest_ids[r] instead of r
    const unsigned char *xin;  // f48 tile
    uint32_t pitch_in;         // bytes per row of the input tile
    const float *fin;          // float32 input (first update), leading dimension ld_in
    uint32_t ld_in;
    unsigned char *xout;
    uint32_t pitch_out;
    float *fout;   // float32 output (last update), may be null
    double *dout;  // float64 output (last update), may be null
    uint32_t ld_out;
    int32_t ncols;
    int32_t wide_io;  // float input/output rows are 16-byte aligned and ncols % 4 == 0
    const double *pi_rows;
    unsigned long long *eps_acc;  // [kXEpsCopies][ncols][2] fixed-point sums of this update
    const double *eps_scale;      // device scalar
    // Plain-grid launches: the CTA of (row block rb, slab) parks its 2 x CPL x 32 fixed-point sums at
    // eps_part[(rb * nslabs + slab) * 2 * CPL * 32] with coalesced stores and x_eps_reduce adds them afterwards. The
    // 512 64-bit reductions a CTA used to issue itself kept its registers resident for another ~3 us (of ~64): 5% of
    // the update (profiles/k1x_epsdbg_r02.jsonl). Null: reduce straight into eps_acc (persistent / multi variants).
    unsigned long long *eps_part;
    int32_t dbg;                  // CY2PX_EPS_DBG (measurement only): 1 no global flush, 2 no epilogue, 4 no prologue + epilogue, 8 no pi load
};

__device__ __forceinline__ double mk_double(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// base + row * pitch as ONE IMAD.WIDE.U32 with a 64-bit addend (left to itself the compiler adds the lane offset in
// the multiply-add and the uniform base pointer with two more adds)
__device__ __forceinline__ const unsigned char *row_ptr(const unsigned char *base, uint32_t row, uint32_t pitch) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(row), "r"(pitch), "l"((uint64_t)base));
    return reinterpret_cast<const unsigned char *>(r);
}

// ---- operand loads: CPL doubles of one source row for this lane ---------------------------------------------------
// Record of (row, slab): heads as CPL/4 planes of 32 x uint4 (CPL = 2: one plane of 32 x uint2), then the tails,
// CPL*LOB/8 bytes per lane. Lane `lane` owns columns (slab*32 + lane)*CPL + k, k < CPL.
template <int LOB, int CPL>
struct XLoad {
    static constexpr int kLoWords = (CPL * LOB + 31) / 32;
    uint32_t h[CPL];
    uint32_t l[kLoWords];
    // hi: record + lane * (CPL >= 4 ? 16 : 8);  lo: record + kHiBytes + lane * (CPL*LOB/8)
    __device__ __forceinline__ void issue(const unsigned char *hi, const unsigned char *lo) {
        if constexpr (CPL == 2) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(hi));
            h[0] = v.x, h[1] = v.y;
        } else {
#pragma unroll
            for (int j = 0; j < CPL / 4; ++j) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(hi + j * 512));
                h[4 * j] = v.x, h[4 * j + 1] = v.y, h[4 * j + 2] = v.z, h[4 * j + 3] = v.w;
            }
        }
        if constexpr (CPL * LOB == 16) {
            l[0] = (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(lo));
        } else if constexpr (CPL * LOB == 32) {
            l[0] = __ldg(reinterpret_cast<const uint32_t *>(lo));
        } else if constexpr (CPL * LOB == 64) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(lo));
            l[0] = v.x, l[1] = v.y;
        } else {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(lo));
            l[0] = v.x, l[1] = v.y, l[2] = v.z, l[3] = v.w;
        }
    }
    // the tail of element k moved to the top of a float64's low word: one integer instruction
    __device__ __forceinline__ uint32_t low_word(int k) const {
        if constexpr (LOB == 16) {
            return (k & 1) ? (l[k / 2] & 0xffff0000u) : (l[k / 2] << 16);
        } else {
            const uint32_t w = l[k / 4];
            switch (k & 3) {
                case 0: return w << 24;
                case 1: return __byte_perm(w, 0u, 0x1444);
                case 2: return __byte_perm(w, 0u, 0x2444);
                default: return w & 0xff000000u;
            }
        }
    }
    __device__ __forceinline__ void fma(double v, double *acc) const {
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = ::fma(v, mk_double(low_word(k), h[k]), acc[k]);
    }
};

// float32 input row (first update): exact widening, CPL values per lane
template <int CPL>
struct FLoad {
    float x[CPL];
    __device__ __forceinline__ void issue(const float *p, int nvalid, bool wide) {
        if (wide) {
            if constexpr (CPL == 2) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
                x[0] = v.x, x[1] = v.y;
            } else {
#pragma unroll
                for (int j = 0; j < CPL / 4; ++j) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(p + 4 * j));
                    x[4 * j] = v.x, x[4 * j + 1] = v.y, x[4 * j + 2] = v.z, x[4 * j + 3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < CPL; ++k) x[k] = k < nvalid ? __ldg(p + k) : 0.f;
        }
    }
    __device__ __forceinline__ void fma(double v, double *acc) const {
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = ::fma(v, (double)x[k], acc[k]);
    }
};

// round a float64 to the stored format: keep LOB bits of the low word, nearest (ties away from zero)
template <int LOB>
__device__ __forceinline__ void x_round(double y, uint32_t &hi, uint32_t &tail, double &yr) {
    constexpr int drop = 32 - LOB;
    long long b = __double_as_longlong(y);
    b += (1LL << (drop - 1));
    hi = (uint32_t)((unsigned long long)b >> 32);
    const uint32_t low = (uint32_t)b;
    tail = low >> drop;
    yr = mk_double(low & ~((1u << drop) - 1u), hi);
}

template <int LOB, int CPL, int WARPS, int U, bool IN_F32, bool OUT_F32, bool EPS>
__device__ __forceinline__ void gather_block_x(const XArgs &a, int32_t row0, int32_t row1, int slab, int copy_sel,
                                               unsigned long long (*s_acc)[CPL][32], unsigned *s_done,
                                               int64_t part_slot = -1) {
    using F = XFmt<LOB, CPL>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (slab * 32 + lane) * CPL;
    const bool active = col < a.ncols;
    const int nvalid = min(CPL, a.ncols - col);
    double uv[CPL], uu[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) uv[k] = uu[k] = 0.0;
    // EPS is compiled in for every launch (the instantiation WITHOUT the partial sums gets a worse schedule from ptxas:
    // 3.69 vs 2.70 ms per update, profiles/k1x_sweep5_noeps_r02.jsonl); a call without a stationary vector passes
    // eps_acc == null and only skips the epilogue
    const bool want_eps = EPS && a.eps_acc != nullptr;
    if (want_eps && !(a.dbg & 4)) {
        for (int t = threadIdx.x; t < 2 * CPL * 32; t += WARPS * 32) (&s_acc[0][0][0])[t] = 0ull;
        if (threadIdx.x == 0) *s_done = 0u;
        __syncthreads();
    }
    if (active) {
        const PairD *__restrict__ pairs = a.pairs;
        const unsigned char *xhi = a.xin + (size_t)slab * F::kRec + lane * (CPL >= 4 ? 16 : 8);
        const unsigned char *xlo = a.xin + (size_t)slab * F::kRec + F::kHiBytes + lane * (CPL * LOB / 8);
        const float *fin = a.fin + col;
        const bool wide = a.wide_io != 0;
        const uint32_t pitch_in = a.pitch_in;
        // Software pipeline over the warp's stream of entries: the (source, value) entries of group g+1 — the next U
        // entries of this row or the first U of the warp's NEXT row, whose row pointers were fetched one row ahead —
        // are requested before the gathers of group g are consumed, so an update costs one exposed L2 round trip per
        // group instead of two dependent ones (entries, then rows); ncu of the unpipelined loop: 74% of the issue
        // latency was the long scoreboard at 31% occupancy, nothing saturated (profiles/README_r02.md).
        auto load_entries = [&](uint4 *q, int32_t p, int32_t end) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                q[u] = (p + u < end) ? __ldg(reinterpret_cast<const uint4 *>(pairs + p + u)) : make_uint4(0u, 0u, 0u, 0u);
        };
        int32_t r = row0 + warp;
        int32_t p = 0, end = 0;
        if (r < row1) p = __ldg(a.indptr + r), end = __ldg(a.indptr + r + 1);
        uint4 q[U];
        load_entries(q, p, end);
        for (; r < row1; r += WARPS) {
            const int32_t rn = r + WARPS;
            int32_t pn = 0, endn = 0;
            if (rn < row1) pn = __ldg(a.indptr + rn), endn = __ldg(a.indptr + rn + 1);
            const double pi = (want_eps && !(a.dbg & 8)) ? __ldg(a.pi_rows + r) : 1.0;
            const int32_t orow = (OUT_F32 && a.out_orig) ? __ldg(a.dest_ids + r) : r;
            double acc[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[k] = 0.0;
            do {  // an empty row runs one group of padding entries (value 0, loads skipped)
                uint4 qn[U];
                if (IN_F32) {
                    FLoad<CPL> x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) x[u].x[k] = 0.f;
                        if (p + u < end) x[u].issue(fin + (uint64_t)q[u].x * a.ld_in, nvalid, wide);
                    }
                    p += U;
                    if (p < end)
                        load_entries(qn, p, end);
                    else
                        load_entries(qn, pn, endn);
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u].fma(mk_double(q[u].z, q[u].w), acc);
                } else {
                    XLoad<LOB, CPL> x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) x[u].h[k] = 0u;
#pragma unroll
                        for (int k = 0; k < XLoad<LOB, CPL>::kLoWords; ++k) x[u].l[k] = 0u;
                        if (p + u < end) x[u].issue(row_ptr(xhi, q[u].x, pitch_in), row_ptr(xlo, q[u].x, pitch_in));
                    }
                    p += U;
                    if (p < end)
                        load_entries(qn, p, end);
                    else
                        load_entries(qn, pn, endn);
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u].fma(mk_double(q[u].z, q[u].w), acc);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) q[u] = qn[u];
            } while (p < end);
            p = pn, end = endn;
            // ---- store the row in the stored format (or as float32 / float64 on the last update)
            uint32_t hi[CPL], tl[CPL];
            double yr[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) x_round<LOB>(acc[k], hi[k], tl[k], yr[k]);
            if (OUT_F32) {
                const uint64_t off = (uint64_t)(uint32_t)orow * a.ld_out + col;
                if (a.fout) {
                    if (wide) {
                        if constexpr (CPL == 2) {
                            *reinterpret_cast<float2 *>(a.fout + off) = make_float2((float)acc[0], (float)acc[1]);
                        } else {
#pragma unroll
                            for (int j = 0; j < CPL / 4; ++j)
                                *reinterpret_cast<float4 *>(a.fout + off + 4 * j) = make_float4(
                                    (float)acc[4 * j], (float)acc[4 * j + 1], (float)acc[4 * j + 2], (float)acc[4 * j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < CPL; ++k)
                            if (k < nvalid) a.fout[off + k] = (float)acc[k];
                    }
                }
                if (a.dout) {
#pragma unroll
                    for (int k = 0; k < CPL; ++k)
                        if (k < nvalid) a.dout[off + k] = acc[k];
                }
            } else {
                unsigned char *rec = a.xout + (uint64_t)(uint32_t)r * a.pitch_out + (size_t)slab * F::kRec;
                unsigned char *ohi = rec + lane * (CPL >= 4 ? 16 : 8), *olo = rec + F::kHiBytes + lane * (CPL * LOB / 8);
                if constexpr (CPL == 2) {
                    *reinterpret_cast<uint2 *>(ohi) = make_uint2(hi[0], hi[1]);
                } else {
#pragma unroll
                    for (int j = 0; j < CPL / 4; ++j)
                        *reinterpret_cast<uint4 *>(ohi + j * 512) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                }
                constexpr int LW = (CPL * LOB + 31) / 32;
                uint32_t lw[LW];
#pragma unroll
                for (int j = 0; j < LW; ++j) lw[j] = 0u;
#pragma unroll
                for (int k = 0; k < CPL; ++k) lw[k * LOB / 32] |= tl[k] << ((k * LOB) % 32);
                if constexpr (CPL * LOB == 16)
                    *reinterpret_cast<unsigned short *>(olo) = (unsigned short)lw[0];
                else if constexpr (CPL * LOB == 32)
                    *reinterpret_cast<uint32_t *>(olo) = lw[0];
                else if constexpr (CPL * LOB == 64)
                    *reinterpret_cast<uint2 *>(olo) = make_uint2(lw[0], lw[1]);
                else
                    *reinterpret_cast<uint4 *>(olo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            if (EPS) {
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const double y = (k < nvalid) ? yr[k] : 0.0;
                    uv[k] = ::fma(y, pi, uv[k]);
                    uu[k] = ::fma(y, y, uu[k]);
                }
            }
        }
    }
    if (want_eps && !(a.dbg & 6)) {
        const double sc = __ldg(a.eps_scale);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            atomicAdd(&s_acc[0][k][lane], (unsigned long long)__double2ll_rn(uv[k] * sc));
            atomicAdd(&s_acc[1][k][lane], (unsigned long long)__double2ll_rn(uu[k] * sc));
        }
        __syncwarp();
        __threadfence_block();
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(s_done, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == WARPS - 1 && !(a.dbg & 1)) {  // every warp's partials are in: one integer atomic per (column, which) and CTA
            __threadfence_block();
            if (part_slot >= 0 && a.eps_part != nullptr) {
                uint4 *dst = reinterpret_cast<uint4 *>(a.eps_part + part_slot * (2 * CPL * 32));
                const volatile unsigned long long *src = &s_acc[0][0][0];
#pragma unroll
                for (int t = lane; t < CPL * 32; t += 32) {
                    const unsigned long long v0 = src[2 * t], v1 = src[2 * t + 1];
                    dst[t] = make_uint4((uint32_t)v0, (uint32_t)(v0 >> 32), (uint32_t)v1, (uint32_t)(v1 >> 32));
                }
            } else
            for (int t = lane; t < 2 * CPL * 32; t += 32) {
                const int which = t / (CPL * 32), k = (t / 32) % CPL, ln = t & 31;
                const int gcol = (slab * 32 + ln) * CPL + k;
                if (gcol < a.ncols)
                    atomicAdd(&a.eps_acc[((int64_t)copy_sel * a.ncols + gcol) * 2 + which],
                              *(volatile unsigned long long *)&s_acc[which][k][ln]);
            }
        }
    }
}

// MINB = resident CTAs per SM the register budget is compiled for, U = row gathers in flight per warp
template <int LOB, int CPL, int WARPS, int MINB, int U, bool IN_F32, bool OUT_F32, bool EPS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_gather_x(int32_t row_begin, int32_t row_end, int32_t rows_per_cta, XArgs a) {
    __shared__ unsigned long long s_acc[2][CPL][32];
    __shared__ unsigned s_done;
    const int32_t row0 = row_begin + blockIdx.x * rows_per_cta;
    gather_block_x<LOB, CPL, WARPS, U, IN_F32, OUT_F32, EPS>(a, row0, min(row0 + rows_per_cta, row_end), blockIdx.y,
                                                             blockIdx.x % kXEpsCopies, s_acc, &s_done,
                                                             (int64_t)blockIdx.x * gridDim.y + blockIdx.y);
}

// Adds the parked per-CTA sums of one update into eps_acc: thread = one (slab, which, column-in-lane, lane) entry and
// one of gridDim.y interleaved shares of the row blocks; integer adds, so the order does not matter.
template <int CPL>
__global__ void __launch_bounds__(256)
x_eps_reduce(int32_t nrb, int32_t nslabs, int32_t ncols, const unsigned long long *__restrict__ part,
             unsigned long long *__restrict__ eps_acc) {
    constexpr int PER = 2 * CPL * 32;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nslabs * PER) return;
    const int slab = o / PER, idx = o - slab * PER;
    const int which = idx / (CPL * 32), k = (idx / 32) % CPL, ln = idx & 31;
    const int gcol = (slab * 32 + ln) * CPL + k;
    if (gcol >= ncols) return;
    const int QS = gridDim.y;
    const unsigned long long *p = part + (size_t)slab * PER + idx;
    const size_t stride = (size_t)nslabs * PER;
    unsigned long long acc = 0ull;
    for (int rb = blockIdx.y; rb < nrb; rb += QS * 8) {
        unsigned long long v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = rb + u * QS < nrb ? __ldcg(p + (size_t)(rb + u * QS) * stride) : 0ull;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    atomicAdd(&eps_acc[((int64_t)(blockIdx.y % kXEpsCopies) * ncols + gcol) * 2 + which], acc);
}

// Persistent variant: num_SMs x MINB resident CTAs draw (column slab, row group) items from one counter in launch
// order — slab-major, row groups ascending, so the wave of CTAs still sweeps one contiguous band of the ordered rows
// (what keeps the gathers in L2) — and a CTA's warps stay on ADJACENT rows for the whole launch: no CTA tail, no block
// scheduling between row groups. counter: one zeroed int per launch.
template <int LOB, int CPL, int WARPS, int MINB, int U, bool EPS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_gather_x_persist(int32_t n_rows, int32_t rows_per_item, int32_t nslabs, int32_t *__restrict__ counter, XArgs a) {
    __shared__ unsigned long long s_acc[2][CPL][32];
    __shared__ unsigned s_done;
    __shared__ int32_t s_item;
    const int32_t nrg = (n_rows + rows_per_item - 1) / rows_per_item;
    const int32_t total = nrg * nslabs;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
        __syncthreads();
        const int32_t item = s_item;
        if (item >= total) break;
        const int32_t slab = item / nrg, rg = item % nrg;
        const int32_t row0 = rg * rows_per_item;
        gather_block_x<LOB, CPL, WARPS, U, false, false, EPS>(a, row0, min(row0 + rows_per_item, n_rows), slab,
                                                               rg % kXEpsCopies, s_acc, &s_done);
        __syncthreads();  // s_item and the eps accumulators are reused by the next item
    }
}

// Persistent MULTI-ITERATION variant (the shape the north_star names: one launch carries a column tile through many
// updates, the ping-pong iterate sized to stay in L2): a cooperative grid of num_SMs x MINB CTAs, a work counter per
// update, grid.sync() between updates. Opt-in (CY2PX_MULTI=1): measured against the default in
// profiles/k1x_sweep_multi_r02.jsonl — the update is bound by the L2 -> SM path either way, so keeping HBM out of it
// (DRAM traffic -> 0) does not shorten it, and the narrow tile an L2-resident iterate needs (128 columns at 50k cells)
// leaves the grid barrier and the tail of every update exposed.
template <int LOB, int CPL, int WARPS, int MINB, int U, bool EPS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_gather_x_multi(int32_t n_rows, int32_t rows_per_item, int32_t nslabs, int32_t n_updates, int32_t *__restrict__ counters,
                    unsigned char *bufA, unsigned char *bufB, int32_t first_src_is_A, size_t eps_stride, XArgs a) {
    cgx::grid_group grid = cgx::this_grid();
    __shared__ unsigned long long s_acc[2][CPL][32];
    __shared__ unsigned s_done;
    __shared__ int32_t s_item;
    const int32_t nrg = (n_rows + rows_per_item - 1) / rows_per_item;
    const int32_t total = nrg * nslabs;
    for (int32_t it = 0; it < n_updates; ++it) {
        const bool src_is_A = ((it & 1) == 0) == (first_src_is_A != 0);
        XArgs b = a;
        b.xin = src_is_A ? bufA : bufB;
        b.xout = src_is_A ? bufB : bufA;
        if (EPS && a.eps_acc) b.eps_acc = a.eps_acc + (size_t)it * eps_stride;
        for (;;) {
            if (threadIdx.x == 0) s_item = atomicAdd(counters + it, 1);
            __syncthreads();
            const int32_t item = s_item;
            if (item >= total) break;
            const int32_t slab = item / nrg, rg = item % nrg;
            const int32_t row0 = rg * rows_per_item;
            gather_block_x<LOB, CPL, WARPS, U, false, false, EPS>(b, row0, min(row0 + rows_per_item, n_rows), slab,
                                                                   rg % kXEpsCopies, s_acc, &s_done);
            __syncthreads();
        }
        __threadfence();
        grid.sync();  // every row of this update is written (and visible) before anyone gathers from it
    }
}

// ---- small kernels ----------------------------------------------------------------------------------------------
// lo != null: the plan holds a float64 matrix as hi + lo (common.cuh); the sum is exact in float64.
__global__ void widen_pairs_kernel(int64_t nnz, const Pair *__restrict__ in, const float *__restrict__ lo,
                                   PairD *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) {
        const Pair e = in[i];
        out[i] = PairD{e.src, 0, lo ? (double)e.val + (double)lo[i] : (double)e.val};
    }
}

__global__ void x_pi_rows_kernel(int32_t n, const double *__restrict__ stationary, const int32_t *__restrict__ dest_ids,
                                 double *__restrict__ out) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r] = stationary[dest_ids ? dest_ids[r] : r];
}

// |x0| column sums in a fixed order: part[chunk][c] = sum over the chunk's rows (ascending), one thread per column
constexpr int kScaleChunks = 64;
__global__ void x_colsum_kernel(int32_t n, int32_t B, const float *__restrict__ X0, double *__restrict__ part) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    const int32_t per = (n + kScaleChunks - 1) / kScaleChunks;
    const int32_t r0 = min(n, (int32_t)blockIdx.y * per), r1 = min(n, r0 + per);
    double s = 0.0;
    for (int32_t r = r0; r < r1; ++r) s += fabs((double)X0[(size_t)r * B + c]);
    part[(size_t)blockIdx.y * B + c] = s;
}

// One CTA: m = max_c sum_i |x0[i,c]|, pmax = max_i |pi_i|, vv = sum pi^2 (fixed order), and the fixed-point scale
//   scale = 2^(56 - ceil(log2(max(m*pmax, m*m)))):  the sums of x.pi and x.x stay below 2^56, with 2^7 of headroom in
// the signed 64-bit accumulators for iterates whose mass grows over the run (T row sums within rounding of 1 give
// (1 + 1e-7)^iters), at a resolution of 2^-56 of the bound per partial sum.
__global__ void __launch_bounds__(1024) eps_scale_kernel(int32_t n, int32_t B, const double *__restrict__ part,
                                                          const double *__restrict__ stationary, double *out_scale,
                                                          double *out_vv) {
    __shared__ double s_m[32], s_p[32], s_v[1024];
    double m = 0.0, pm = 0.0, v = 0.0;
    for (int32_t c = threadIdx.x; c < B; c += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < kScaleChunks; ++k) s += part[(size_t)k * B + c];
        m = fmax(m, s);
    }
    // vv in a fixed order: thread t sums elements t, t+1024, ...; then a fixed tree
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double p = stationary[i];
        pm = fmax(pm, fabs(p));
        v = ::fma(p, p, v);
    }
    s_v[threadIdx.x] = v;
    for (int o = 16; o; o >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        pm = fmax(pm, __shfl_xor_sync(0xffffffffu, pm, o));
    }
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m, s_p[threadIdx.x >> 5] = pm;
    __syncthreads();
    for (int w = 512; w; w >>= 1) {
        if ((int)threadIdx.x < w) s_v[threadIdx.x] += s_v[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; ++k) m = fmax(m, s_m[k]), pm = fmax(pm, s_p[k]);
        double bound = fmax(m * pm, m * m);
        if (!(bound > 0.0) || !isfinite(bound)) bound = 1.0;
        int e;
        frexp(bound, &e);  // bound = f * 2^e, f in [0.5, 1)  ->  bound <= 2^e
        *out_scale = ldexp(1.0, 56 - e);
        *out_vv = s_v[0];
    }
}

// eps[i] = clip(1 - uv/sqrt(uu*vv), 0, 2) from the integer sums (copies added as integers: order-free)
__global__ void x_eps_finalize(int32_t iters, int32_t tile_cols, int32_t col0, int32_t B,
                               const unsigned long long *__restrict__ acc, const double *__restrict__ scale,
                               const double *__restrict__ vv, double *__restrict__ eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)iters * tile_cols) return;
    const int32_t it = (int32_t)(i / tile_cols), c = (int32_t)(i % tile_cols);
    const unsigned long long *a = acc + ((int64_t)it * kXEpsCopies * tile_cols + c) * 2;
    unsigned long long uv = 0, uu = 0;
    for (int32_t k = 0; k < kXEpsCopies; ++k) {
        uv += a[(int64_t)k * tile_cols * 2];
        uu += a[(int64_t)k * tile_cols * 2 + 1];
    }
    const double inv = 1.0 / scale[0];
    const double duv = (double)(long long)uv * inv, duu = (double)(long long)uu * inv;
    const double d = 1.0 - duv / sqrt(duu * vv[0]);
    eps[(int64_t)it * B + col0 + c] = fmin(fmax(d, 0.0), 2.0);
}

// ---- host side --------------------------------------------------------------------------------------------------
static inline size_t x_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int x_env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

struct XConfig {
    int lob, cpl, warps, rows_per_cta, minb, u, persist;
};

static XConfig x_config(int lo_bits) {
    XConfig c;
    c.lob = lo_bits == 8 ? 8 : 16;
    c.cpl = x_env_int("CY2PX_CPL", 8) == 4 ? 4 : 8;
    c.minb = x_env_int("CY2PX_MINB", 2);
    c.persist = x_env_int("CY2PX_PERSIST", 0);
    const int u = x_env_int("CY2PX_U", 4);
    c.u = u == 2 ? 2 : 4;
    const int w = x_env_int("CY2PX_WARPS", 8);
    c.warps = w >= 16 ? 16 : 8;
    const int r = x_env_int("CY2PX_ROWS", 0);
    c.rows_per_cta = r > 0 ? std::max(r, c.warps) : c.warps * 8;
    return c;
}

constexpr int kXCounterSlots = 4096;  // per-launch work counters of the persistent variant, zeroed per block of launches
struct XWs {
    size_t eps_acc, scal, part, bufA, bufB, pi_rows, counters, eps_part, total;
    int32_t tile;      // columns per tile (multiple of the slab width)
    uint32_t pitch;    // bytes per row of a scratch tile
};

static XWs x_ws_layout(const cy2p_plan *plan, int32_t B, int32_t iters, int want_eps, const XConfig &c) {
    XWs w;
    const int slab_cols = 32 * c.cpl;
    const size_t rec = (size_t)slab_cols * 4 + (size_t)slab_cols * c.lob / 8;
    int64_t bt = x_env_int("CY2P_TILE", 0);
    if (bt <= 0) {
        const size_t budget = (size_t)3 << 29;  // 1.5 GiB for the two ping-pong tiles
        bt = (int64_t)(budget / ((size_t)plan->n * 2 * rec)) * slab_cols;
        if (bt > 2048) bt = 2048;
    }
    bt = bt / slab_cols * slab_cols;
    if (bt < slab_cols) bt = slab_cols;
    const int64_t Bpad = ((int64_t)B + slab_cols - 1) / slab_cols * slab_cols;
    if (bt > Bpad) bt = Bpad;
    w.tile = (int32_t)bt;
    w.pitch = (uint32_t)((bt / slab_cols) * rec);
    size_t off = 0;
    w.eps_acc = off;
    if (want_eps) off = x_align_up(off + sizeof(unsigned long long) * 2 * (size_t)iters * (size_t)w.tile * kXEpsCopies, 256);
    w.scal = off;
    off = x_align_up(off + 4 * sizeof(double), 256);
    w.part = off;
    if (want_eps) off = x_align_up(off + sizeof(double) * (size_t)kScaleChunks * (size_t)B, 256);
    w.bufA = off;
    off = x_align_up(off + (size_t)plan->n * w.pitch, 256);
    w.bufB = off;
    off = x_align_up(off + (size_t)plan->n * w.pitch, 256);
    w.pi_rows = off;
    if (want_eps) off = x_align_up(off + sizeof(double) * (size_t)plan->n, 256);
    w.counters = off;
    off = x_align_up(off + sizeof(int32_t) * kXCounterSlots, 256);
    w.eps_part = off;
    if (want_eps) {  // one parked block of sums per CTA of the densest plain launch (row blocks of >= 32 rows)
        const int rpc = c.rows_per_cta < 64 ? (c.rows_per_cta < 32 ? 32 : c.rows_per_cta) : 64;
        const size_t nrb = ((size_t)plan->n + rpc - 1) / rpc;
        off = x_align_up(off + sizeof(unsigned long long) * nrb * (size_t)(w.tile / slab_cols) * 2 * slab_cols, 256);
    }
    w.total = off;
    return w;
}

static int32_t ensure_pairsD(cy2p_plan *plan, const Pair *src, const float *lo, PairD **dst, cudaStream_t st) {
    if (*dst || !src) return CY2P_OK;
    const size_t bytes = sizeof(PairD) * (size_t)(plan->nnz ? plan->nnz : 1);
    CY2P_CUDA(cudaMalloc((void **)dst, bytes));
    plan->bytes_owned += bytes;
    if (plan->nnz) {
        count_launch();
        widen_pairs_kernel<<<(unsigned)((plan->nnz + 255) / 256), 256, 0, st>>>(plan->nnz, src, lo, *dst);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

struct XPersist {
    int32_t *counter = nullptr;  // zeroed int for this launch, or null: plain grid
    int num_sms = 0;
};

template <int LOB, int CPL, int WARPS, int MINB, int U, bool IN_F32, bool OUT_F32>
static void x_launch_t(const XArgs &a, int32_t n, int rows_per_cta, cudaStream_t st, XPersist ps = XPersist()) {
    const int nslabs = (a.ncols + 32 * CPL - 1) / (32 * CPL);
    if constexpr (!IN_F32 && !OUT_F32) {
        if (ps.counter) {
            count_launch();
            const int ctas = ps.num_sms * MINB;
            spmm_gather_x_persist<LOB, CPL, WARPS, MINB, U, true><<<ctas, WARPS * 32, 0, st>>>(n, rows_per_cta, nslabs, ps.counter, a);
            return;
        }
    }
    const dim3 grid((unsigned)((n + rows_per_cta - 1) / rows_per_cta), (unsigned)nslabs);
    count_launch();
    spmm_gather_x<LOB, CPL, WARPS, MINB, U, IN_F32, OUT_F32, true><<<grid, WARPS * 32, 0, st>>>(0, n, rows_per_cta, a);
    if (a.eps_acc != nullptr && a.eps_part != nullptr) {
        const int outs = nslabs * 2 * CPL * 32;
        const int qs = grid.x >= 256 ? 16 : (grid.x >= 16 ? 4 : 1);
        count_launch();
        x_eps_reduce<CPL><<<dim3((unsigned)((outs + 255) / 256), (unsigned)qs), 256, 0, st>>>((int32_t)grid.x, nslabs, a.ncols,
                                                                                         a.eps_part, a.eps_acc);
    }
}

// tile -> tile updates (the bulk of a run): the tuning knobs select the instantiation. Measured on B200 at config 2
// (profiles/k1x_sweep2_r02.jsonl): the update runs against the L2 -> SM bandwidth cap, and what decides how close it
// gets is the number of gathered bytes a SM has in flight (24 KB: 4.5 ms per update, 49 KB: 3.1-3.8, 73 KB: 2.9-3.0,
// 98 KB: 2.7) — in-flight bytes are registers, so the best shape is few fat warps: 8 columns per lane, 4 row gathers
// in flight, 128 registers, 16 warps per SM. Shapes that spill (64-register budgets) lose 15-30%.
template <int LOB, int CPL, int U>
static void x_launch_hot(const XArgs &a, int32_t n, const XConfig &c, cudaStream_t st, XPersist ps) {
    if (c.warps == 16)
        x_launch_t<LOB, CPL, 16, 1, U, false, false>(a, n, c.rows_per_cta, st, ps);
    else if (c.minb == 3)
        x_launch_t<LOB, CPL, 8, 3, U, false, false>(a, n, c.rows_per_cta, st, ps);
    else
        x_launch_t<LOB, CPL, 8, 2, U, false, false>(a, n, c.rows_per_cta, st, ps);
}

template <int LOB, int CPL>
static void x_launch_fmt(const XArgs &a, int32_t n, const XConfig &c, bool in_f32, bool out_f32, cudaStream_t st,
                         XPersist ps) {
    if (!in_f32 && !out_f32) {
        if (c.u == 2)
            x_launch_hot<LOB, CPL, 2>(a, n, c, st, ps);
        else
            x_launch_hot<LOB, CPL, 4>(a, n, c, st, ps);
        return;
    }
    const int rpc = 64;  // first / last update of a run: float32 rows in or out
    if (in_f32 && out_f32)
        x_launch_t<LOB, CPL, 8, 2, 2, true, true>(a, n, rpc, st);
    else if (in_f32)
        x_launch_t<LOB, CPL, 8, 2, 2, true, false>(a, n, rpc, st);
    else
        x_launch_t<LOB, CPL, 8, 2, 2, false, true>(a, n, rpc, st);
}

static void x_launch(const XArgs &a, int32_t n, const XConfig &c, bool in_f32, bool out_f32, cudaStream_t st,
                     XPersist ps = XPersist()) {
    if (c.lob == 16) {
        if (c.cpl == 8)
            x_launch_fmt<16, 8>(a, n, c, in_f32, out_f32, st, ps);
        else
            x_launch_fmt<16, 4>(a, n, c, in_f32, out_f32, st, ps);
    } else {
        if (c.cpl == 8)
            x_launch_fmt<8, 8>(a, n, c, in_f32, out_f32, st, ps);
        else
            x_launch_fmt<8, 4>(a, n, c, in_f32, out_f32, st, ps);
    }
}

}  // namespace cy2p

using namespace cy2p;

extern "C" {

size_t cy2p_propagate_x_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters, int32_t want_eps, int32_t tail_bits) {
    if (!plan || B < 1 || iters < 0) return 0;
    return x_ws_layout(plan, B, iters, want_eps, x_config(tail_bits)).total;
}

int32_t cy2p_propagate_x_column_tile(const cy2p_plan *plan, int32_t B, int32_t tail_bits) {
    if (!plan || B < 1) return 0;
    return x_ws_layout(plan, B, 0, 0, x_config(tail_bits)).tile;
}

int32_t cy2p_propagate_x(cy2p_plan *plan, const float *d_X0, int32_t B, int32_t iters, const double *d_stationary,
                         float *d_Xfinal, double *d_Xfinal64, double *d_eps, int32_t tail_bits, void *d_ws,
                         size_t ws_bytes, void *stream) {
    CY2P_REQUIRE(plan && d_X0, "null plan / X0");
    CY2P_REQUIRE(B >= 1 && iters >= 0, "B >= 1 and iters >= 0 required");
    CY2P_REQUIRE(tail_bits == 16 || tail_bits == 8, "tail_bits must be 16 (f48) or 8 (f40)");
    CY2P_REQUIRE(!d_eps || d_stationary, "eps requested without a stationary vector");
    CY2P_REQUIRE(d_Xfinal || d_Xfinal64, "nothing to write: both final outputs are null");
    const XConfig c = x_config(tail_bits);
    const XWs w = x_ws_layout(plan, B, iters, d_eps != nullptr, c);
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    const int32_t n = plan->n;
    char *ws = (char *)d_ws;
    if (iters == 0) {
        if (d_Xfinal && d_Xfinal != d_X0)
            CY2P_CUDA(cudaMemcpyAsync(d_Xfinal, d_X0, sizeof(float) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));
        CY2P_REQUIRE(!d_Xfinal64, "iters == 0 with a float64 output is not supported");
        return CY2P_OK;
    }
    // locality ordering: worth building when the work amortises the one-off host pass (same rule as cy2p_propagate)
    if (plan->reorder_state == 0 && (int64_t)B * iters >= 4096 && B >= 32) {
        const int32_t rc = ensure_reordered(plan, st);
        if (rc != CY2P_OK) return rc;
    }
    const bool reord = plan->reorder_state == 1;
    int32_t rc;
    if (reord) {
        if ((rc = ensure_pairsD(plan, plan->d_r_pairs_pp, plan->d_r_lo_pp, &plan->d_r_pairsD_pp, st)) != CY2P_OK) return rc;
        if ((rc = ensure_pairsD(plan, plan->d_r_pairs_op, plan->d_r_lo_op, &plan->d_r_pairsD_op, st)) != CY2P_OK) return rc;
    } else {
        if ((rc = ensure_pairsD(plan, plan->d_t_pairs, plan->d_t_lo, &plan->d_t_pairsD, st)) != CY2P_OK) return rc;
    }
    unsigned long long *eps_acc = (unsigned long long *)(ws + w.eps_acc);
    double *scal = (double *)(ws + w.scal);  // [0] fixed-point scale, [1] |stationary|^2
    double *pi_rows = nullptr;
    if (d_eps) {
        pi_rows = (double *)(ws + w.pi_rows);
        double *part = (double *)(ws + w.part);
        count_launch(3);
        x_colsum_kernel<<<dim3((unsigned)((B + 255) / 256), kScaleChunks), 256, 0, st>>>(n, B, d_X0, part);
        eps_scale_kernel<<<1, 1024, 0, st>>>(n, B, part, d_stationary, scal, scal + 1);
        x_pi_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_stationary, reord ? plan->d_order : nullptr, pi_rows);
    }
    const bool wide_all = (B % 4 == 0) && (((uintptr_t)d_X0 & 15) == 0) && (!d_Xfinal || ((uintptr_t)d_Xfinal & 15) == 0);
    unsigned char *bufA = (unsigned char *)(ws + w.bufA), *bufB = (unsigned char *)(ws + w.bufB);
    int32_t *counters = (int32_t *)(ws + w.counters);
    int64_t launch_no = 0;
    for (int32_t c0 = 0; c0 < B; c0 += w.tile) {
        const int32_t nc = (B - c0 < w.tile) ? B - c0 : w.tile;
        if (d_eps)
            CY2P_CUDA(cudaMemsetAsync(eps_acc, 0, sizeof(unsigned long long) * 2 * (size_t)iters * (size_t)nc * kXEpsCopies, st));
        unsigned char *src = nullptr;
        const int multi_env = x_env_int("CY2PX_MULTI", 0);
        const bool multi = multi_env && iters >= 4 && c.cpl == 4 && c.warps == 8 && iters - 2 <= kXCounterSlots;
        for (int32_t it = 0; it < iters; ++it) {
            const bool first = it == 0, last = it + 1 == iters;
            XArgs a = {};
            a.indptr = reord ? plan->d_r_indptr : plan->d_t_indptr;
            a.pairs = reord ? (first ? plan->d_r_pairsD_op : plan->d_r_pairsD_pp) : plan->d_t_pairsD;
            a.dest_ids = reord ? plan->d_order : nullptr;
            a.out_orig = reord ? 1 : 0;
            a.xin = src;
            a.pitch_in = w.pitch;
            a.fin = d_X0 + c0;
            a.ld_in = (uint32_t)B;
            unsigned char *dst = (src == bufA) ? bufB : bufA;
            a.xout = dst;
            a.pitch_out = w.pitch;
            a.fout = d_Xfinal ? d_Xfinal + c0 : nullptr;
            a.dout = d_Xfinal64 ? d_Xfinal64 + c0 : nullptr;
            a.ld_out = (uint32_t)B;
            a.ncols = nc;
            a.wide_io = (wide_all && nc % c.cpl == 0 && c0 % 4 == 0) ? 1 : 0;
            a.pi_rows = pi_rows;
            a.eps_acc = d_eps ? eps_acc + 2 * (size_t)it * (size_t)nc * kXEpsCopies : nullptr;
            a.eps_scale = scal;
            a.dbg = x_env_int("CY2PX_EPS_DBG", 0);
            a.eps_part = (d_eps && c.rows_per_cta >= 32 && x_env_int("CY2PX_EPS_PARK", 1))
                             ? (unsigned long long *)(ws + w.eps_part) : nullptr;
            if (multi && it == 1) {  // updates 1 .. iters-2 in ONE cooperative launch
                const int32_t n_upd = iters - 2;
                CY2P_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * kXCounterSlots, st));
                int32_t n_ = n, rpi = c.rows_per_cta, nslabs = (nc + 32 * 4 - 1) / (32 * 4), srcA = (src == bufA) ? 1 : 0;
                size_t eps_stride = 2 * (size_t)nc * kXEpsCopies;
                void *args[] = {&n_, &rpi, &nslabs, (void *)&n_upd, &counters, &bufA, &bufB, &srcA, &eps_stride, &a};
                const dim3 grid((unsigned)(plan->num_sms * 2));
                count_launch();
                cudaError_t le;
                if (c.lob == 16)
                    le = cudaLaunchCooperativeKernel((void *)spmm_gather_x_multi<16, 4, 8, 2, 4, true>, grid, dim3(256), args, 0, st);
                else
                    le = cudaLaunchCooperativeKernel((void *)spmm_gather_x_multi<8, 4, 8, 2, 4, true>, grid, dim3(256), args, 0, st);
                CY2P_CUDA(le);
                if (n_upd & 1) src = dst;  // an odd number of updates ends in the other buffer
                it += n_upd - 1;
                continue;
            }
            XPersist ps;
            if (c.persist && !first && !last) {
                const int64_t slot = launch_no++ % kXCounterSlots;
                if (slot == 0) CY2P_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * kXCounterSlots, st));
                ps.counter = counters + slot;
                ps.num_sms = plan->num_sms;
            }
            x_launch(a, n, c, first, last, st, ps);
            src = dst;
        }
        if (d_eps) {
            const int64_t cnt = (int64_t)iters * nc;
            count_launch();
            x_eps_finalize<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(iters, nc, c0, B, eps_acc, scal, scal + 1, d_eps);
        }
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"

// ---- cross-GPU barrier between the updates of the row-partitioned run (cy2path_b200/rowpart.py) -------------------
// One tiny kernel, stream-ordered after the rank's update kernel (whose peer stores are complete when it retires):
// thread g of the single CTA writes this rank's epoch into slot `rank` of rank g's flag array (a peer store over
// NVLink; its own slot for g == rank) and then spins until slot g of its OWN array shows that epoch: when every thread
// is through, every rank's update of this round has retired. The epoch comes from a device counter the kernel itself
// advances, so the launch has no per-call argument and can sit in a CUDA graph that is replayed for every pair of
// updates (at B = 1 an update on 8 GPUs is ~10 us of work: the host-side launch path would otherwise bound it).
// A rank that never arrives must not hang the others for good: the spin gives up after ~2 s and raises *d_error.
namespace cy2p {
__global__ void xbarrier_kernel(int32_t *const *peer_flags, volatile int32_t *my_flags, int32_t *epoch_counter,
                                int32_t rank, int32_t world, int32_t *d_error) {
    __shared__ int32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = ++(*epoch_counter);
    __syncthreads();
    const int32_t epoch = s_epoch;
    const int g = threadIdx.x;
    if (g < world) {
        __threadfence_system();
        *(volatile int32_t *)(peer_flags[g] + rank) = epoch;
        const long long t0 = clock64();
        while (my_flags[g] - epoch < 0) {
            if (clock64() - t0 > 4000000000LL) {
                atomicExch(d_error, 1);
                break;
            }
        }
        __threadfence_system();
    }
}
}  // namespace cy2p

extern "C" int32_t cy2p_xbarrier(void *const *d_peer_flags, int32_t *d_my_flags, int32_t *d_epoch_counter, int32_t rank,
                                 int32_t world, int32_t *d_error, void *stream) {
    CY2P_REQUIRE(d_peer_flags && d_my_flags && d_epoch_counter && d_error, "null pointer");
    CY2P_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "1 <= world <= 64 and 0 <= rank < world required");
    count_launch();
    xbarrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((int32_t *const *)d_peer_flags, d_my_flags, d_epoch_counter, rank,
                                                        world, d_error);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}
