// K0 — plan creation: validate T (CSR), build T^T (destination-major CSR with sources ascending),
// and the CDF tables of the chain sampler.
//
// Replaces: the T.T CSC view + csc_matvec operand handling scipy performs on every call of
// reference sampling.py:40-42, and extract_nonzero_entries (reference simulation.py:15-53).
#include <cub/cub.cuh>
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cy2p {

static thread_local char g_err[512] = "";

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------- validation + degree statistics
// flags[0] = first error code (0 ok), flags[1] = offending row/entry, flags[2] = max out-degree,
// flags[3] = max in-degree, flags[4] = CDF width W, flags[5] = some row is not strictly increasing in its column indices
__global__ void validate_rows_kernel(int32_t n, int64_t nnz, const int32_t *__restrict__ indptr, int32_t *flags) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && (indptr[0] != 0 || (int64_t)indptr[n] != nnz)) atomicCAS(&flags[0], 0, 1);
    if (i >= n) return;
    int32_t len = indptr[i + 1] - indptr[i];
    if (len < 0) {
        if (atomicCAS(&flags[0], 0, 2) == 0) flags[1] = i;
        return;
    }
    atomicMax(&flags[2], len);
}

__global__ void count_columns_kernel(int64_t nnz, int32_t n, const int32_t *__restrict__ indices,
                                     int32_t *__restrict__ indeg, int32_t *flags) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int32_t c = indices[p];
    if (c < 0 || c >= n) {
        if (atomicCAS(&flags[0], 0, 3) == 0) flags[1] = (int32_t)min(p, (int64_t)INT32_MAX);
        return;
    }
    atomicAdd(&indeg[c], 1);
}

__global__ void max_kernel(int32_t n, const int32_t *__restrict__ v, int32_t *out) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    int32_t m = i < n ? v[i] : 0;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// one thread per source row: row id of every stored entry, entry ordinal as the sort payload
__global__ void expand_rows_kernel(int32_t n, const int32_t *__restrict__ indptr, int32_t *__restrict__ rowid,
                                   int32_t *__restrict__ ordinal) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p) {
        rowid[p] = i;
        ordinal[p] = p;
    }
}

__global__ void fill_pairs_kernel(int64_t nnz, const int32_t *__restrict__ sorted_ordinal,
                                  const int32_t *__restrict__ rowid, const float *__restrict__ data,
                                  Pair *__restrict__ pairs) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    int32_t p = sorted_ordinal[q];
    Pair e;
    e.src = rowid[p];
    e.val = data[p];
    pairs[q] = e;
}

__global__ void fill_lo_kernel(int64_t nnz, const int32_t *__restrict__ sorted_ordinal, const float *__restrict__ lo_csr,
                               float *__restrict__ t_lo) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nnz) t_lo[q] = lo_csr[sorted_ordinal[q]];
}

// CDF width W = max over rows of the number of non-zeros of the DENSE row (reference
// simulation.py:28-41 uses count_nonzero(T.toarray())): distinct columns whose summed value != 0.
__global__ void cdf_width_kernel(int32_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                 const float *__restrict__ data, int32_t *flags) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    int32_t cnt = 0;
    if (i < n) {
        int32_t s = indptr[i], e = indptr[i + 1];
        bool strictly_increasing = true;
        for (int32_t p = s + 1; p < e; ++p) strictly_increasing &= indices[p] > indices[p - 1];
        if (strictly_increasing) {
            for (int32_t p = s; p < e; ++p) cnt += data[p] != 0.f;
        } else {  // duplicates / unsorted: O(len^2), rows are short
            atomicExch(&flags[5], 1);
            for (int32_t p = s; p < e; ++p) {
                bool first = true;
                for (int32_t q = s; q < p; ++q) first &= indices[q] != indices[p];
                if (!first) continue;
                float sum = 0.f;
                for (int32_t q = p; q < e; ++q)
                    if (indices[q] == indices[p]) sum += data[q];
                cnt += sum != 0.f;
            }
        }
    }
    int32_t m = cnt;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&flags[4], m);
}

// extract_nonzero_entries, one thread per row. fp32 sequential adds in storage order; the rounding is
// numpy's round(x, 3) on float32: rint(x * 1000f) / 1000f, each op rounded to fp32 (no FMA contraction).
// lo != null (float64 matrix): the reference's cumsum runs in the matrix's dtype (simulation.py:24-25), is cast to
// float32 on assignment (:44-46) and rounded there (:49-51) — sequential float64 adds of hi + lo, one cast, same rounding.
__global__ void cdf_build_kernel(int32_t n, int32_t W, const int32_t *__restrict__ indptr,
                                 const int32_t *__restrict__ indices, const float *__restrict__ data,
                                 const float *__restrict__ lo, int32_t *__restrict__ idx, float *__restrict__ cum,
                                 int32_t *flag_too_long) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t s = indptr[i], e = indptr[i + 1];
    if (e - s > W) {
        atomicExch(flag_too_long, 1);
        e = s + W;
    }
    float run = 0.f;
    double run64 = 0.0;
    int32_t j = 0;
    for (int32_t p = s; p < e; ++p, ++j) {
        if (lo != nullptr) {
            run64 = __dadd_rn(run64, __dadd_rn((double)data[p], (double)lo[p]));
            run = __double2float_rn(run64);
        } else {
            run = __fadd_rn(run, data[p]);
        }
        idx[(size_t)i * W + j] = indices[p];
        cum[(size_t)i * W + j] = __fdiv_rn(rintf(__fmul_rn(run, 1000.f)), 1000.f);
    }
    for (; j < W; ++j) {
        idx[(size_t)i * W + j] = 0;
        cum[(size_t)i * W + j] = 0.f;
    }
}

// Temporaries come from the device's stream-ordered pool (cudaMallocAsync): plan creation sits inside the timed
// end-to-end path of the small single-start case, and cudaMalloc/cudaFree cost ~0.1-0.5 ms each.
static void keep_pool_cached(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}

template <typename T>
static int32_t dmalloc(cy2p_plan *plan, T **ptr, size_t count, cudaStream_t st) {
    size_t bytes = sizeof(T) * (count ? count : 1);
    CY2P_CUDA(cudaMallocAsync((void **)ptr, bytes, st));
    plan->bytes_owned += bytes;
    return CY2P_OK;
}

static void plan_free(cy2p_plan *p) {
    if (!p) return;
    cudaFree(p->d_indptr);
    cudaFree(p->d_indices);
    cudaFree(p->d_data);
    cudaFree(p->d_t_indptr);
    cudaFree(p->d_t_pairs);
    cudaFree(p->d_order);
    cudaFree(p->d_r_indptr);
    cudaFree(p->d_r_pairs_pp);
    cudaFree(p->d_r_pairs_op);
    cudaFree(p->d_blk_row);
    cudaFree(p->d_blk_uptr);
    cudaFree(p->d_blk_usrc);
    cudaFree(p->d_r_pairs_lp);
    if (p->side_stream) cudaStreamDestroy(p->side_stream);
    for (cudaEvent_t e : p->side_events)
        if (e) cudaEventDestroy(e);
    cudaFree(p->d_data_lo);
    cudaFree(p->d_t_lo);
    cudaFree(p->d_r_lo_pp);
    cudaFree(p->d_r_lo_op);
    cudaFree(p->d_t_pairsD);
    cudaFree(p->d_r_pairsD_pp);
    cudaFree(p->d_r_pairsD_op);
    delete p;
}

static int32_t plan_build(int32_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, const float *data,
                          cudaMemcpyKind kind, int32_t device, cudaStream_t st, cy2p_plan **out,
                          const float *lo = nullptr /* residuals of a float64 matrix, same order and memory space as data */) {
    CY2P_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    CY2P_REQUIRE(n > 0 && nnz >= 0 && nnz < (int64_t)INT32_MAX, "n must be > 0 and 0 <= nnz < 2^31");
    CY2P_REQUIRE(indptr && (nnz == 0 || (indices && data)), "null CSR array");
    CY2P_CUDA(cudaSetDevice(device));
    cy2p_plan *plan = new cy2p_plan();
    plan->n = n;
    plan->nnz = nnz;
    plan->device = device;
    cudaDeviceProp prop;
    CY2P_CUDA(cudaGetDeviceProperties(&prop, device));
    plan->num_sms = prop.multiProcessorCount;
    plan->l2_bytes = (size_t)prop.l2CacheSize;

    int32_t rc;
#define TRY(x)              \
    do {                    \
        rc = (x);           \
        if (rc != CY2P_OK) { \
            plan_free(plan); \
            return rc;      \
        }                   \
    } while (0)
#define TRY_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));         \
            plan_free(plan);                                                                        \
            return CY2P_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

    keep_pool_cached(device);
    TRY(dmalloc(plan, &plan->d_indptr, (size_t)n + 1, st));
    TRY(dmalloc(plan, &plan->d_indices, (size_t)nnz, st));
    TRY(dmalloc(plan, &plan->d_data, (size_t)nnz, st));
    TRY(dmalloc(plan, &plan->d_t_indptr, (size_t)n + 1, st));
    TRY(dmalloc(plan, &plan->d_t_pairs, (size_t)nnz, st));
    TRY_CUDA(cudaMemcpyAsync(plan->d_indptr, indptr, sizeof(int32_t) * ((size_t)n + 1), kind, st));
    if (nnz) {
        TRY_CUDA(cudaMemcpyAsync(plan->d_indices, indices, sizeof(int32_t) * (size_t)nnz, kind, st));
        TRY_CUDA(cudaMemcpyAsync(plan->d_data, data, sizeof(float) * (size_t)nnz, kind, st));
        if (lo != nullptr) {
            TRY(dmalloc(plan, &plan->d_data_lo, (size_t)nnz, st));
            TRY(dmalloc(plan, &plan->d_t_lo, (size_t)nnz, st));
            TRY_CUDA(cudaMemcpyAsync(plan->d_data_lo, lo, sizeof(float) * (size_t)nnz, kind, st));
        }
    }

    // temporaries (freed before return)
    int32_t *flags = nullptr, *indeg = nullptr, *rowid = nullptr, *ord_in = nullptr, *ord_out = nullptr,
            *key_out = nullptr;
    void *cub_tmp = nullptr;
    auto free_tmp = [&]() {
        cudaFreeAsync(flags, st);
        cudaFreeAsync(indeg, st);
        cudaFreeAsync(rowid, st);
        cudaFreeAsync(ord_in, st);
        cudaFreeAsync(ord_out, st);
        cudaFreeAsync(key_out, st);
        cudaFreeAsync(cub_tmp, st);
    };
#define TRY_CUDA_T(call)                                                                            \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));         \
            free_tmp();                                                                             \
            plan_free(plan);                                                                        \
            return CY2P_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

    keep_pool_cached(device);
    const size_t nz1 = (size_t)(nnz ? nnz : 1);
    TRY_CUDA_T(cudaMallocAsync((void **)&flags, sizeof(int32_t) * 8, st));
    TRY_CUDA_T(cudaMallocAsync((void **)&indeg, sizeof(int32_t) * ((size_t)n + 1), st));
    TRY_CUDA_T(cudaMallocAsync((void **)&rowid, sizeof(int32_t) * nz1, st));
    TRY_CUDA_T(cudaMallocAsync((void **)&ord_in, sizeof(int32_t) * nz1, st));
    TRY_CUDA_T(cudaMallocAsync((void **)&ord_out, sizeof(int32_t) * nz1, st));
    TRY_CUDA_T(cudaMallocAsync((void **)&key_out, sizeof(int32_t) * nz1, st));
    TRY_CUDA_T(cudaMemsetAsync(flags, 0, sizeof(int32_t) * 8, st));
    TRY_CUDA_T(cudaMemsetAsync(indeg, 0, sizeof(int32_t) * ((size_t)n + 1), st));

    const int TB = 256;
    const unsigned gn = (unsigned)((n + TB - 1) / TB), gz = (unsigned)((nnz + TB - 1) / TB);
    count_launch(); validate_rows_kernel<<<gn, TB, 0, st>>>(n, nnz, plan->d_indptr, flags);
    int32_t h_flags[8];
    TRY_CUDA_T(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    TRY_CUDA_T(cudaStreamSynchronize(st));
    if (h_flags[0] != 0) {
        set_error(h_flags[0] == 1 ? "bad CSR: indptr[0] != 0 or indptr[n] != nnz"
                                  : "bad CSR: indptr decreases at row %d", h_flags[1]);
        free_tmp();
        plan_free(plan);
        return CY2P_ERR_BAD_CSR;
    }
    if (nnz) {
        count_launch();
        count_columns_kernel<<<gz, TB, 0, st>>>(nnz, n, plan->d_indices, indeg, flags);
        TRY_CUDA_T(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        TRY_CUDA_T(cudaStreamSynchronize(st));
        if (h_flags[0] != 0) {
            set_error("bad CSR: column index out of range at entry %d", h_flags[1]);
            free_tmp();
            plan_free(plan);
            return CY2P_ERR_BAD_CSR;
        }
    }
    count_launch(); max_kernel<<<gn, TB, 0, st>>>(n, indeg, &flags[3]);
    count_launch(); cdf_width_kernel<<<gn, TB, 0, st>>>(n, plan->d_indptr, plan->d_indices, plan->d_data, flags);

    // T^T row pointer = exclusive scan of the in-degrees (n+1 entries, last = nnz)
    size_t tmp_scan = 0, tmp_sort = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, indeg, plan->d_t_indptr, n + 1, st);
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < (int64_t)n) ++end_bit;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, plan->d_indices, key_out, ord_in, ord_out, (int)nnz, 0,
                                    end_bit, st);
    size_t tmp_bytes = tmp_scan > tmp_sort ? tmp_scan : tmp_sort;
    TRY_CUDA_T(cudaMallocAsync(&cub_tmp, tmp_bytes ? tmp_bytes : 1, st));
    TRY_CUDA_T(cub::DeviceScan::ExclusiveSum(cub_tmp, tmp_scan, indeg, plan->d_t_indptr, n + 1, st));
    if (nnz) {
        count_launch();
        expand_rows_kernel<<<gn, TB, 0, st>>>(n, plan->d_indptr, rowid, ord_in);
        // LSD radix sort is stable: entries of one destination keep ascending source order
        TRY_CUDA_T(cub::DeviceRadixSort::SortPairs(cub_tmp, tmp_sort, plan->d_indices, key_out, ord_in, ord_out,
                                                   (int)nnz, 0, end_bit, st));
        count_launch(); fill_pairs_kernel<<<gz, TB, 0, st>>>(nnz, ord_out, rowid, plan->d_data, plan->d_t_pairs);
        if (plan->d_data_lo) {
            count_launch();
            fill_lo_kernel<<<gz, TB, 0, st>>>(nnz, ord_out, plan->d_data_lo, plan->d_t_lo);
        }
    }
    TRY_CUDA_T(cudaGetLastError());
    TRY_CUDA_T(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    TRY_CUDA_T(cudaStreamSynchronize(st));
    plan->max_out_deg = h_flags[2];
    plan->max_in_deg = h_flags[3];
    plan->W = h_flags[4];
    plan->canonical_rows = h_flags[5] == 0;
    free_tmp();
    *out = plan;
    return CY2P_OK;
#undef TRY
#undef TRY_CUDA
#undef TRY_CUDA_T
}

}  // namespace cy2p

using namespace cy2p;

extern "C" {

int32_t cy2p_abi_version(void) { return CY2P_ABI_VERSION; }
const char *cy2p_last_error(void) { return g_err; }
int64_t cy2p_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int32_t cy2p_plan_create(int32_t n, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                         const float *d_data, int32_t device, void *stream, cy2p_plan **out) {
    return plan_build(n, nnz, d_indptr, d_indices, d_data, cudaMemcpyDeviceToDevice, device, (cudaStream_t)stream,
                      out);
}

int32_t cy2p_plan_create_host(int32_t n, int64_t nnz, const int32_t *h_indptr, const int32_t *h_indices,
                              const float *h_data, int32_t device, void *stream, cy2p_plan **out) {
    return plan_build(n, nnz, h_indptr, h_indices, h_data, cudaMemcpyHostToDevice, device, (cudaStream_t)stream,
                      out);
}

int32_t cy2p_plan_create_host_f64(int32_t n, int64_t nnz, const int32_t *h_indptr, const int32_t *h_indices,
                                  const double *h_data, int32_t device, void *stream, cy2p_plan **out) {
    CY2P_REQUIRE(nnz >= 0 && (nnz == 0 || h_data), "null data");
    std::vector<float> hi((size_t)nnz), lo((size_t)nnz);
    bool any = false;
    for (int64_t p = 0; p < nnz; ++p) {
        hi[p] = (float)h_data[p];
        lo[p] = (float)(h_data[p] - (double)hi[p]);
        any |= lo[p] != 0.f;
    }
    // (a float64 array whose values are all float32-representable is an ordinary float32 matrix)
    return plan_build(n, nnz, h_indptr, h_indices, hi.data(), cudaMemcpyHostToDevice, device, (cudaStream_t)stream, out,
                      any ? lo.data() : nullptr);
}

int32_t cy2p_plan_destroy(cy2p_plan *plan) {
    if (plan) {
        cudaSetDevice(plan->device);
        plan_free(plan);
    }
    return CY2P_OK;
}

int32_t cy2p_plan_info(const cy2p_plan *plan, int64_t info[8]) {
    CY2P_REQUIRE(plan && info, "null plan/info");
    info[0] = plan->n;
    info[1] = plan->nnz;
    info[2] = plan->max_in_deg;
    info[3] = plan->max_out_deg;
    info[4] = plan->W;
    info[5] = plan->device;
    info[6] = plan->reorder_state;
    info[7] = (int64_t)plan->bytes_owned;
    return CY2P_OK;
}

int32_t cy2p_cdf_build(const cy2p_plan *plan, int32_t *d_idx, float *d_cum, int32_t W, void *stream) {
    CY2P_REQUIRE(plan && d_idx && d_cum, "null pointer");
    CY2P_REQUIRE(W >= 1, "W must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    int32_t *flag = nullptr;
    CY2P_CUDA(cudaMalloc((void **)&flag, sizeof(int32_t)));
    CY2P_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    const int TB = 128;
    count_launch(); cdf_build_kernel<<<(plan->n + TB - 1) / TB, TB, 0, st>>>(plan->n, W, plan->d_indptr, plan->d_indices,
                                                             plan->d_data, plan->d_data_lo, d_idx, d_cum, flag);
    int32_t h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(flag);
    if (e != cudaSuccess) {
        set_error("cdf_build: %s", cudaGetErrorString(e));
        return CY2P_ERR_CUDA;
    }
    if (h) {
        set_error("index %d is out of bounds for axis 1 with size %d (a row stores more entries than W)", W, W);
        return CY2P_ERR_CDF_TABLE;
    }
    return CY2P_OK;
}

}  // extern "C"
