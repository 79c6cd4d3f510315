// cy2path_b200 — shared definitions of the sm_100a hot-path library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/cy2path_b200.h"

namespace cy2p {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);  // bookkeeping for cy2p_launch_count()

#define CY2P_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cy2p::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return CY2P_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define CY2P_REQUIRE(cond, msg)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            cy2p::set_error("invalid argument: %s", msg);        \
            return CY2P_ERR_INVALID_ARG;                         \
        }                                                        \
    } while (0)

// (source cell, transition probability) of one stored entry of T^T; 8 bytes so two entries
// travel in one 128-bit load.
struct __align__(8) Pair {
    int32_t src;
    float val;
};

// The same entry with the probability widened to float64 (exact): operand of the DFMA in the extended-precision
// batched update (propagate_x.cu) without a per-product conversion; one 128-bit load.
struct __align__(16) PairD {
    int32_t src;
    int32_t pad;
    double val;
};

}  // namespace cy2p

struct cy2p_plan {
    int32_t n = 0;
    int64_t nnz = 0;
    int32_t device = 0;
    int32_t num_sms = 0;
    int32_t max_in_deg = 0, max_out_deg = 0, W = 0;
    std::vector<int32_t> h_t_indptr;  // host copy of d_t_indptr (filled on demand for small graphs)
    int32_t cluster_ctas = 0;     // cluster size of the single-start kernel: 0 = not probed yet, -1 = unavailable
    bool canonical_rows = false;  // every row of T has strictly increasing column indices (no duplicates)
    size_t bytes_owned = 0;
    size_t l2_bytes = 0;
    // T as given (CSR, device copies owned by the plan)
    int32_t *d_indptr = nullptr;
    int32_t *d_indices = nullptr;
    float *d_data = nullptr;
    // T^T: destination-major CSR, sources ascending inside a row (scipy csc_matvec's add order)
    int32_t *d_t_indptr = nullptr;
    cy2p::Pair *d_t_pairs = nullptr;
    // Locality ordering (built lazily by reorder.cu for batched propagation): destinations in Cuthill-McKee
    // order so that the rows a CTA gathers lie in a narrow window of the permuted iterate (L2 hits instead of HBM)
    int32_t reorder_state = 0;             // 0 = not tried, 1 = active, -1 = rejected (no gain) / disabled
    int32_t *d_order = nullptr;            // [n] original cell id of permuted position q
    int32_t *d_r_indptr = nullptr;         // [n+1] T^T row pointers in permuted destination order
    cy2p::Pair *d_r_pairs_pp = nullptr;    // sources as permuted positions  (permuted iterate -> permuted iterate)
    cy2p::Pair *d_r_pairs_op = nullptr;    // sources as original ids        (original iterate -> permuted iterate)
    double span_identity = 0.0, span_reordered = 0.0;  // mean |pos(i) - pos(j)| over the stored entries
    // float64-valued copies of the three pair arrays (built on first use by propagate_x.cu)
    cy2p::PairD *d_t_pairsD = nullptr, *d_r_pairsD_pp = nullptr, *d_r_pairsD_op = nullptr;
    // A float64 transition matrix (cy2p_plan_create_host_f64) is held as value = hi + lo with hi = float32(value) in
    // the arrays above and lo = float32(value - hi) here, entry for entry: (double)hi + (double)lo restores 48 mantissa
    // bits, which is what the float64-arithmetic paths (single-start float64, batched f48) multiply by. Null for a
    // float32 matrix.
    float *d_data_lo = nullptr;  // CSR order (with d_data)
    float *d_t_lo = nullptr;     // T^T order (with d_t_pairs)
    float *d_r_lo_pp = nullptr, *d_r_lo_op = nullptr;  // with d_r_pairs_pp / d_r_pairs_op
    // Staged gather (propagate.cu spmm_staged): the permuted rows cut into blocks whose distinct sources fit a CTA's
    // shared memory; sources are named by their index in the block's list
    int32_t nblk = 0;                      // 0 = not built
    int32_t blk_ucap = 0;                  // max distinct sources of a block
    int32_t *d_blk_row = nullptr;          // [nblk+1] first permuted row of each block
    int32_t *d_blk_uptr = nullptr;         // [nblk+1] offsets into d_blk_usrc
    int32_t *d_blk_usrc = nullptr;         // distinct sources (permuted positions, ascending) of each block
    cy2p::Pair *d_r_pairs_lp = nullptr;    // d_r_pairs_pp with src replaced by the index in the block's list
    double blk_reuse = 0.0;                // stored entries / distinct sources, averaged over the blocks
    // side stream + events of cy2p_fhmm_loss_grad (fhmm.cu): the node-level passes run beside the pass over D
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_events[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace cy2p {
int32_t ensure_reordered(cy2p_plan *plan, cudaStream_t st);  // reorder.cu
}
