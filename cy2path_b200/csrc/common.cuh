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

// Row blocks of the TMA-staged propagation kernel: a block is a run of consecutive (permuted)
// destination rows whose union of source cells fits the shared-memory stage.
struct RowBlock {
    int32_t row_begin, row_end;  // destination rows [begin, end) in plan order
    int32_t src_begin, src_end;  // range into plan->d_block_src (unique sources of the block)
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
    // staged kernel structures (built lazily by propagate_staged.cu)
    int32_t num_blocks = 0;
    cy2p::RowBlock *d_blocks = nullptr;
    int32_t *d_block_src = nullptr;   // concatenated unique-source lists
    cy2p::Pair *d_local_pairs = nullptr;  // d_t_pairs with src replaced by the slot in the block's stage
    int32_t *d_perm = nullptr;        // plan order -> original cell id (nullptr = identity)
    int32_t *d_iperm = nullptr;
};
