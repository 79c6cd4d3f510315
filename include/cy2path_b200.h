/* cy2path_b200 — C ABI of the B200-native hot path of cy2path.
 *
 * The reference (aron0093/cy2path) has no FFI/plugin interface: its boundary is a Python
 * function API over AnnData. This header is the C-ABI that sits directly
 * underneath that API; each entry point names the reference function it replaces. The
 * Python mirror of the reference API (cy2path_b200/{sampling,simulation,models}) binds these
 * symbols with ctypes (cy2path_b200/_lib.py); INTEGRATION.md shows the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross the boundary
 *  - pointers named d_* are DEVICE pointers on the plan's device, h_* are HOST pointers
 *  - every function returns an int32 status (CY2P_OK == 0); no exceptions, no aborts;
 *    cy2p_last_error() returns a thread-local message for the last failure
 *  - kernels never allocate: work space comes from the caller (cy2p_*_ws_bytes); only
 *    cy2p_plan_create allocates (plan-owned, freed by cy2p_plan_destroy)
 *  - stream is a cudaStream_t passed as void*; all work is enqueued on it (functions that
 *    must return a host value synchronise that stream and say so)
 *  - re-entrant; the only globals are the thread-local error string and the launch counter;
 *    one plan per (T, device)
 */
#ifndef CY2PATH_B200_H
#define CY2PATH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: round-2 additions, all additive (cy2p_propagate_x*, cy2p_plan_create_host_f64, cy2p_order_host, cy2p_xbarrier,
 * cy2p_propagate_rows_persist*); every version-1 entry point keeps its signature and meaning. */
#define CY2P_ABI_VERSION 2

enum cy2p_status {
    CY2P_OK = 0,
    CY2P_ERR_INVALID_ARG = 1,   /* null pointer, negative size, misaligned buffer           */
    CY2P_ERR_BAD_CSR = 2,       /* indptr not monotone / column index out of range          */
    CY2P_ERR_CDF_OVERSHOOT = 3, /* a uniform exceeded the last CDF entry of a row (the      */
                                /* reference raises IndexError at simulation.py:78-81)      */
    CY2P_ERR_CUDA = 4,          /* a CUDA call failed; message carries cudaGetErrorString   */
    CY2P_ERR_WORKSPACE = 5,     /* caller's work space too small                            */
    CY2P_ERR_CDF_TABLE = 6      /* a row stores more entries than the table width W (the    */
                                /* reference raises IndexError at simulation.py:43-46)      */
};

typedef struct cy2p_plan cy2p_plan;

int32_t cy2p_abi_version(void);
const char *cy2p_last_error(void);
/* Number of kernels of this library launched so far by the process (bench.py's gpu_launches). */
int64_t cy2p_launch_count(void);

/* ---- plan: device-resident T (CSR), T^T (CSR, stable: sources ascending), packed (col,val) ----
 * Replaces the implicit ``adata.obsp[matrix_key]`` operand of sampling.py:40-42 and the T.T CSC
 * view scipy builds per call. Validates the CSR. Synchronises `stream` (returns validation). */
int32_t cy2p_plan_create(int32_t n, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                         const float *d_data, int32_t device, void *stream, cy2p_plan **out);
/* Same, from host arrays (copies H2D inside). */
int32_t cy2p_plan_create_host(int32_t n, int64_t nnz, const int32_t *h_indptr, const int32_t *h_indices,
                              const float *h_data, int32_t device, void *stream, cy2p_plan **out);
/* Same, for a float64 matrix (a user-supplied ``adata.obsp[matrix_key]`` of dtype float64; sampling.py:40-42 then
 * multiplies in float64 and simulation.py:24-25 accumulates the CDF in float64). Values are held as float32 pairs
 * hi + lo (48 mantissa bits, 2^-48 relative): cy2p_propagate_f64 with B == 1, cy2p_propagate_x and cy2p_cdf_build use
 * hi + lo; every other entry point uses hi = float32(value). Values that are all float32-representable give an
 * ordinary plan. */
int32_t cy2p_plan_create_host_f64(int32_t n, int64_t nnz, const int32_t *h_indptr, const int32_t *h_indices,
                                  const double *h_data, int32_t device, void *stream, cy2p_plan **out);
int32_t cy2p_plan_destroy(cy2p_plan *plan);
/* info[0]=n info[1]=nnz info[2]=max in-degree info[3]=max out-degree (stored) info[4]=CDF table width W
 * info[5]=device info[6]=locality ordering (0 not built yet, 1 active, -1 rejected) info[7]=bytes of device memory owned */
int32_t cy2p_plan_info(const cy2p_plan *plan, int64_t info[8]);

/* ---- K1: MSM state-probability propagation  X <- X . T, `iters` times ----------------------
 * Replaces iterate_state_probability's loop (sampling.py:38-45): the product (:40-42) and the
 * cosine distance to the stationary vector (:44-45), for B start distributions at once.
 *   d_X0        [n][B] float32, cell-major (row = cell, B contiguous);  B == 1 is the reference's case
 *   d_stationary[n] float64 or NULL (then d_eps must be NULL)
 *   d_Xfinal    [n][B] float32: the iterate after `iters` updates (may be NULL if history covers it)
 *   d_history   NULL, or [ceil(iters/history_stride)][n][B] float32: row h holds the iterate BEFORE
 *               update h*history_stride (history_stride==1 reproduces state_history_max_iter)
 *   d_eps       NULL, or [iters][B] float64: eps[i][b] = 1 - cos(x_{i+1}[:,b], stationary), clipped [0,2]
 *   d_ws        work space of at least cy2p_propagate_ws_bytes(plan, B, iters) bytes, 256-byte aligned
 * Arithmetic: fp32 products and row sums in ascending source order (the order scipy's csc_matvec
 * adds them, there in float64); the eps dot products are summed per thread (4 rows) in the iterate's precision
 * and across threads / CTAs in float64 (agreement with a float64 recomputation: ~1e-7 for fp32 iterates). */
size_t cy2p_propagate_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters);
/* Width of the column tiles cy2p_propagate carries through all iterations (sized to stay L2-resident). */
int32_t cy2p_propagate_column_tile(const cy2p_plan *plan, int32_t B);
int32_t cy2p_propagate(cy2p_plan *plan, const float *d_X0, int32_t B, int32_t iters,
                       const double *d_stationary, float *d_Xfinal, float *d_history,
                       int32_t history_stride, double *d_eps, void *d_ws, size_t ws_bytes, void *stream);

/* float64 iterate: same contract with double X0 / Xfinal / history. Used for B == 1 — the reference's
 * own case, whose history is float64 (sampling.py:33) — and for long batched runs: a float32 iterate
 * cannot represent the ~1e-8 per-step growth that float32-rounded row sums give the reference's
 * float64 iterate, so its relative L1 distance to the reference grows linearly with the iteration
 * count (DESIGN.md "precision"). */
size_t cy2p_propagate_f64_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters);
int32_t cy2p_propagate_f64(cy2p_plan *plan, const double *d_X0, int32_t B, int32_t iters,
                           const double *d_stationary, double *d_Xfinal, double *d_history,
                           int32_t history_stride, double *d_eps, void *d_ws, size_t ws_bytes, void *stream);

/* Extended-precision iterate for the BATCHED case (the default of cy2path_b200.sampling.propagate_batch): same
 * update, but the products and row sums are float64 (like scipy's upcast product at sampling.py:40-42) and the
 * iterate between updates is a truncated float64 — sign, 11-bit exponent and 20 + tail_bits mantissa bits, i.e.
 * 6 bytes per element for tail_bits = 16 ("f48"), 5 for tail_bits = 8 ("f40") — in a tiled work-space layout
 * (csrc/propagate_x.cu). A float32 iterate leaves the reference's float64 iterate by a relative L1 distance of
 * ~2e-8 per update (4e-5 after 2,000); f48 stays below 1e-9, f40 below 2e-7 (tools/precision_lab.py).
 *   d_X0        [n][B] float32 (exact in the wider format)
 *   d_Xfinal    [n][B] float32 or NULL;  d_Xfinal64 [n][B] float64 or NULL (at least one)
 *   d_eps       NULL or [iters][B] float64, BIT-REPRODUCIBLE: float64 partial sums per thread in a fixed order,
 *               64-bit fixed-point integer atomics across threads (integer addition is associative)
 *   want_eps    for the work-space size: 1 if d_eps will be non-NULL */
size_t cy2p_propagate_x_ws_bytes(const cy2p_plan *plan, int32_t B, int32_t iters, int32_t want_eps, int32_t tail_bits);
int32_t cy2p_propagate_x_column_tile(const cy2p_plan *plan, int32_t B, int32_t tail_bits);
int32_t cy2p_propagate_x(cy2p_plan *plan, const float *d_X0, int32_t B, int32_t iters,
                         const double *d_stationary, float *d_Xfinal, double *d_Xfinal64, double *d_eps,
                         int32_t tail_bits, void *d_ws, size_t ws_bytes, void *stream);

/* Row-partitioned variant (BASELINE config 5): computes only destination rows [row_begin,row_end)
 * of one update, Y[row_begin:row_end, :] = (X . T)[row_begin:row_end, :], reading the whole X.
 * The caller (cy2path_b200/dist.py) all-gathers the slices between iterations. */
int32_t cy2p_propagate_rows(cy2p_plan *plan, const float *d_X, float *d_Y, int32_t B,
                            int32_t row_begin, int32_t row_end, void *stream);
int32_t cy2p_propagate_rows_f64(cy2p_plan *plan, const double *d_X, double *d_Y, int32_t B,
                                int32_t row_begin, int32_t row_end, void *stream);

/* Locality order of the plan (reorder.cu): d_order[q] = original cell id at position q of the plan's Cuthill-McKee
 * order (identity if the ordering was rejected because the input is already well ordered). Builds the ordering
 * on first use (host pass, synchronises `stream`). The *_ordered row update takes X and Y already in that order,
 * so a row-partitioned run permutes the start distributions once and un-permutes the result once. */
int32_t cy2p_plan_order(cy2p_plan *plan, int32_t *d_order, void *stream);
int32_t cy2p_propagate_rows_ordered(cy2p_plan *plan, const float *d_Xp, float *d_Yp, int32_t B,
                                    int32_t q_begin, int32_t q_end, void *stream);
int32_t cy2p_propagate_rows_ordered_f64(cy2p_plan *plan, const double *d_Xp, double *d_Yp, int32_t B,
                                        int32_t q_begin, int32_t q_end, void *stream);

/* Host-only (no device): the locality order a plan would build for this CSR matrix (Cuthill-McKee + groups of `group`
 * cells that share sources; group <= 1: plain Cuthill-McKee). h_order_out[q] = original cell id at position q.
 * Used by the row-partitioned driver to order the whole graph once on the host and give each rank a plan of its own
 * row slice only (cy2path_b200/rowpart.py). */
int32_t cy2p_order_host(int32_t n, int64_t nnz, const int32_t *h_indptr, const int32_t *h_indices, int32_t group,
                        int32_t *h_order_out);

/* Fused update + all-gather for the row partition: the kernel stores every finished row of [q_begin, q_end) into
 * ALL `npeers` next-iterate buffers (d_peer_out: device array of npeers base pointers, each an [n][B] iterate in
 * plan order; one of them is this rank's own, the others are peer-mapped over NVLink), so the exchange overlaps the
 * gather row by row instead of following it as a separate collective. The caller provides the peer mappings
 * (CUDA IPC) and a cross-rank barrier between updates (cy2path_b200/dist.py). */
int32_t cy2p_propagate_rows_allgather(cy2p_plan *plan, const float *d_Xp, void *const *d_peer_out, int32_t npeers,
                                      int32_t B, int32_t q_begin, int32_t q_end, void *stream);
int32_t cy2p_propagate_rows_allgather_f64(cy2p_plan *plan, const double *d_Xp, void *const *d_peer_out,
                                          int32_t npeers, int32_t B, int32_t q_begin, int32_t q_end, void *stream);

/* Halo push: like cy2p_propagate_rows_allgather, but a finished row q goes only to the peers that gather it:
 * d_row_peers[q] has bit g set when peer g owns a destination row with source q (or is the computing rank itself);
 * at most 32 peers. In the plan's locality order a rank's rows reference ~16-22% of the other ranks' rows, so the
 * exchange shrinks ~4x against the full all-gather (cy2path_b200/dist.py::propagate_row_partitioned_halo). */
int32_t cy2p_propagate_rows_push(cy2p_plan *plan, const float *d_Xp, void *const *d_peer_out, int32_t npeers,
                                 const uint32_t *d_row_peers, int32_t B, int32_t q_begin, int32_t q_end,
                                 void *stream);
int32_t cy2p_propagate_rows_push_f64(cy2p_plan *plan, const double *d_Xp, void *const *d_peer_out, int32_t npeers,
                                     const uint32_t *d_row_peers, int32_t B, int32_t q_begin, int32_t q_end,
                                     void *stream);
int32_t cy2p_enable_peer_access(int32_t device, int32_t peer);

/* Row-partitioned single-start run (B == 1), ALL `iters` updates in ONE cooperative launch: per update the rank's rows
 * [q_begin, q_end) are formed from its full-size iterate and pushed to the ranks that gather them (as
 * cy2p_propagate_rows_push), followed by a grid barrier, the cross-GPU flag exchange of cy2p_xbarrier done by one CTA,
 * and a second grid barrier — no launch and no host round trip per update. d_peer_buf0 / d_peer_buf1: device arrays of
 * `world` pointers to every rank's two ping-pong iterates [n] (update 0 reads buffer 0 of this rank); flags / epoch
 * counter / error word as for cy2p_xbarrier. All ranks must call it with the same `iters`. */
int32_t cy2p_propagate_rows_persist(cy2p_plan *plan, void *const *d_peer_buf0, void *const *d_peer_buf1, int32_t world,
                                    int32_t rank, const uint32_t *d_row_peers, int32_t q_begin, int32_t q_end,
                                    int32_t iters, void *const *d_peer_flags, int32_t *d_my_flags,
                                    int32_t *d_epoch_counter, int32_t *d_error, void *stream);
int32_t cy2p_propagate_rows_persist_f64(cy2p_plan *plan, void *const *d_peer_buf0, void *const *d_peer_buf1,
                                        int32_t world, int32_t rank, const uint32_t *d_row_peers, int32_t q_begin,
                                        int32_t q_end, int32_t iters, void *const *d_peer_flags, int32_t *d_my_flags,
                                        int32_t *d_epoch_counter, int32_t *d_error, void *stream);

/* Cross-GPU barrier between two updates of the row-partitioned run, as ONE small kernel on `stream` (so it orders
 * after the rank's update kernel and can be captured in a CUDA graph): every rank writes its epoch into its slot of
 * every rank's flag array (d_peer_flags: device array of `world` pointers to the ranks' int32[world] flag arrays in
 * peer-mapped memory, this rank's own at index `rank`) and waits until its own array (d_my_flags) shows that epoch in
 * every slot. The epoch is *d_epoch_counter + 1 (advanced by the kernel; all ranks must start from equal counters and
 * zeroed flags). A wait that lasts ~2 s gives up and sets *d_error = 1 instead of hanging the device. */
int32_t cy2p_xbarrier(void *const *d_peer_flags, int32_t *d_my_flags, int32_t *d_epoch_counter, int32_t rank,
                      int32_t world, int32_t *d_error, void *stream);

/* ---- K0/K2: chain sampling ------------------------------------------------------------------
 * cy2p_cdf_build replaces extract_nonzero_entries (simulation.py:15-53): idx int32 [n][W] and
 * cum float32 [n][W] = round3(float32 sequential cumsum of the row), zero padded; W from plan_info.
 * Bit-exact with the reference (fp32 adds in storage order, rint(x*1000)/1000 in fp32, no FMA). */
int32_t cy2p_cdf_build(const cy2p_plan *plan, int32_t *d_idx, float *d_cum, int32_t W, void *stream);

/* cy2p_sample_chains replaces the joblib fan-out of iterate_markov_chain (simulation.py:57-87,
 * :133-153) given the uniforms each chain consumes:
 *   d_uniforms [C][steps] float64, d_init_state [C] int32
 *   d_states   [C][steps+1] int32,  d_probs [C][steps] float32 (T[prev,next], duplicates summed)
 *   d_err      [4] int32 device scratch: on CDF overshoot {1, chain, step, state}
 * j = first index with (double)r <= (double)cum[state][j]  (numpy>=2 promotion), bit-exact paths.
 * Synchronises `stream` and returns CY2P_ERR_CDF_OVERSHOOT if any chain overshot. */
int32_t cy2p_sample_chains(const cy2p_plan *plan, const int32_t *d_idx, const float *d_cum, int32_t W,
                           const int32_t *d_init_state, const double *d_uniforms, int64_t C, int32_t steps,
                           int32_t *d_states, float *d_probs, int32_t *d_err, void *stream);

/* Empirical state history (simulation.py:156-162): hist[k][i] = float32(count_k(i) / C) for the
 * first `steps1` columns of d_states [C][ld_states]; unvisited entries are 0 (the reference leaves
 * them uninitialised). d_counts is int32 scratch [steps1][n]. */
int32_t cy2p_chain_history(const int32_t *d_states, int64_t C, int32_t ld_states, int32_t steps1, int32_t n,
                           int32_t *d_counts, float *d_hist, void *stream);

/* ---- K3: factorial latent dynamic model (FHMM) ------------------------------------------------
 * cy2p_fhmm_forward replaces FHMM.forward_model + log_transform_params (_FHMM.py:84-150,
 * methods.py:6-25). All tensors float32 device, C-contiguous:
 *   theta_pi [S][L], theta_w [L], theta_E [S][Le][N] (Le = 1 if restricted else L), theta_A [L][S][S]
 *   d_pred   [I][N] or NULL;  d_hidden [I][S][L];  d_joint [I][S][L][N] or NULL (4 GB at config 3)
 *   d_logE   [S][Le][N] out: log-softmax of theta_E (kept for the loss pass);
 *   d_small  out, [cy2p_fhmm_small_floats(S,L)] floats: log_pi, log_w, log_A, log_A_mix (see fhmm.cu)
 *   d_ws     >= cy2p_fhmm_ws_bytes(S,L,N,I) bytes */
size_t cy2p_fhmm_ws_bytes(int32_t S, int32_t L, int32_t N, int32_t I);
int32_t cy2p_fhmm_small_floats(int32_t S, int32_t L);
int32_t cy2p_fhmm_forward(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                          const float *theta_pi, const float *theta_w, const float *theta_E,
                          const float *theta_A, float *d_pred, float *d_hidden, float *d_joint,
                          float *d_logE, float *d_small, void *d_ws, size_t ws_bytes, void *stream);

/* cy2p_fhmm_loss_grad replaces one epoch's forward + loss + loss.backward() of trainer.train
 * (trainer.py:134-196) and compute_log_likelihood (methods.py:28-40):
 *   d_D      [I][N] float32 simulation history;  plan carries T (sparse) for the exclusivity and
 *            TPM regularisers (the reference multiplies by a densified T, trainer.py:175,190)
 *   weights  host float[4] = {sparsity, exclusivity, orthogonality, TPM}
 *   d_terms  [8] float32 out: loss, divergence, log_likelihood, sparsity, orthogonality,
 *            exclusivity, regularisation, (reserved)
 *   d_g*     gradients of `loss` w.r.t. the four unnormalised parameters (same shapes) */
int32_t cy2p_fhmm_loss_grad(cy2p_plan *plan, int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                            const float *theta_pi, const float *theta_w, const float *theta_E,
                            const float *theta_A, const float *d_D, const float *h_weights,
                            float *d_terms, float *d_gpi, float *d_gw, float *d_gE, float *d_gA,
                            void *d_ws, size_t ws_bytes, void *stream);

/* cy2p_fhmm_decode replaces FHMM.filtering / smoothing / viterbi (_FHMM.py:153-281) in one call:
 *   d_log_alpha [I][S][L], d_log_probs [I][L]           (filtering)
 *   d_log_beta [I][S][L], d_log_gamma [I][S][L]         (smoothing; uses the last row of D at every step, as
 *                                                        the reference does, _FHMM.py:209,223)
 *   d_log_delta [I][S][L], d_psi [I][S][L] (float), d_log_max [I][L], d_best_path [L][I] (float)   (viterbi)
 * d_ws: the same work space as cy2p_fhmm_forward. */
int32_t cy2p_fhmm_decode(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                         const float *theta_w, const float *theta_E, const float *theta_A, const float *d_D,
                         float *d_log_alpha, float *d_log_probs, float *d_log_beta, float *d_log_gamma,
                         float *d_log_delta, float *d_psi, float *d_log_max, float *d_best_path, void *d_ws,
                         size_t ws_bytes, void *stream);

/* ---- K3b: latent state-space model with state-level lineage weights (LSSM) ----------------------
 * Same call shapes as the FHMM entry points; replaces LSSM.forward_model (_LSSM.py:83-117), one epoch of
 * trainer.train on an LSSM (trainer.py:134-196; the trainer is shared by the model classes) and
 * LSSM.filtering / smoothing / viterbi (_LSSM.py:120-260). Parameter shapes differ:
 *   theta_pi [S], theta_w [S][L], theta_E [S][Le][N], theta_A [S][S];  d_hidden [I][S];
 *   d_small  out, [cy2p_lssm_small_floats(S,L)] floats: log_pi [S], log_w [S][L], log_A [S][S]
 *   gradients d_gpi [S], d_gw [S][L], d_gE [S][Le][N], d_gA [S][S];  d_ws >= cy2p_lssm_ws_bytes(S,L,N,I) */
size_t cy2p_lssm_ws_bytes(int32_t S, int32_t L, int32_t N, int32_t I);
int32_t cy2p_lssm_small_floats(int32_t S, int32_t L);
int32_t cy2p_lssm_forward(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                          const float *theta_pi, const float *theta_w, const float *theta_E,
                          const float *theta_A, float *d_pred, float *d_hidden, float *d_joint,
                          float *d_logE, float *d_small, void *d_ws, size_t ws_bytes, void *stream);
int32_t cy2p_lssm_loss_grad(cy2p_plan *plan, int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                            const float *theta_pi, const float *theta_w, const float *theta_E,
                            const float *theta_A, const float *d_D, const float *h_weights,
                            float *d_terms, float *d_gpi, float *d_gw, float *d_gE, float *d_gA,
                            void *d_ws, size_t ws_bytes, void *stream);
int32_t cy2p_lssm_decode(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                         const float *theta_w, const float *theta_E, const float *theta_A, const float *d_D,
                         float *d_log_alpha, float *d_log_probs, float *d_log_beta, float *d_log_gamma,
                         float *d_log_delta, float *d_psi, float *d_log_max, float *d_best_path, void *d_ws,
                         size_t ws_bytes, void *stream);

/* ---- K3c: latent state-space model with node-level lineage weights (NSSM) -----------------------
 * Replaces NSSM.forward_model (_NSSM.py:76-118), one epoch of trainer.train on an NSSM and
 * NSSM.filtering / smoothing / viterbi (_NSSM.py:120-249). Parameter shapes:
 *   theta_pi [S], theta_cw [Sc][N][L] (Sc = 1 if restricted else S), theta_E [S][N], theta_A [S][S]
 *   d_hidden [I][S]; d_joint [I][S][L][N] (the layout forward_model returns after its permute, _NSSM.py:111);
 *   d_logE [S][N] and d_logcw [Sc][N][L]: the two log-softmaxes; d_small: the LSSM layout with
 *   log_w = log r[s][l], r[s][l] = sum_n E[s][n] cw[s'][n][l] (cy2p_lssm_small_floats(S,L) floats)
 *   gradients d_gpi [S], d_gcw [Sc][N][L], d_gE [S][N], d_gA [S][S];  requires S*L <= 64. */
size_t cy2p_nssm_ws_bytes(int32_t S, int32_t L, int32_t N, int32_t I);
int32_t cy2p_nssm_forward(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                          const float *theta_pi, const float *theta_cw, const float *theta_E,
                          const float *theta_A, float *d_pred, float *d_hidden, float *d_joint,
                          float *d_logE, float *d_logcw, float *d_small, void *d_ws, size_t ws_bytes,
                          void *stream);
int32_t cy2p_nssm_loss_grad(cy2p_plan *plan, int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                            const float *theta_pi, const float *theta_cw, const float *theta_E,
                            const float *theta_A, const float *d_D, const float *h_weights,
                            float *d_terms, float *d_gpi, float *d_gcw, float *d_gE, float *d_gA,
                            void *d_ws, size_t ws_bytes, void *stream);
int32_t cy2p_nssm_decode(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                         const float *theta_cw, const float *theta_E, const float *theta_A, const float *d_D,
                         float *d_log_alpha, float *d_log_probs, float *d_log_beta, float *d_log_gamma,
                         float *d_log_delta, float *d_psi, float *d_log_max, float *d_best_path, void *d_ws,
                         size_t ws_bytes, void *stream);

/* ---- K4: pairwise distance matrices between sampled trajectories ------------------------------------
 * Replace dtaidistance.dtw_ndim.distance_matrix_fast (cytopath.py:92-94) and compute_Hausdorff_distance
 * (cytopath.py:30-45) of the lineage clustering. d_series: float64 device [C][len][dpad], the point dimension
 * zero-padded to dpad (a multiple of 8, <= 64), 16-byte aligned; d_out: float64 [C][C], symmetric, zero diagonal.
 *   DTW: squared-Euclidean local cost, D[i][j] = cost + min(D[i-1][j], D[i][j-1], D[i-1][j-1]), sqrt at the end.
 *   Hausdorff: max(max_a min_b |a-b|, max_b min_a |a-b|), Euclidean.
 * d_ws >= cy2p_traj_ws_bytes(C, len) bytes. */
size_t cy2p_traj_ws_bytes(int32_t C, int32_t len);
int32_t cy2p_dtw_matrix(const double *d_series, int32_t C, int32_t len, int32_t dpad, double *d_out, void *d_ws,
                        size_t ws_bytes, void *stream);
int32_t cy2p_hausdorff_matrix(const double *d_series, int32_t C, int32_t len, int32_t dpad, double *d_out,
                              void *d_ws, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CY2PATH_B200_H */
