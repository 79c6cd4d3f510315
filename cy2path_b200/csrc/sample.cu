// K2 — Markov-chain sampling by warp-cooperative inverse-CDF search (reference simulation.py:57-87)
// and K2b — the per-step empirical state histogram (reference simulation.py:156-162).
//
// One warp walks one chain. Per step the warp reads the CDF row and the index row of the current
// state with one coalesced load each (lane = table column), compares the step's uniform against
// the row in float64 — numpy>=2 promotes `np.float64 <= float32[]` to float64 —
// and the first set bit of the ballot is the reference's `np.where(...)[0][0]`. Rows wider than 32
// are searched in 32-column chunks, stopping at the first chunk with a hit. Uniforms are fetched 32
// steps at a time (one coalesced 256-byte read per warp) and states/probabilities are written back
// 32 steps at a time, so the streaming traffic is coalesced although the walk itself is a chain of
// dependent loads.
#include "common.cuh"

namespace cy2p {

__global__ void __launch_bounds__(256)
sample_chains_kernel(int64_t C, int32_t steps, int32_t W, const int32_t *__restrict__ idx,
                     const float *__restrict__ cum, const int32_t *__restrict__ indptr,
                     const int32_t *__restrict__ indices, const float *__restrict__ data,
                     const int32_t *__restrict__ init_state, const double *__restrict__ uniforms,
                     int32_t *__restrict__ states, float *__restrict__ probs, int32_t *err) {
    const int lane = threadIdx.x & 31;
    const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= C) return;
    const double *u = uniforms + chain * (int64_t)steps;
    int32_t *out_s = states + chain * ((int64_t)steps + 1);
    float *out_p = probs + chain * (int64_t)steps;

    int32_t state = init_state[chain];
    if (lane == 0) out_s[0] = state;
    bool dead = false;

    for (int32_t k0 = 0; k0 < steps; k0 += 32) {
        const int32_t m = min(32, steps - k0);
        const double my_u = (lane < m) ? u[k0 + lane] : 0.0;
        int32_t my_state = -1;
        float my_prob = 0.f;
        for (int32_t s = 0; s < m && !dead; ++s) {
            const double r = __shfl_sync(0xffffffffu, my_u, s);
            const int64_t base = (int64_t)state * W;
            int32_t nxt = -1;
            for (int32_t c0 = 0; c0 < W; c0 += 32) {
                const int32_t c = c0 + lane;
                const bool in = c < W;
                const float cv = in ? __ldg(cum + base + c) : 0.f;
                const int32_t iv = in ? __ldg(idx + base + c) : 0;
                const unsigned hit = __ballot_sync(0xffffffffu, in && (r <= (double)cv));
                if (hit) {
                    nxt = __shfl_sync(0xffffffffu, iv, __ffs(hit) - 1);
                    break;
                }
            }
            if (nxt < 0) {  // r above every CDF entry of the row: the reference raises IndexError here
                if (lane == 0 && atomicCAS(&err[0], 0, 1) == 0) {
                    err[1] = (int32_t)min(chain, (int64_t)INT32_MAX);
                    err[2] = k0 + s + 1;
                    err[3] = state;
                }
                dead = true;
                break;
            }
            // T[prev, next]: sum of the stored entries of row `prev` whose column is `next`, in
            // storage order (scipy scalar indexing sums duplicates)
            const int32_t rb = __ldg(indptr + state), re = __ldg(indptr + state + 1);
            float pr = 0.f;
            for (int32_t p0 = rb; p0 < re; p0 += 32) {
                const int32_t p = p0 + lane;
                const bool match = p < re && __ldg(indices + p) == nxt;
                const float v = match ? __ldg(data + p) : 0.f;
                unsigned mm = __ballot_sync(0xffffffffu, match);
                while (mm) {  // ascending entry order
                    const int src = __ffs(mm) - 1;
                    pr = __fadd_rn(pr, __shfl_sync(0xffffffffu, v, src));
                    mm &= mm - 1;
                }
            }
            state = nxt;
            if (lane == s) {
                my_state = nxt;
                my_prob = pr;
            }
        }
        if (lane < m) {
            out_s[k0 + 1 + lane] = my_state;  // -1 marks steps after an overshoot
            out_p[k0 + lane] = my_prob;
        }
    }
}

__global__ void hist_count_kernel(int64_t C, int32_t ld_states, int32_t steps1, int32_t n,
                                  const int32_t *__restrict__ states, int32_t *__restrict__ counts) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * steps1) return;
    const int64_t c = t / steps1;
    const int32_t k = (int32_t)(t % steps1);
    const int32_t s = states[c * ld_states + k];
    if (s >= 0 && s < n) atomicAdd(&counts[(int64_t)k * n + s], 1);
}

// counts / counts.sum() in float64 (numpy int64 / int64 true division), stored to float32
__global__ void hist_normalise_kernel(int64_t total, int64_t C, const int32_t *__restrict__ counts,
                                      float *__restrict__ hist) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    hist[t] = (float)((double)counts[t] / (double)C);
}

}  // namespace cy2p

using namespace cy2p;

extern "C" {

int32_t cy2p_sample_chains(const cy2p_plan *plan, const int32_t *d_idx, const float *d_cum, int32_t W,
                           const int32_t *d_init_state, const double *d_uniforms, int64_t C, int32_t steps,
                           int32_t *d_states, float *d_probs, int32_t *d_err, void *stream) {
    CY2P_REQUIRE(plan && d_idx && d_cum && d_init_state && d_states && d_err, "null pointer");
    CY2P_REQUIRE(C >= 0 && steps >= 0 && W >= 1, "C >= 0, steps >= 0, W >= 1 required");
    CY2P_REQUIRE(steps == 0 || (d_uniforms && d_probs), "null uniforms / probs");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    CY2P_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t) * 4, st));
    if (C > 0) {
        const int warps = 8;
        const unsigned grid = (unsigned)((C + warps - 1) / warps);
        count_launch(); sample_chains_kernel<<<grid, warps * 32, 0, st>>>(C, steps, W, d_idx, d_cum, plan->d_indptr, plan->d_indices,
                                                         plan->d_data, d_init_state, d_uniforms, d_states, d_probs,
                                                         d_err);
        CY2P_CUDA(cudaGetLastError());
    }
    int32_t h[4] = {0, 0, 0, 0};
    CY2P_CUDA(cudaMemcpyAsync(h, d_err, sizeof(h), cudaMemcpyDeviceToHost, st));
    CY2P_CUDA(cudaStreamSynchronize(st));
    if (h[0]) {
        set_error("CDF overshoot: chain %d step %d state %d (uniform above the last CDF entry)", h[1], h[2], h[3]);
        return CY2P_ERR_CDF_OVERSHOOT;
    }
    return CY2P_OK;
}

int32_t cy2p_chain_history(const int32_t *d_states, int64_t C, int32_t ld_states, int32_t steps1, int32_t n,
                           int32_t *d_counts, float *d_hist, void *stream) {
    CY2P_REQUIRE(d_states && d_counts && d_hist, "null pointer");
    CY2P_REQUIRE(C >= 1 && steps1 >= 1 && steps1 <= ld_states && n >= 1, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)steps1 * n;
    CY2P_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * (size_t)total, st));
    const int64_t work = C * steps1;
    count_launch(); hist_count_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(C, ld_states, steps1, n, d_states, d_counts);
    count_launch(); hist_normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, C, d_counts, d_hist);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"
