// K2 — Markov-chain sampling by lane-cooperative inverse-CDF search (reference simulation.py:57-87)
// and K2b — the per-step empirical state histogram (reference simulation.py:156-162).
//
// A group of LANES (8, 16 or 32) lanes walks one chain, so a warp carries 32/LANES chains. Per step the
// group reads the CDF row and the index row of the current state LANES columns at a time (coalesced),
// compares the step's uniform against the row in float64 — numpy>=2 promotes `np.float64 <= float32[]`
// to float64 — and the first set bit of the group's ballot is the reference's
// `np.where(...)[0][0]`; the search stops at the first chunk with a hit. Uniforms are fetched LANES steps at a
// time and states/probabilities written LANES steps at a time, so the streaming traffic is coalesced although
// the walk itself is a chain of dependent loads. The first ncu capture showed the 32-lane version to be
// instruction-issue bound (82% issue slots): the narrow groups divide the per-step scalar work (shuffles,
// ballot, address arithmetic) among 2-4 chains per warp.
//
// T[prev, next]: for matrices whose rows have strictly increasing columns (scipy-canonical; checked at plan
// creation) column j of the table IS entry j of the CSR row, so the probability is data[indptr[prev] + j].
// Otherwise (duplicate / unsorted columns) the row is scanned and duplicates are summed in storage order,
// which is what scipy's scalar indexing returns.
#include <stdlib.h>

#include "common.cuh"

namespace cy2p {

template <int LANES, bool CANONICAL>
__global__ void __launch_bounds__(256)
sample_chains_kernel(int64_t C, int32_t steps, int32_t W, const int32_t *__restrict__ idx,
                     const float *__restrict__ cum, const int32_t *__restrict__ indptr,
                     const int32_t *__restrict__ indices, const float *__restrict__ data,
                     const int32_t *__restrict__ init_state, const double *__restrict__ uniforms,
                     int32_t *__restrict__ states, float *__restrict__ probs, int32_t *err) {
    constexpr int GROUPS = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LANES;             // lane within the group
    const int gshift = (lane / LANES) * LANES;  // first lane of the group
    const unsigned gmask = (unsigned)((LANES == 32 ? 0xffffffffull : ((1ull << LANES) - 1ull)) << gshift);
    const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t chain = warp_global * GROUPS + lane / LANES;
    const bool live = chain < C;
    const int64_t ch = live ? chain : 0;
    const double *u = uniforms + ch * (int64_t)steps;
    int32_t *out_s = states + ch * ((int64_t)steps + 1);
    float *out_p = probs + ch * (int64_t)steps;

    int32_t state = init_state[ch];
    if (live && gl == 0) out_s[0] = state;
    bool dead = !live;

    for (int32_t k0 = 0; k0 < steps; k0 += LANES) {
        const int32_t m = min(LANES, steps - k0);
        const double my_u = (live && gl < m) ? u[k0 + gl] : 0.0;
        int32_t my_state = -1;
        float my_prob = 0.f;
        for (int32_t s = 0; s < m; ++s) {
            const double r = __shfl_sync(0xffffffffu, my_u, gshift + s);
            const int64_t base = (int64_t)state * W;
            int32_t nxt = -1, col = 0;
            // chunked first-hit search; all groups of the warp iterate together until each has finished
            bool searching = !dead;
            for (int32_t c0 = 0; c0 < W; c0 += LANES) {
                const int32_t c = c0 + gl;
                const bool in = searching && c < W;
                const float cv = in ? __ldg(cum + base + c) : 0.f;
                const int32_t iv = in ? __ldg(idx + base + c) : 0;
                const unsigned hit = (__ballot_sync(0xffffffffu, in && (r <= (double)cv)) & gmask) >> gshift;
                const int first = hit ? __ffs(hit) - 1 : 0;
                const int32_t got = __shfl_sync(0xffffffffu, iv, gshift + first);
                if (searching && hit) {
                    nxt = got;
                    col = c0 + first;
                    searching = false;
                }
                if (!__any_sync(0xffffffffu, searching)) break;
            }
            if (!dead && nxt < 0) {  // r above every CDF entry of the row: the reference raises IndexError here
                if (gl == 0 && atomicCAS(&err[0], 0, 1) == 0) {
                    err[1] = (int32_t)min(chain, (int64_t)INT32_MAX);
                    err[2] = k0 + s + 1;
                    err[3] = state;
                }
                dead = true;
            }
            float pr = 0.f;
            if (CANONICAL) {
                if (!dead && gl == 0) pr = __ldg(data + __ldg(indptr + state) + col);
                pr = __shfl_sync(0xffffffffu, pr, gshift);
            } else {
                // sum of the stored entries of row `prev` whose column is `next`, in storage order
                // (one chain per warp on this path, so the loops below are warp-uniform)
                static_assert(CANONICAL || LANES == 32, "the duplicate-summing path runs one chain per warp");
                const int32_t rb = dead ? 0 : __ldg(indptr + state), re = dead ? 0 : __ldg(indptr + state + 1);
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
            }
            if (!dead) {
                state = nxt;
                if (gl == s) {
                    my_state = nxt;
                    my_prob = pr;
                }
            }
        }
        if (live && gl < m) {
            out_s[k0 + 1 + gl] = my_state;  // -1 marks steps after an overshoot
            out_p[k0 + gl] = my_prob;
        }
    }
}

// Canonical rows no longer than LANES * NCH: the whole CDF / index row of the current state (NCH loads of each per
// lane) and the row pointer are requested at once, so a step costs two dependent L2 round trips (row, then T[prev,next])
// instead of one per chunk plus two; with 8 lanes per chain a wave of resident threads also carries twice the chains of
// the 16-lane chunked search. Measured at config 2 (100,000 chains x 2,000 steps, W = 30): see profiles/README_r02.md.
template <int LANES, int NCH>
__global__ void __launch_bounds__(256)
sample_chains_prefetch(int64_t C, int32_t steps, int32_t W, const int32_t *__restrict__ idx,
                       const float *__restrict__ cum, const int32_t *__restrict__ indptr,
                       const float *__restrict__ data, const int32_t *__restrict__ init_state,
                       const double *__restrict__ uniforms, int32_t *__restrict__ states, float *__restrict__ probs,
                       int32_t *err) {
    constexpr int GROUPS = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LANES;
    const int gshift = (lane / LANES) * LANES;
    const unsigned gmask = (unsigned)((LANES == 32 ? 0xffffffffull : ((1ull << LANES) - 1ull)) << gshift);
    const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t chain = warp_global * GROUPS + lane / LANES;
    const bool live = chain < C;
    const int64_t ch = live ? chain : 0;
    const double *u = uniforms + ch * (int64_t)steps;
    int32_t *out_s = states + ch * ((int64_t)steps + 1);
    float *out_p = probs + ch * (int64_t)steps;

    int32_t state = init_state[ch];
    if (live && gl == 0) out_s[0] = state;
    bool dead = !live;

    for (int32_t k0 = 0; k0 < steps; k0 += LANES) {
        const int32_t m = min(LANES, steps - k0);
        const double my_u = (live && gl < m) ? u[k0 + gl] : 0.0;
        int32_t my_state = -1;
        float my_prob = 0.f;
        for (int32_t s = 0; s < m; ++s) {
            const double r = __shfl_sync(0xffffffffu, my_u, gshift + s);
            const int64_t base = (int64_t)state * W;
            float cv[NCH];
            int32_t iv[NCH];
            const int32_t rp = dead ? 0 : __ldg(indptr + state);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const int32_t c = j * LANES + gl;
                const bool in = !dead && c < W;
                cv[j] = in ? __ldg(cum + base + c) : -1.f;  // uniforms are >= 0: a padding lane never hits
                iv[j] = in ? __ldg(idx + base + c) : 0;
            }
            int32_t nxt = -1, col = 0;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const unsigned hit = (__ballot_sync(0xffffffffu, cv[j] >= 0.f && r <= (double)cv[j]) & gmask) >> gshift;
                const int first = hit ? __ffs(hit) - 1 : 0;
                const int32_t got = __shfl_sync(0xffffffffu, iv[j], gshift + first);
                if (nxt < 0 && hit) {
                    nxt = got;
                    col = j * LANES + first;
                }
            }
            if (!dead && nxt < 0) {  // r above every CDF entry of the row: the reference raises IndexError here
                if (gl == 0 && atomicCAS(&err[0], 0, 1) == 0) {
                    err[1] = (int32_t)min(chain, (int64_t)INT32_MAX);
                    err[2] = k0 + s + 1;
                    err[3] = state;
                }
                dead = true;
            }
            float pr = 0.f;
            if (!dead && gl == 0) pr = __ldg(data + rp + col);
            pr = __shfl_sync(0xffffffffu, pr, gshift);
            if (!dead) {
                state = nxt;
                if (gl == s) {
                    my_state = nxt;
                    my_prob = pr;
                }
            }
        }
        if (live && gl < m) {
            out_s[k0 + 1 + gl] = my_state;  // -1 marks steps after an overshoot
            out_p[k0 + gl] = my_prob;
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

template <int LANES, bool CANONICAL>
static void launch_sampler(int64_t C, int32_t steps, int32_t W, const int32_t *idx, const float *cum,
                           const cy2p_plan *plan, const int32_t *init_state, const double *uniforms, int32_t *states,
                           float *probs, int32_t *err, cudaStream_t st) {
    const int warps = 8, groups = 32 / LANES;
    const int64_t chains_per_cta = (int64_t)warps * groups;
    const unsigned grid = (unsigned)((C + chains_per_cta - 1) / chains_per_cta);
    count_launch();
    sample_chains_kernel<LANES, CANONICAL><<<grid, warps * 32, 0, st>>>(C, steps, W, idx, cum, plan->d_indptr,
                                                                       plan->d_indices, plan->d_data, init_state,
                                                                       uniforms, states, probs, err);
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
        // the fast probability lookup needs table column j == CSR entry j: canonical rows no longer than W
        const bool canonical = plan->canonical_rows && plan->max_out_deg <= W;
        const char *ev = getenv("CY2P_SAMPLER_LANES");  // tuning knob
        int lanes = (ev && *ev) ? atoi(ev) : (W <= 8 ? 8 : 16);
#define CY2P_SAMPLE(L_, C_) \
    launch_sampler<L_, C_>(C, steps, W, d_idx, d_cum, plan, d_init_state, d_uniforms, d_states, d_probs, d_err, st)
#define CY2P_PREFETCH(L_, N_)                                                                                       \
    do {                                                                                                            \
        const int64_t per_cta = 8 * (32 / L_);                                                                      \
        count_launch();                                                                                             \
        sample_chains_prefetch<L_, N_><<<(unsigned)((C + per_cta - 1) / per_cta), 256, 0, st>>>(                    \
            C, steps, W, d_idx, d_cum, plan->d_indptr, plan->d_data, d_init_state, d_uniforms, d_states, d_probs,   \
            d_err);                                                                                                 \
    } while (0)
        const char *pv = getenv("CY2P_SAMPLER_PREFETCH");  // 0: the chunked search of round 1 (measurement)
        const bool prefetch = canonical && !(ev && *ev) && !(pv && *pv && atoi(pv) == 0);
        const char *lv = getenv("CY2P_SAMPLER_PF_LANES");  // tuning knob: lanes per chain of the prefetching kernel
        const int pf_lanes = (lv && *lv) ? atoi(lv) : 8;
        if (prefetch && W <= 32 && pf_lanes == 4) CY2P_PREFETCH(4, 8);
        else if (prefetch && W <= 32 && pf_lanes == 16) CY2P_PREFETCH(16, 2);
        else if (prefetch && W <= 32) CY2P_PREFETCH(8, 4);
        else if (prefetch && W <= 64) CY2P_PREFETCH(8, 8);
        else if (!canonical) CY2P_SAMPLE(32, false);
        else if (lanes <= 8) CY2P_SAMPLE(8, true);
        else if (lanes <= 16) CY2P_SAMPLE(16, true);
        else CY2P_SAMPLE(32, true);
#undef CY2P_SAMPLE
#undef CY2P_PREFETCH
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
    count_launch();
    hist_count_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(C, ld_states, steps1, n, d_states, d_counts);
    count_launch();
    hist_normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, C, d_counts, d_hist);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"
