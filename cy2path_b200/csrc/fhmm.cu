// K3 — latent dynamic models: forward pass, training loss and ANALYTIC gradients, decoders.
// FHMM first (the north_star's model); LSSM and NSSM (single latent chain with state- / node-level lineage weights,
// reference models/_LSSM.py, _NSSM.py) reuse its kernels and are found at the end of the device section.
//
// Replaces FHMM.forward_model + log_transform_params (reference models/_FHMM.py:84-150,
// models/methods.py:6-25), compute_log_likelihood (methods.py:28-40) and one epoch's loss + autograd
// backward of trainer.train (models/trainer.py:134-196).
//
// The reference materialises joint[I,S,L,N] (4 GB at BASELINE config 3) and differentiates through I
// in-place slice writes (O(I^2 S L N) backward). The joint factorises,
//     joint[t,s,l,n] = logE[s,e(l),n] + hidden[t,s,l] + log w[l]          (_FHMM.py:126-141)
// so everything the loss needs is   pred[t,n] = log sum_k G[t,k] E[k,n]   with the small matrix
// G[t,k] = w_l H_t[s,l] (k = s when restricted, (s,l) otherwise) and the time average
// C[s,l] = w_l mean_t H_t[s,l]. One pass over D produces pred, the KL divergence, dLoss/dlogE and
// dLoss/dG together (D is read once); the S x L recurrences run forward and backward inside one CTA.
//
// Notation: K = S*Le rows of the emission matrix (Le = 1 restricted, L otherwise); e(l) = 0 or l.
#include <float.h>

#include "common.cuh"

namespace cy2p {

constexpr int kMaxS = 32, kMaxL = 16, kMaxK = 64;  // one warp per lineage, lane = latent state

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// --------------------------------------------------------------------------------------------------
// log_softmax over the N nodes of every emission row (methods.py:17-19), two launches, grid (row, chunk):
//   pass 1: per-chunk (max, sum exp(x - max)) partials;  pass 2: combine the row's partials, write
//   logE = x - lse and accumulate row_sum[row] += sum exp(logE) in float64 (log of it is the ~1e-7 residual the
//   reference's compute_log_likelihood subtracts, methods.py:33-34).
constexpr int kSoftmaxChunk = 4096;

__global__ void __launch_bounds__(256) emission_softmax_partials(int N, const float *__restrict__ theta,
                                                                 float2 *__restrict__ partials) {
    const int row = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
    const float *x = theta + (size_t)row * N;
    const int i0 = chunk * kSoftmaxChunk, i1 = min(N, i0 + kSoftmaxChunk);
    __shared__ float sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m = -INFINITY;
    for (int i = i0 + threadIdx.x; i < i1; i += 256) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    m = sm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    __syncthreads();
    float z = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += 256) z += expf(x[i] - m);
    z = warp_sum(z);
    if (lane == 0) sm[warp] = z;
    __syncthreads();
    if (threadIdx.x == 0) {
        z = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) z += sm[w];
        partials[row * nchunks + chunk] = make_float2(m, z);
    }
}

__global__ void __launch_bounds__(256) emission_softmax_write(int N, const float *__restrict__ theta,
                                                              const float2 *__restrict__ partials,
                                                              float *__restrict__ logE, double *__restrict__ row_sum,
                                                              float *__restrict__ row_lse_out = nullptr) {
    const int row = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
    float m = -INFINITY;
    for (int c = 0; c < nchunks; ++c) m = fmaxf(m, partials[row * nchunks + c].x);
    double z = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const float2 pz = partials[row * nchunks + c];
        z += (double)pz.y * (double)expf(pz.x - m);
    }
    const float lse = m + (float)log(z);
    if (row_lse_out && chunk == 0 && threadIdx.x == 0) row_lse_out[row] = lse;
    const float *x = theta + (size_t)row * N;
    float *y = logE + (size_t)row * N;
    const int i0 = chunk * kSoftmaxChunk, i1 = min(N, i0 + kSoftmaxChunk);
    double r = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
        const float v = x[i] - lse;
        y[i] = v;
        r += (double)expf(v);
    }
    r = warp_sum(r);
    __shared__ double sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        r = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += sm[w];
        atomicAdd(&row_sum[row], r);
    }
}

// --------------------------------------------------------------------------------------------------
// Small forward (one CTA): the three small log_softmaxes, the mixed TPM, the hidden-state recurrence
// (one warp per lineage, float64 probability domain; hidden = log H afterwards), G and C.
//   small layout: log_pi[S*L] | log_w[L] | log_A[L*S*S] | log_Amix[S*S]
//   aux layout (floats): Hbar[S*L] | logC[S*L] | logc[S] | logq[L] | sparsity[1]
struct SmallOffsets {
    int log_pi, log_w, log_A, log_Amix, total;
    __host__ __device__ SmallOffsets(int S, int L) {
        log_pi = 0;
        log_w = log_pi + S * L;
        log_A = log_w + L;
        log_Amix = log_A + L * S * S;
        total = log_Amix + S * S;
    }
};
struct AuxOffsets {
    int Hbar, logC, logc, logq, sparsity, total;
    __host__ __device__ AuxOffsets(int S, int L) {
        Hbar = 0;
        logC = Hbar + S * L;
        logc = logC + S * L;
        logq = logc + S;
        sparsity = logq + L;
        total = sparsity + 4;
    }
};

// hf_smem != 0: the launch reserved I*S*L more floats of shared memory; H is kept there as float for the time average
// and the mixing matrix (both used to re-read Hlin from global memory in loops of dependent L2 round trips).
__global__ void __launch_bounds__(1024) fhmm_small_forward(int S, int L, int I, int restricted, int KP, int P,
                                                           const float *__restrict__ theta_pi,
                                                           const float *__restrict__ theta_w,
                                                           const float *__restrict__ theta_A, float *__restrict__ small,
                                                           float *__restrict__ hidden, float *__restrict__ G,
                                                           float *__restrict__ aux, double *__restrict__ Hlin,
                                                           double *__restrict__ Apow, int hf_smem,
                                                           double *__restrict__ Xblk, int *__restrict__ tickets) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    float *log_pi = small + so.log_pi, *log_w = small + so.log_w, *log_A = small + so.log_A,
          *log_Amix = small + so.log_Amix;
    extern __shared__ float sm[];
    float *s_logA = sm;                   // [L][S][S]
    float *s_h = s_logA + L * S * S;      // [2][S][L]
    float *s_logw = s_h + 2 * S * L;      // [L]
    float *s_amix = s_logw + L;           // [S][S]
    const int tid = threadIdx.x, nt = blockDim.x;

    // last-CTA tickets of the multi-CTA recurrence kernels that follow in this call (forward expand, backward blocks,
    // backward expand): zeroed here, by the first kernel of every forward / loss+grad call, whichever variants run later
    if (tickets != nullptr && tid < 4) tickets[tid] = 0;
    // the three raw parameter blocks come in with one sweep of independent loads; the softmaxes then run out of shared
    // memory (their max / sum loops would otherwise wait for one L2 round trip per element)
    for (int i = tid; i < L * S * S; i += nt) s_logA[i] = theta_A[i];
    for (int i = tid; i < S * L; i += nt) s_h[i] = theta_pi[i];
    for (int i = tid; i < L; i += nt) s_logw[i] = theta_w[i];
    __syncthreads();
    // log_softmax(theta_pi, dim=0): per lineage l over states
    for (int l = tid; l < L; l += nt) {
        float m = -INFINITY;
        for (int s = 0; s < S; ++s) m = fmaxf(m, s_h[s * L + l]);
        float z = 0.f;
        for (int s = 0; s < S; ++s) z += expf(s_h[s * L + l] - m);
        const float lse = m + logf(z);
        for (int s = 0; s < S; ++s) log_pi[s * L + l] = s_h[s * L + l] - lse;
    }
    if (tid == 32) {  // (a lane of the second warp: the first is busy with theta_pi)
        float m = -INFINITY;
        for (int l = 0; l < L; ++l) m = fmaxf(m, s_logw[l]);
        float z = 0.f;
        for (int l = 0; l < L; ++l) z += expf(s_logw[l] - m);
        const float lse = m + logf(z);
        for (int l = 0; l < L; ++l) {
            const float v = s_logw[l] - lse;
            log_w[l] = v;
            s_logw[l] = v;
        }
    }
    for (int r = nt - 1 - tid; r < L * S; r += nt) {  // rows (l, s) over destination state, from the far end of the CTA
        float *x = s_logA + (size_t)r * S;
        float m = -INFINITY;
        for (int j = 0; j < S; ++j) m = fmaxf(m, x[j]);
        float z = 0.f;
        for (int j = 0; j < S; ++j) z += expf(x[j] - m);
        const float lse = m + logf(z);
        for (int j = 0; j < S; ++j) {
            const float v = x[j] - lse;
            log_A[(size_t)r * S + j] = v;
            x[j] = v;
        }
    }
    __syncthreads();
    // log A_mix[s,s'] = lse_l(log A[l,s,s'] + log w[l])   (_FHMM.py:93-96); sparsity = -mean diag (trainer.py:148)
    for (int e = tid; e < S * S; e += nt) {
        float m = -INFINITY;
        for (int l = 0; l < L; ++l) m = fmaxf(m, s_logA[l * S * S + e] + s_logw[l]);
        float z = 0.f;
        for (int l = 0; l < L; ++l) z += expf(s_logA[l * S * S + e] + s_logw[l] - m);
        log_Amix[e] = m + logf(z);
        s_amix[e] = m + logf(z);
    }
    __syncthreads();
    if (tid == 0) {
        float d = 0.f;
        for (int s = 0; s < S; ++s) d += s_amix[s * S + s];
        aux[ao.sparsity] = -d / (float)S;
    }

    // Hidden recurrence H_t[., l] = H_{t-1}[., l] . A_l in the probability domain, float64 (the reference iterates
    // lse_s(hidden[t-1][s,l] + log A[l,s,s']) in float32, _FHMM.py:116-124). A is constant over t, so the I-step
    // chain is cut into blocks of P steps:
    //   level 1  A^1 .. A^P                      (P-1 small matrix products, block-parallel)
    //   level 2  X_b = X_{b-1} . A^P, X_0 = pi   (ceil(I/P) dependent steps, one warp per lineage)
    //   level 3  H_{bP+j} = X_b . A^j            (all (t,s,l) in parallel)
    // which leaves ~P + I/P dependent steps instead of I.
    const int SL = S * L, SS = S * S;
    const int warp = tid >> 5, lane = tid & 31;
    const int NB = (I + P - 1) / P;
    double *s_pow = reinterpret_cast<double *>(sm + ((L * SS + 2 * SL + L + SS + 1) & ~1));  // [P][L][S][S]: A^(j+1)
    double *s_X = s_pow + (size_t)P * L * SS;                                         // [NB][S][L] (unused if P == 1)
    double *s_part = s_X + (P > 1 ? (size_t)NB * SL : 0);                              // [blockDim]
    float *s_Hf = reinterpret_cast<float *>(s_part + nt);                              // [I][S][L] when hf_smem
    for (int i = tid; i < L * SS; i += nt) {
        const double a = exp((double)s_logA[i]);
        s_pow[i] = a;
        Apow[i] = a;
    }
    __syncthreads();
    for (int j = 1; j < P; ++j) {  // A^(j+1) = A^j . A
        const double *prev = s_pow + (size_t)(j - 1) * L * SS;
        double *cur = s_pow + (size_t)j * L * SS;
        for (int i = tid; i < L * SS; i += nt) {
            const int l = i / SS, r = (i / S) % S, c = i % S;
            double acc = 0.0;
            for (int k = 0; k < S; ++k) acc = fma(prev[(l * S + r) * S + k], s_pow[(l * S + k) * S + c], acc);
            cur[i] = acc;
            Apow[(size_t)j * L * SS + i] = acc;
        }
        __syncthreads();
    }
    if (warp < L) {  // level 2
        const int l = warp;
        const double *AP = s_pow + (size_t)(P - 1) * L * SS + (size_t)l * SS;
        double h = lane < S ? exp((double)log_pi[lane * L + l]) : 0.0;
        for (int b = 0; b < NB; ++b) {
            if (lane < S) (P > 1 ? s_X : Hlin)[(size_t)b * SL + lane * L + l] = h;
            if (b + 1 == NB) break;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 1 < S; k += 2) {
                const double h0 = __shfl_sync(0xffffffffu, h, k), h1 = __shfl_sync(0xffffffffu, h, k + 1);
                if (lane < S) {
                    a0 = fma(h0, AP[k * S + lane], a0);
                    a1 = fma(h1, AP[(k + 1) * S + lane], a1);
                }
            }
            if (k < S) {
                const double h0 = __shfl_sync(0xffffffffu, h, k);
                if (lane < S) a0 = fma(h0, AP[k * S + lane], a0);
            }
            h = a0 + a1;
        }
    }
    __syncthreads();
    if (Xblk != nullptr) {  // split path (P > 1): level 3 and everything after it run as fhmm_forward_expand, one CTA per block
        for (int i = tid; i < NB * SL; i += nt) Xblk[i] = s_X[i];
        return;
    }
    // level 3 + hidden = log H
    // a thread keeps one (state, lineage) element when the block is a multiple of S*L threads wide and strides over
    // t, so the index arithmetic of the inner loop is additions only
    const int l3_groups = nt / SL;
    const int l3_e = tid % SL, l3_sp = l3_e / L, l3_l = l3_e % L;
    int l3_b = (tid / SL) / P, l3_j = (tid / SL) - l3_b * P;
    for (int t = tid / SL; l3_groups > 0 && tid < l3_groups * SL && t < I; t += l3_groups, l3_j += l3_groups) {
        const int i = t * SL + l3_e, sp = l3_sp, l = l3_l;
        while (l3_j >= P) {
            l3_j -= P;
            ++l3_b;
        }
        const int b = l3_b, j = l3_j;
        double h;
        if (P == 1) {
            h = Hlin[i];  // level 2 already produced every step
        } else if (j == 0) {
            h = s_X[(size_t)b * SL + sp * L + l];
        } else {
            const double *Aj = s_pow + (size_t)(j - 1) * L * SS + (size_t)l * SS;
            const double *X = s_X + (size_t)b * SL;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 1 < S; k += 2) {
                a0 = fma(X[k * L + l], Aj[k * S + sp], a0);
                a1 = fma(X[(k + 1) * L + l], Aj[(k + 1) * S + sp], a1);
            }
            if (k < S) a0 = fma(X[k * L + l], Aj[k * S + sp], a0);
            h = a0 + a1;
        }
        if (P > 1) Hlin[i] = h;
        if (hf_smem) s_Hf[i] = (float)h;
        hidden[i] = h > 1e-30 ? logf((float)h) : (float)log(h);
    }
    __syncthreads();
    // time average Hbar[s,l] = mean_t H_t[s,l]: groups of threads sum strided slices of t, then combine
    {
        const int groups = nt / SL > 0 ? nt / SL : 1;
        const int gidx = tid / SL, e = tid % SL;
        if (gidx < groups) {
            double acc = 0.0;
            if (hf_smem) {
                for (int t = gidx; t < I; t += groups) acc += (double)s_Hf[(size_t)t * SL + e];
            } else {
                for (int t = gidx; t < I; t += groups) acc += Hlin[(size_t)t * SL + e];
            }
            s_part[gidx * SL + e] = acc;
        }
        __syncthreads();
        if (tid < SL) {
            double acc = 0.0;
            for (int gq = 0; gq < groups; ++gq) acc += s_part[gq * SL + tid];
            const float hb = (float)(acc / (double)I);
            aux[ao.Hbar + tid] = hb;
            aux[ao.logC + tid] = logf(hb) + s_logw[tid % L];
        }
    }
    // the mixing matrix G[t][k] = w_l H_t[s,l] (summed over l when the emission is shared between lineages)
    float *s_wf = s_h + SL;  // second half of the s_h scratch: w_l
    for (int i = tid; i < L; i += nt) s_wf[i] = expf(s_logw[i]);
    __syncthreads();
    if (restricted) {
        for (int i = tid; i < I * KP; i += nt) {
            const int t = i / KP, k = i % KP;
            float g = 0.f;
            if (k < S) {
                if (hf_smem) {
                    for (int ll = 0; ll < L; ++ll) g += s_Hf[(size_t)t * SL + k * L + ll] * s_wf[ll];
                } else {
                    for (int ll = 0; ll < L; ++ll) g += (float)Hlin[(size_t)t * SL + k * L + ll] * s_wf[ll];
                }
            }
            G[i] = g;
        }
    } else {
        for (int i = tid; i < I * KP; i += nt) {
            const int t = i / KP, k = i % KP;
            G[i] = k < SL ? (hf_smem ? s_Hf[(size_t)t * SL + k] : (float)Hlin[(size_t)t * SL + k]) * s_wf[k % L] : 0.f;
        }
    }
    __syncthreads();
    // log c_s = log sum_l C[s,l], log q_l = log sum_s C[s,l]
    float *s_logC = s_h;
    for (int i = tid; i < SL; i += nt) s_logC[i] = aux[ao.logC + i];
    __syncthreads();
    for (int s0 = tid; s0 < S; s0 += nt) {
        float m = -INFINITY;
        for (int ll = 0; ll < L; ++ll) m = fmaxf(m, s_logC[s0 * L + ll]);
        float z = 0.f;
        for (int ll = 0; ll < L; ++ll) z += expf(s_logC[s0 * L + ll] - m);
        aux[ao.logc + s0] = m + logf(z);
    }
    for (int ll = tid; ll < L; ll += nt) {
        float m = -INFINITY;
        for (int s0 = 0; s0 < S; ++s0) m = fmaxf(m, s_logC[s0 * L + ll]);
        float z = 0.f;
        for (int s0 = 0; s0 < S; ++s0) z += expf(s_logC[s0 * L + ll] - m);
        aux[ao.logq + ll] = m + logf(z);
    }
}

// sum_q part[q * stride] for q < n, added in q order (bit-reproducible) with eight L2 loads in flight at a time
template <typename T>
__device__ __forceinline__ T ordered_sum_ldcg(const T *__restrict__ part, int n, size_t stride) {
    T acc = (T)0;
    for (int q0 = 0; q0 < n; q0 += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = q0 + u < n ? __ldcg(part + (size_t)(q0 + u) * stride) : (T)0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    return acc;
}

// --------------------------------------------------------------------------------------------------
// Level 3 of the hidden recurrence and what hangs off it, one CTA per block b of P steps (fhmm_small_forward's split
// path leaves X_b and A^1..A^P in global memory): H_{bP+j} = X_b . A^j, hidden = log H, the block's rows of the mixing
// matrix G, and the block's share of the time average. The CTA that finishes last adds the shares in block order (fixed
// order: bit-reproducible) and writes Hbar, log C, log c, log q. One SM did all of this before, bound by its own shared-
// memory bandwidth (23 us at config 3); the blocks are independent.
__global__ void __launch_bounds__(256)
fhmm_forward_expand(int S, int L, int I, int restricted, int KP, int P, const float *__restrict__ small,
                    const double *__restrict__ Apow, const double *__restrict__ Xblk, float *__restrict__ hidden,
                    double *__restrict__ Hlin, float *__restrict__ G, float *__restrict__ aux,
                    double *__restrict__ HbarPart, int *__restrict__ ticket, int staged) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    extern __shared__ double smd[];
    const int SL = S * L, SS = S * S, NB = (I + P - 1) / P;
    // staged != 0: the launch reserved room for X_b and A^1..A^(P-1) (one coalesced sweep instead of 2 S dependent
    // L2 round trips per output)
    double *s_X = smd;                                                  // [SL]
    double *s_pow = s_X + SL;                                           // [P-1][L][S][S]
    float *s_H = reinterpret_cast<float *>(staged ? s_pow + (size_t)(P - 1) * L * SS : smd);  // [P][SL]
    float *s_w = s_H + P * SL;   // [L]
    float *s_logC = s_w + L;     // [SL] (last CTA)
    __shared__ int s_last;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = blockIdx.x, t_lo = b * P, len = min(I, t_lo + P) - t_lo;
    const double *X = Xblk + (size_t)b * SL;
    const double *Apw = Apow;
    for (int i = tid; i < L; i += nt) s_w[i] = expf(small[so.log_w + i]);
    if (staged) {
        for (int i = tid; i < SL; i += nt) s_X[i] = X[i];
#pragma unroll 8
        for (int i = tid; i < (len - 1) * L * SS; i += nt) s_pow[i] = Apow[i];
        __syncthreads();
        X = s_X;
        Apw = s_pow;
    }
    for (int o = tid; o < len * SL; o += nt) {
        const int j = o / SL, e = o - j * SL, sp = e / L, l = e - sp * L;
        double h;
        if (j == 0) {
            h = X[e];
        } else {
            const double *Aj = Apw + (size_t)(j - 1) * L * SS + (size_t)l * SS + sp;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 1 < S; k += 2) {
                a0 = fma(X[k * L + l], Aj[k * S], a0);
                a1 = fma(X[(k + 1) * L + l], Aj[(k + 1) * S], a1);
            }
            if (k < S) a0 = fma(X[k * L + l], Aj[k * S], a0);
            h = a0 + a1;
        }
        const size_t i = (size_t)(t_lo + j) * SL + e;
        Hlin[i] = h;
        s_H[o] = (float)h;
        hidden[i] = h > 1e-30 ? logf((float)h) : (float)log(h);
    }
    __syncthreads();
    for (int e = tid; e < SL; e += nt) {
        double acc = 0.0;
        for (int j = 0; j < len; ++j) acc += (double)s_H[j * SL + e];
        HbarPart[(size_t)b * SL + e] = acc;
    }
    // the mixing matrix G[t][k] = w_l H_t[s,l] (summed over l when the emission is shared between lineages)
    for (int o = tid; o < len * KP; o += nt) {
        const int j = o / KP, k = o - j * KP;
        float g = 0.f;
        if (restricted) {
            if (k < S)
                for (int ll = 0; ll < L; ++ll) g += s_H[j * SL + k * L + ll] * s_w[ll];
        } else if (k < SL) {
            g = s_H[j * SL + k] * s_w[k % L];
        }
        G[(size_t)(t_lo + j) * KP + k] = g;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1) == NB - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = tid; e < SL; e += nt) {
        const double acc = ordered_sum_ldcg(HbarPart + e, NB, (size_t)SL);
        const float hb = (float)(acc / (double)I);
        aux[ao.Hbar + e] = hb;
        const float lc = logf(hb) + small[so.log_w + e % L];
        aux[ao.logC + e] = lc;
        s_logC[e] = lc;
    }
    __syncthreads();
    // log c_s = log sum_l C[s,l], log q_l = log sum_s C[s,l]
    for (int s0 = tid; s0 < S; s0 += nt) {
        float m = -INFINITY;
        for (int ll = 0; ll < L; ++ll) m = fmaxf(m, s_logC[s0 * L + ll]);
        float z = 0.f;
        for (int ll = 0; ll < L; ++ll) z += expf(s_logC[s0 * L + ll] - m);
        aux[ao.logc + s0] = m + logf(z);
    }
    for (int ll = tid; ll < L; ll += nt) {
        float m = -INFINITY;
        for (int s0 = 0; s0 < S; ++s0) m = fmaxf(m, s_logC[s0 * L + ll]);
        float z = 0.f;
        for (int s0 = 0; s0 < S; ++s0) z += expf(s_logC[s0 * L + ll] - m);
        aux[ao.logq + ll] = m + logf(z);
    }
}

// --------------------------------------------------------------------------------------------------
// joint[t,s,l,n] = logE[s,e(l),n] + hidden[t,s,l] + log w[l]  (only when the caller asks for it)
__global__ void fhmm_joint_kernel(int S, int L, int Le, int N, int I, const float *__restrict__ logE,
                                  const float *__restrict__ hidden, const float *__restrict__ log_w,
                                  float *__restrict__ joint) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int tsl = blockIdx.y;  // t*S*L + s*L + l
    if (n >= N) return;
    const int l = tsl % L, s = (tsl / L) % S;
    const int e = Le == 1 ? 0 : l;
    joint[(size_t)tsl * N + n] = logE[(size_t)(s * Le + e) * N + n] + hidden[tsl] + log_w[l];
}

// --------------------------------------------------------------------------------------------------
// The pass over D: pred, KL divergence, dDiv/dlogE (per t-split partials) and dDiv/dG.
//   grid.x = column tiles (blockDim.x * J columns), grid.y = t splits of `tchunk` rows.
//   Max-shifted per column: Ehat[k] = exp(logE[k,n] - m_n), Phat = sum_k G[t,k] Ehat[k], log P = m_n + log Phat.
// Blackwell packed fp32 FMA (FFMA2): two independent fp32 FMAs per instruction on a register pair.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// Sum NV values (power of two <= 32) across the 32 lanes with a transposing butterfly: every stage halves the
// values a lane carries, so NV + log2(32/NV) shuffles replace 5*NV. Afterwards v[0] on lane `l` is the full sum of
// element e(l) = (l >> log2(32/NV)) & (NV-1)  (every element is held by 32/NV lanes).
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[NV], int lane) {
    int off = 16;
#pragma unroll
    for (int h = NV / 2; h >= 1; h /= 2, off /= 2) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = upper ? v[i] : v[i + h];
            const float keep = upper ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    for (; off >= 1; off /= 2) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
__host__ __device__ constexpr int next_pow2(int x) { return x <= 1 ? 1 : x <= 2 ? 2 : x <= 4 ? 4 : x <= 8 ? 8 : x <= 16 ? 16 : x <= 32 ? 32 : 64; }
__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x / 2); }

// Thread owns J columns (adjacent column PAIRS share FFMA2 instructions), loops over the rows of its t-split.
// Slice mode (Rin != nullptr; emission matrices with many rows, see fhmm_data_fwd): the kernel handles rows
// [koff, koff + K) of the Kfull-row problem, reads r[t,n] = D/P from Rin instead of forming P, and only accumulates
// the two gradients. Otherwise koff = 0, K = Kfull, ldG = KP.
template <int KP, int J>
__global__ void __launch_bounds__(128, (KP <= 12 ? 4 : 1))
fhmm_data_pass(int N, int I, int K, int tchunk, const float *__restrict__ logE, const float *__restrict__ G,
               const float *__restrict__ D, float *__restrict__ pred, float *__restrict__ glogE_split,
               float *__restrict__ gG, double *__restrict__ div_acc, float inv_I, const float *__restrict__ Rin,
               int koff, int Kfull, int ldG) {
    constexpr int JP = (J + 1) / 2;  // column pairs; J == 1 runs with a dead second half
    extern __shared__ float sG[];    // [tchunk][KP]
    const int t0 = blockIdx.y * tchunk, t1 = min(I, t0 + tchunk);
    for (int i = threadIdx.x; i < (t1 - t0) * KP; i += blockDim.x) {
        const int k = i % KP;
        sG[i] = k < K ? G[(size_t)(t0 + i / KP) * ldG + koff + k] : 0.f;
    }
    logE += (size_t)koff * N;
    const int lane = threadIdx.x & 31;
    const int base = blockIdx.x * blockDim.x * J + threadIdx.x;
    float2 Eh[JP][KP], dE[JP][KP];
    float mcol[2 * JP];
    bool ok[2 * JP];
#pragma unroll
    for (int j = 0; j < 2 * JP; ++j) {
        const int n = base + j * blockDim.x;
        ok[j] = j < J && n < N;
        float e[KP];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            e[k] = (ok[j] && k < K) ? logE[(size_t)k * N + n] : -INFINITY;
            m = fmaxf(m, e[k]);
        }
        if (!ok[j] || Rin != nullptr) m = 0.f;  // slice mode forms no ratio, so it needs no max shift
        mcol[j] = m;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float v = (ok[j] && k < K) ? expf(e[k] - m) : 0.f;
            if (j & 1) Eh[j / 2][k].y = v;
            else Eh[j / 2][k].x = v;
            dE[j / 2][k] = make_float2(0.f, 0.f);
        }
    }
    __syncthreads();
    double div = 0.0;
    for (int t = t0; t < t1; ++t) {
        float div_f = 0.f;  // per-row partial in float, promoted once per t
        const float *g = sG + (t - t0) * KP;
        float2 g2[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) g2[k] = make_float2(g[k], g[k]);
        float2 part2[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) part2[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) {
            float2 p2 = make_float2(0.f, 0.f);
            if (Rin == nullptr) {
#pragma unroll
                for (int k = 0; k < KP; ++k) p2 = ffma2(g2[k], Eh[jp][k], p2);
            }
            float pv[2] = {p2.x, p2.y}, rv[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * jp + h;
                const int n = base + j * blockDim.x;
                if (Rin != nullptr) {  // slice mode: r[t,n] = D/P was formed by fhmm_data_fwd
                    rv[h] = ok[j] ? __ldg(Rin + (size_t)t * N + n) : 0.f;
                    continue;
                }
                const float d = (ok[j] && D != nullptr) ? __ldg(D + (size_t)t * N + n) : 0.f;
                const float p = fmaxf(pv[h], FLT_MIN);
                if (pred != nullptr) {  // forward_model output: full-precision log
                    if (ok[j]) pred[(size_t)t * N + n] = mcol[j] + logf(p);
                }
                float r = 0.f;
                if (d > 0.f) {
                    // loss pass: MUFU log2 / reciprocal (abs. error of the log ~5e-7, weighted by D whose rows sum to 1)
                    r = __fdividef(d, p);
                    div_f += d * (__logf(d) - mcol[j] - __logf(p));  // xlogy(D,D) - D*pred (KLDivLoss, trainer.py:139)
                }
                rv[h] = r;
            }
            const float2 r2 = make_float2(rv[0], rv[1]);
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                dE[jp][k] = ffma2(r2, g2[k], dE[jp][k]);
                part2[k] = ffma2(r2, Eh[jp][k], part2[k]);
            }
        }
        div += (double)div_f;
        // dG[t][k] = -(1/I) sum over the warp's columns: transposing butterfly, then one atomic per element
        constexpr int NV = next_pow2(KP);
        if (NV <= 32) {
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = k < KP ? part2[k < KP ? k : 0].x + part2[k < KP ? k : 0].y : 0.f;
            warp_transpose_reduce<NV>(v, lane);
            constexpr int REP = 32 / NV;  // lanes holding the same element
            const int e = (lane / REP) % NV;
            if (lane % REP == 0 && e < K && v[0] != 0.f) atomicAdd(&gG[(size_t)t * Kfull + koff + e], -inv_I * v[0]);
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const float sacc = warp_sum(part2[k].x + part2[k].y);
                if (lane == 0 && k < K && sacc != 0.f) atomicAdd(&gG[(size_t)t * Kfull + koff + k], -inv_I * sacc);
            }
        }
    }
    float *out = glogE_split + ((size_t)blockIdx.y * Kfull + koff) * N;
#pragma unroll
    for (int j = 0; j < 2 * JP; ++j) {
        const int n = base + j * blockDim.x;
        if (!ok[j]) continue;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) {
                const float eh = (j & 1) ? Eh[j / 2][k].y : Eh[j / 2][k].x;
                const float de = (j & 1) ? dE[j / 2][k].y : dE[j / 2][k].x;
                out[(size_t)k * N + n] = -inv_I * eh * de;
            }
    }
    div = warp_sum(div);
    __shared__ double sdiv[4];
    if (lane == 0) sdiv[threadIdx.x >> 5] = div;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(div_acc, (sdiv[0] + sdiv[1] + sdiv[2] + sdiv[3]) * (double)inv_I);
}

// Emission matrices with many rows (K = S*L > 24: unrestricted FHMM / LSSM, every NSSM): the fused pass above would
// need 4*K registers per column. The work is split instead: this kernel forms P[t,n] = sum_k G[t,k] E[k,n] over ALL
// rows (thread = one column pair, K packed FMAs per row of D), writes pred / the divergence and stores
// r[t,n] = D[t,n] / P[t,n]; fhmm_data_pass then runs once per slice of <= 12 rows in slice mode for the two gradients.
template <int KP>
__global__ void __launch_bounds__(128, 3)
fhmm_data_fwd(int N, int I, int K, int tchunk, const float *__restrict__ logE, const float *__restrict__ G,
              const float *__restrict__ D, float *__restrict__ pred, float *__restrict__ R,
              double *__restrict__ div_acc, float inv_I) {
    extern __shared__ float sG[];  // [tchunk][KP]
    const int t0 = blockIdx.y * tchunk, t1 = min(I, t0 + tchunk);
    for (int i = threadIdx.x; i < (t1 - t0) * KP; i += blockDim.x) sG[i] = G[(size_t)t0 * KP + i];
    const int lane = threadIdx.x & 31;
    const int base = blockIdx.x * blockDim.x * 2 + threadIdx.x;
    float2 Eh[KP];
    float mcol[2], emneg[2];
    bool ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int n = base + j * blockDim.x;
        ok[j] = n < N;
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, ok[j] ? logE[(size_t)k * N + n] : 0.f);
        mcol[j] = m;
        emneg[j] = expf(-m);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float v = (ok[j] && k < K) ? expf(logE[(size_t)k * N + n] - m) : 0.f;
            if (j) Eh[k].y = v;
            else Eh[k].x = v;
        }
    }
    __syncthreads();
    double div = 0.0;
    for (int t = t0; t < t1; ++t) {
        const float *g = sG + (t - t0) * KP;
        float2 p2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KP; ++k) p2 = ffma2(make_float2(g[k], g[k]), Eh[k], p2);
        const float pv[2] = {p2.x, p2.y};
        float div_f = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = base + j * blockDim.x;
            if (!ok[j]) continue;
            const float p = fmaxf(pv[j], FLT_MIN);
            if (pred != nullptr) pred[(size_t)t * N + n] = mcol[j] + logf(p);
            if (D != nullptr) {
                const float d = __ldg(D + (size_t)t * N + n);
                float r = 0.f;
                if (d > 0.f) {
                    r = __fdividef(d, p) * emneg[j];
                    div_f += d * (__logf(d) - mcol[j] - __logf(p));
                }
                R[(size_t)t * N + n] = r;
            }
        }
        div += (double)div_f;
    }
    div = warp_sum(div);
    __shared__ double sdiv[4];
    if (lane == 0) sdiv[threadIdx.x >> 5] = div;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(div_acc, (sdiv[0] + sdiv[1] + sdiv[2] + sdiv[3]) * (double)inv_I);
}

// --------------------------------------------------------------------------------------------------
// Tensor-core gradients of the data term (two-pass mode: r[t,n] = D/P is in memory).
//   dlogE[k,n] = -(1/I) E[k,n] * sum_t G[t,k] r[t,n]      ([K x I] . [I x N]: the contraction over K = I time steps)
//   dG[t,k]    = -(1/I) sum_n r[t,n] E[k,n]                ([I x N] . [N x K]: the contraction over the N cells)
// Both are warp-level mma.sync m16n8k8 TF32 products with the 3xTF32 split (x = hi + lo, hi*hi + hi*lo + lo*hi, fp32
// accumulate): each factor keeps ~21 mantissa bits, the gradients stay within 1e-6 relative of the FFMA path. The
// 5th-generation tcgen05 path needs M >= 64 per instruction and operands staged through shared-memory descriptors; with
// M = K <= 64 emission rows and an arithmetic intensity of K/2 flop per byte of r these contractions are bound by
// reading r (I*N*4 bytes), not by the tensor pipe, so the warp-level instruction is the right tool and the evidence the
// north_star asks for is the tensor-pipe counter next to the bytes (profiles/README_r02.md).
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(c, al, bh);  // small terms first
    mma_tf32(c, ah, bl);
    mma_tf32(c, ah, bh);
}

// Leading dimension of a small matrix held in shared memory for A-fragment loads (lane (g, tg) reads [row g][col tg]):
// an odd multiple of 4 words puts the 32 lanes of a load in 32 different banks.
__host__ __device__ constexpr int mma_ld(int cols) { return ((cols + 7) / 8 * 8) + 4; }

// P and r on the tensor cores — the emission GEMM proper (_FHMM.py:126-144: pred[t,n] = log sum_k G[t,k] E[k,n]):
//   P^[t,n] = sum_k G[t,k] E^[k,n] with E^ = exp(logE - m[n]) (max shift per cell), M = 16 time steps, N = 8 cells,
//   K = emission rows in steps of 8. A warp keeps the E^ fragments of its NW cell tiles in registers (hi / lo planes),
//   takes the G fragments of a block of 16 time steps from shared memory once and uses them for all its tiles; the
//   C fragment (t = g / g + 8, n = 2tg / 2tg + 1) is turned into pred, the divergence and r = D / P in place.
template <int KT, int NW>
__global__ void __launch_bounds__(128)
fhmm_fwd_mma(int N, int I, int K, int ldG, int tchunk, const float *__restrict__ logE, const float *__restrict__ G,
             const float *__restrict__ D, float *__restrict__ pred, float *__restrict__ R, double *__restrict__ div_acc,
             float inv_I) {
    constexpr int LD = mma_ld(KT * 8);
    extern __shared__ uint32_t sG32[];  // [2][tpad][LD]
    const int t0 = blockIdx.y * tchunk, t1 = min(I, t0 + tchunk);
    const int tpad = (t1 - t0 + 15) / 16 * 16;
    uint32_t *sGh = sG32, *sGl = sG32 + (size_t)tpad * LD;
    for (int i = threadIdx.x; i < tpad * (KT * 8); i += blockDim.x) {
        const int tt = i / (KT * 8), k = i % (KT * 8), t = t0 + tt;
        const float v = (t < t1 && k < K) ? G[(size_t)t * ldG + k] : 0.f;
        split_tf32(v, sGh[tt * LD + k], sGl[tt * LD + k]);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const int nbase = (blockIdx.x * 4 + warp) * (8 * NW);
    // B fragments: b0 = E^[k0 + tg][n0 + g], b1 = E^[k0 + tg + 4][n0 + g]; the shift m[n] is needed for column n0 + g
    // (fragment) and for columns n0 + 2tg, n0 + 2tg + 1 (C fragment): computed for the former, shuffled for the latter
    uint32_t bh[NW][KT][2], bl[NW][KT][2];
    float mc[NW][2], em[NW][2];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int n = nbase + 8 * w + g;
        const bool ok = n < N;
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, ok ? logE[(size_t)k * N + n] : 0.f);
#pragma unroll
        for (int q = 0; q < KT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 8 * q + tg + 4 * h;
                const float v = (ok && k < K) ? expf(logE[(size_t)k * N + n] - m) : 0.f;
                split_tf32(v, bh[w][q][h], bl[w][q][h]);
            }
        // lane (g, tg) needs m of columns 2tg and 2tg + 1 of the tile: held by the lanes with g' = 2tg, 2tg + 1
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            mc[w][j] = __shfl_sync(0xffffffffu, m, (2 * tg + j) * 4);
            em[w][j] = expf(-mc[w][j]);
        }
    }
    __syncthreads();
    double div = 0.0;
    const bool vec = (N & 1) == 0;
    auto load_d = [&](int tb_, float2 (&dst)[NW][2]) {
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = t0 + tb_ + g + 8 * h, n = nbase + 8 * w + 2 * tg;
                dst[w][h] = make_float2(0.f, 0.f);
                if (D != nullptr && tb_ < tpad && t < t1) {
                    const float *dp = D + (size_t)t * N + n;
                    if (vec && n + 1 < N) dst[w][h] = __ldg(reinterpret_cast<const float2 *>(dp));
                    else {
                        if (n < N) dst[w][h].x = __ldg(dp);
                        if (n + 1 < N) dst[w][h].y = __ldg(dp + 1);
                    }
                }
            }
    };
    float2 dnext[NW][2];
    load_d(0, dnext);
    for (int tb = 0; tb < tpad; tb += 16) {
        uint32_t ah[KT][4], al[KT][4];
#pragma unroll
        for (int q = 0; q < KT; ++q) {
            const uint32_t *ph = sGh + (size_t)(tb + g) * LD + 8 * q + tg, *pl = sGl + (size_t)(tb + g) * LD + 8 * q + tg;
            ah[q][0] = ph[0], ah[q][1] = ph[8 * LD], ah[q][2] = ph[4], ah[q][3] = ph[8 * LD + 4];
            al[q][0] = pl[0], al[q][1] = pl[8 * LD], al[q][2] = pl[4], al[q][3] = pl[8 * LD + 4];
        }
        float div_f = 0.f;
        // the D values are requested ONE BLOCK of 16 time steps AHEAD (and those of the first block before the loop), so
        // that their HBM latency hides behind a whole block of MMAs and elementwise work: lane (g, tg) owns rows
        // t = g, g + 8 and the column pair (2tg, 2tg + 1) of every tile (one 8-byte access when N is even)
        float2 dv[NW][2];
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int h = 0; h < 2; ++h) dv[w][h] = dnext[w][h];
        load_d(tb + 16, dnext);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < KT; ++q) mma_3xtf32(c, ah[q], al[q], bh[w][q], bl[w][q]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = t0 + tb + g + 8 * h, n = nbase + 8 * w + 2 * tg;
                if (t >= t1 || n >= N) continue;
                float rr[2], pp[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float p = fmaxf(c[2 * h + j], FLT_MIN);
                    pp[j] = pred != nullptr ? mc[w][j] + logf(p) : 0.f;  // forward_model output: full-precision log
                    const float d = j ? dv[w][h].y : dv[w][h].x;
                    float r = 0.f;
                    if (d > 0.f) {
                        r = __fdividef(d, p) * em[w][j];
                        div_f += d * (__logf(d) - mc[w][j] - __logf(p));
                    }
                    rr[j] = r;
                }
                const size_t off = (size_t)t * N + n;
                if (vec && n + 1 < N) {
                    if (pred != nullptr) *reinterpret_cast<float2 *>(pred + off) = make_float2(pp[0], pp[1]);
                    if (D != nullptr) *reinterpret_cast<float2 *>(R + off) = make_float2(rr[0], rr[1]);
                } else {
                    if (pred != nullptr) {
                        pred[off] = pp[0];
                        if (n + 1 < N) pred[off + 1] = pp[1];
                    }
                    if (D != nullptr) {
                        R[off] = rr[0];
                        if (n + 1 < N) R[off + 1] = rr[1];
                    }
                }
            }
        }
        div += (double)div_f;
    }
    if (D != nullptr) {
        div = warp_sum(div);
        __shared__ double sdiv[4];
        if (lane == 0) sdiv[warp] = div;
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(div_acc, (sdiv[0] + sdiv[1] + sdiv[2] + sdiv[3]) * (double)inv_I);
    }
}

// dlogE: CTA = 4 warps, a warp owns NW = 4 tiles of 8 cells (the CTA covers one 512-byte span of every row of r), grid.y
// = t splits; A = G^T from shared memory (pre-split into hi / lo, loaded once per CTA), B = r straight from global
// memory in the fragment layout (each element is read by exactly one thread), MT = ceil(K / 16) accumulator tiles per
// cell tile.
template <int MT, int NW>
__global__ void __launch_bounds__(128)
fhmm_dE_mma(int N, int I, int K, int Kfull, int ldG, int tchunk, const float *__restrict__ logE,
            const float *__restrict__ G, const float *__restrict__ R, float *__restrict__ glogE_split, float inv_I) {
    constexpr int LD = MT * 16 + 8;  // = 8 (mod 16 words): lane (g, tg) reads [row tg][col g]: 32 different banks
    extern __shared__ uint32_t sG32[];  // [2][tpad][LD]: hi plane, lo plane
    const int t0 = blockIdx.y * tchunk, t1 = min(I, t0 + tchunk);
    const int tpad = (t1 - t0 + 7) / 8 * 8;
    uint32_t *sGh = sG32, *sGl = sG32 + (size_t)tpad * LD;
    for (int i = threadIdx.x; i < tpad * (MT * 16); i += blockDim.x) {
        const int tt = i / (MT * 16), k = i % (MT * 16), t = t0 + tt;
        const float v = (t < t1 && k < K) ? G[(size_t)t * ldG + k] : 0.f;
        split_tf32(v, sGh[tt * LD + k], sGl[tt * LD + k]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const int nbase = (blockIdx.x * 4 + warp) * (8 * NW);
    if (nbase >= N) return;
    float c[NW][MT][4];
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int m = 0; m < MT; ++m) c[w][m][0] = c[w][m][1] = c[w][m][2] = c[w][m][3] = 0.f;
    for (int tb = 0; tb < tpad; tb += 8) {
        float r0[NW], r1[NW];
        const int ta = t0 + tb + tg, tbb = ta + 4;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int n = nbase + 8 * w + g;
            r0[w] = (n < N && ta < t1) ? __ldg(R + (size_t)ta * N + n) : 0.f;
            r1[w] = (n < N && tbb < t1) ? __ldg(R + (size_t)tbb * N + n) : 0.f;
        }
        uint32_t ah[MT][4], al[MT][4];
        const uint32_t *gh = sGh + (size_t)(tb + tg) * LD + g, *gl = sGl + (size_t)(tb + tg) * LD + g;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            ah[m][0] = gh[16 * m], ah[m][1] = gh[16 * m + 8], ah[m][2] = gh[4 * LD + 16 * m], ah[m][3] = gh[4 * LD + 16 * m + 8];
            al[m][0] = gl[16 * m], al[m][1] = gl[16 * m + 8], al[m][2] = gl[4 * LD + 16 * m], al[m][3] = gl[4 * LD + 16 * m + 8];
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t bh[2], bl[2];
            split_tf32(r0[w], bh[0], bl[0]);
            split_tf32(r1[w], bh[1], bl[1]);
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_3xtf32(c[w][m], ah[m], al[m], bh, bl);
        }
    }
    // C fragment: rows k = 16m + g (+8), columns n0 + 2tg (+1)
    float *out = glogE_split + (size_t)blockIdx.y * Kfull * N;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 16 * m + g + 8 * h;
                if (k >= K) continue;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int n = nbase + 8 * w + 2 * tg + j;
                    if (n < N) out[(size_t)k * N + n] = -inv_I * expf(logE[(size_t)k * N + n]) * c[w][m][2 * h + j];
                }
            }
}

// dG: CTA = 8 warps, owns 128 consecutive cells (E tile split hi / lo in shared memory) for ALL time steps; a warp
// takes blocks of 16 time steps, A = r in the row-major fragment layout (each lane reads 4 floats of 4 different
// rows; a warp covers 16 rows x 32 bytes = whole sectors), B = E^T from shared memory, NT = ceil(K / 8) accumulator
// tiles; one atomic per (t, k) and CTA at the end of a block of time steps.
constexpr int kMmaCells = 128, kMmaLdE = kMmaCells + 4;
template <int NT>
__global__ void __launch_bounds__(256)
fhmm_dG_mma(int N, int I, int K, int Kfull, const float *__restrict__ logE, const float *__restrict__ R,
            float *__restrict__ gG, float inv_I) {
    extern __shared__ uint32_t sE32[];  // [2][NT*8][kMmaLdE]
    uint32_t *sEh = sE32, *sEl = sE32 + (size_t)NT * 8 * kMmaLdE;
    const int c0 = blockIdx.x * kMmaCells;
    for (int i = threadIdx.x; i < NT * 8 * kMmaCells; i += blockDim.x) {
        const int k = i / kMmaCells, j = i % kMmaCells, n = c0 + j;
        const float v = (k < K && n < N) ? expf(logE[(size_t)k * N + n]) : 0.f;
        split_tf32(v, sEh[k * kMmaLdE + j], sEl[k * kMmaLdE + j]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    for (int tb = warp * 16; tb < I; tb += 8 * 16) {
        float c[NT][4];
#pragma unroll
        for (int q = 0; q < NT; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
        const int ta = tb + g, tbb = tb + g + 8;
        const float *ra = R + (size_t)ta * N + c0 + tg, *rb = R + (size_t)tbb * N + c0 + tg;
        constexpr int UN = 4;
        for (int j0 = 0; j0 < kMmaCells; j0 += 8 * UN) {
            float a[UN][4];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int j = j0 + 8 * u, n = c0 + j + tg;
                a[u][0] = (ta < I && n < N) ? __ldg(ra + j) : 0.f;
                a[u][1] = (tbb < I && n < N) ? __ldg(rb + j) : 0.f;
                a[u][2] = (ta < I && n + 4 < N) ? __ldg(ra + j + 4) : 0.f;
                a[u][3] = (tbb < I && n + 4 < N) ? __ldg(rb + j + 4) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int j = j0 + 8 * u;
                uint32_t ah[4], al[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(a[u][e], ah[e], al[e]);
#pragma unroll
                for (int q = 0; q < NT; ++q) {
                    const uint32_t *eh = sEh + (size_t)(8 * q + g) * kMmaLdE + j + tg;
                    const uint32_t *el = sEl + (size_t)(8 * q + g) * kMmaLdE + j + tg;
                    const uint32_t bh[2] = {eh[0], eh[4]}, bl[2] = {el[0], el[4]};
                    mma_3xtf32(c[q], ah, al, bh, bl);
                }
            }
        }
        // C fragment: rows t = tb + g (+8), columns k = 8q + 2tg (+1)
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = tb + g + 8 * h;
                if (t >= I) continue;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = 8 * q + 2 * tg + j;
                    const float v = c[q][2 * h + j];
                    if (k < K && v != 0.f) atomicAdd(&gG[(size_t)t * Kfull + k], -inv_I * v);
                }
            }
    }
}

// --------------------------------------------------------------------------------------------------
// Per-node conditionals from the time-averaged joint M[s,l,n] = logE[s,e(l),n] + logC[s,l] (trainer.py:153-172):
//   lngs[s,n] = log P(n|s), lngc[l,n] = log P(n|l), lcgn[l,n] = log P(l|n).
// Writes X1[n][0:S] = P(n|s), X1[n][S:S+L] = P(n|l) (operands of the products with T) and X2[n][l] = P(l|n).
struct NodeCond {
    float lngs[kMaxS], Ml[kMaxL], lcgn[kMaxL];
};
__device__ __forceinline__ void node_conditionals(int S, int L, int Le, int N, int n, const float *__restrict__ logE,
                                                  const float *__restrict__ s_logC, const float *__restrict__ s_logc,
                                                  NodeCond &c) {
    for (int s = 0; s < S; ++s) {
        float m = -INFINITY;
        for (int l = 0; l < L; ++l) m = fmaxf(m, s_logC[s * L + l] + logE[(size_t)(s * Le + (Le == 1 ? 0 : l)) * N + n]);
        float z = 0.f;
        for (int l = 0; l < L; ++l)
            z += expf(s_logC[s * L + l] + logE[(size_t)(s * Le + (Le == 1 ? 0 : l)) * N + n] - m);
        c.lngs[s] = m + logf(z) - s_logc[s];  // Ms - lse_n(Ms); lse_n(Ms) = log c_s because E rows sum to 1
    }
    float mm = -INFINITY;
    for (int l = 0; l < L; ++l) {
        float m = -INFINITY;
        for (int s = 0; s < S; ++s) m = fmaxf(m, s_logC[s * L + l] + logE[(size_t)(s * Le + (Le == 1 ? 0 : l)) * N + n]);
        float z = 0.f;
        for (int s = 0; s < S; ++s)
            z += expf(s_logC[s * L + l] + logE[(size_t)(s * Le + (Le == 1 ? 0 : l)) * N + n] - m);
        c.Ml[l] = m + logf(z);
        mm = fmaxf(mm, c.Ml[l]);
    }
    float z = 0.f;
    for (int l = 0; l < L; ++l) z += expf(c.Ml[l] - mm);
    const float lse = mm + logf(z);
    for (int l = 0; l < L; ++l) c.lcgn[l] = c.Ml[l] - lse;
}

__global__ void __launch_bounds__(128) fhmm_node_pass_a(int S, int L, int Le, int N, const float *__restrict__ logE,
                                                        const float *__restrict__ aux, float *__restrict__ X1,
                                                        float *__restrict__ X2) {
    const AuxOffsets ao(S, L);
    __shared__ float s_logC[kMaxS * kMaxL], s_logc[kMaxS], s_logq[kMaxL];
    for (int i = threadIdx.x; i < S * L; i += blockDim.x) s_logC[i] = aux[ao.logC + i];
    for (int i = threadIdx.x; i < S; i += blockDim.x) s_logc[i] = aux[ao.logc + i];
    for (int i = threadIdx.x; i < L; i += blockDim.x) s_logq[i] = aux[ao.logq + i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    NodeCond c;
    node_conditionals(S, L, Le, N, n, logE, s_logC, s_logc, c);
    float *x1 = X1 + (size_t)n * (S + L);
    for (int s = 0; s < S; ++s) x1[s] = expf(c.lngs[s]);
    for (int l = 0; l < L; ++l) x1[S + l] = expf(c.Ml[l] - s_logq[l]);
    for (int l = 0; l < L; ++l) X2[(size_t)n * L + l] = expf(c.lcgn[l]);
}

// TP[n][l] = sum_m T[n,m] X2[m][l]   (row gather on T itself; B = L is tiny)
__global__ void csr_times_small(int n_rows, int B, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                const float *__restrict__ data, const float *__restrict__ X, float *__restrict__ Y) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)n_rows * B) return;
    const int r = (int)(gid / B), b = (int)(gid % B);
    float acc = 0.f;
    const int32_t beg = indptr[r], end = indptr[r + 1];
    for (int32_t p0 = beg; p0 < end; p0 += 8) {  // eight entries, then their eight gathers, in flight at a time
        int32_t ix[8];
        float dv[8], xv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool in = p0 + k < end;
            ix[k] = in ? __ldg(indices + p0 + k) : 0;
            dv[k] = in ? __ldg(data + p0 + k) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) xv[k] = p0 + k < end ? __ldg(X + (size_t)ix[k] * B + b) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(dv[k], xv[k], acc);
    }
    Y[gid] = acc;
}

// scalars layout (doubles): div | d[L] | KL[S] | tsum[S] | reg | orth_unused | rowsum[K] | glogC[S*L] | ll
struct ScalarOffsets {
    int div, d, KL, tsum, reg, rowsum, glogC, ll, total;
    __host__ __device__ ScalarOffsets(int S, int L, int K) {
        div = 0;
        d = 1;
        KL = d + L;
        tsum = KL + S;
        reg = tsum + S;
        rowsum = reg + 1;
        glogC = rowsum + K;
        ll = glogC + S * L;
        total = ll + 2;
    }
};

// Reductions over the nodes: d_l = sum_m RS[l,m] P(l|m) (trainer.py:175-182), KL_s (utils.py:149-161),
// sum_n target[s,n] and the TPM regulariser value (trainer.py:188-193). Y1 = X1 . T  ([n][S+L]).
__global__ void __launch_bounds__(128) fhmm_node_pass_b(int S, int L, int Le, int N, const float *__restrict__ logE,
                                                        const float *__restrict__ aux, const float *__restrict__ X1,
                                                        const float *__restrict__ X2, const float *__restrict__ Y1,
                                                        double *__restrict__ scal, int K) {
    const ScalarOffsets sc(S, L, K);
    const AuxOffsets ao(S, L);
    __shared__ double acc[2 * kMaxS + kMaxL + 1];
    __shared__ float s_logC[kMaxS * kMaxL], s_logc[kMaxS];
    const int nq = 2 * S + L + 1;
    for (int i = threadIdx.x; i < nq; i += blockDim.x) acc[i] = 0.0;
    for (int i = threadIdx.x; i < S * L; i += blockDim.x) s_logC[i] = aux[ao.logC + i];
    for (int i = threadIdx.x; i < S; i += blockDim.x) s_logc[i] = aux[ao.logc + i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = n < N;
    const int lane = threadIdx.x & 31;
    const float *x1 = X1 + (size_t)(ok ? n : 0) * (S + L);
    const float *y1 = Y1 + (size_t)(ok ? n : 0) * (S + L);
    NodeCond c;
    node_conditionals(S, L, Le, N, ok ? n : 0, logE, s_logC, s_logc, c);
    float cen = 0.f;
    for (int s = 0; s < S; ++s) cen += x1[s];
    const float lcen = logf(cen / (float)S);  // centroid = lse_s(lngs) - log S
    float reg = 0.f;
    for (int s = 0; s < S; ++s) {
        const float ps = ok ? x1[s] : 0.f, tg = ok ? y1[s] : 0.f;
        const float lps = c.lngs[s];  // log domain: finite even where exp() underflows
        float kl = ps > 0.f ? ps * (lps - lcen) : 0.f;
        float rg = tg > 0.f ? tg * (logf(tg) - lps) : 0.f;  // xlogy(target,target) - target*lngs
        reg += rg;
        kl = warp_sum(kl);
        const float ts = warp_sum(tg);
        if (lane == 0) {
            atomicAdd(&acc[s], (double)kl);
            atomicAdd(&acc[S + s], (double)ts);
        }
    }
    reg = warp_sum(reg);
    if (lane == 0) atomicAdd(&acc[2 * S + L], (double)reg);
    for (int l = 0; l < L; ++l) {
        float v = ok ? y1[S + l] * X2[(size_t)n * L + l] : 0.f;
        v = warp_sum(v);
        if (lane == 0) atomicAdd(&acc[2 * S + l], (double)v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        double *dst = i < S ? &scal[sc.KL + i]
                            : i < 2 * S ? &scal[sc.tsum + i - S] : i < 2 * S + L ? &scal[sc.d + i - 2 * S] : &scal[sc.reg];
        atomicAdd(dst, acc[i]);
    }
}

// Gradient of the node-level regularisers w.r.t. logE and logC, summed with the divergence partials:
// writes the total dLoss/dlogE[k][n] into gE and accumulates its row sums and dLoss/dlogC.
__global__ void __launch_bounds__(128)
fhmm_node_pass_c(int S, int L, int Le, int N, int K, int nsplit, const float *__restrict__ logE,
                 const float *__restrict__ aux, const float *__restrict__ X1, const float *__restrict__ X2,
                 const float *__restrict__ Y1, const float *__restrict__ TP, const float *__restrict__ glogE_split,
                 double *__restrict__ scal, float w_excl, float w_orth, float w_tpm, float *__restrict__ gE) {
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    __shared__ float s_logC[kMaxS * kMaxL], s_logc[kMaxS], s_sumg[kMaxS], s_gd[kMaxL], s_d[kMaxL];
    __shared__ double acc[kMaxK + kMaxS * kMaxL];
    for (int i = threadIdx.x; i < S * L; i += blockDim.x) s_logC[i] = aux[ao.logC + i];
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        s_logc[i] = aux[ao.logc + i];
        // sum_n g_lngs[s,n] = (w_orth/S) KL_s - (w_tpm/S) sum_n target[s,n]
        s_sumg[i] = (float)((w_orth * scal[sc.KL + i] - w_tpm * scal[sc.tsum + i]) / (double)S);
    }
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const double d = scal[sc.d + i];
        s_d[i] = (float)d;
        s_gd[i] = L > 1 ? (float)(-w_excl / ((double)L * d)) : 0.f;  // d(-mean_l log d_l) * weight
    }
    for (int i = threadIdx.x; i < K + S * L; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = n < N;
    const int lane = threadIdx.x & 31;
    const int nn = ok ? n : 0;
    NodeCond c;
    node_conditionals(S, L, Le, N, nn, logE, s_logC, s_logc, c);
    const float *x1 = X1 + (size_t)nn * (S + L), *y1 = Y1 + (size_t)nn * (S + L);
    const float *x2 = X2 + (size_t)nn * L, *tp = TP + (size_t)nn * L;
    // dLoss/dMs[s] through lngs = Ms - lse_n(Ms)
    float cen = 0.f;
    for (int s = 0; s < S; ++s) cen += x1[s];
    const float lcen = logf(cen / (float)S);
    float gMs[kMaxS], gMl[kMaxL];
    for (int s = 0; s < S; ++s) {
        const float ps = x1[s];
        const float g_lngs = (ps > 0.f ? w_orth / (float)S * ps * (c.lngs[s] - lcen) : 0.f) - w_tpm / (float)S * y1[s];
        gMs[s] = g_lngs - ps * s_sumg[s];
    }
    // dLoss/dMl[l] through lngc (softmax over n) and lcgn (softmax over l)
    float sum_glcgn = 0.f;
    for (int l = 0; l < L; ++l) sum_glcgn += x2[l] * s_gd[l] * y1[S + l];
    for (int l = 0; l < L; ++l) {
        const float a = s_gd[l] * x1[S + l] * (tp[l] - s_d[l]);
        const float b = x2[l] * s_gd[l] * y1[S + l] - x2[l] * sum_glcgn;
        gMl[l] = L > 1 ? a + b : 0.f;
    }
    // chain to M[s,l,n] = logE[s,e,n] + logC[s,l]:  gM = softmax-weight(Ms) * gMs + softmax-weight(Ml) * gMl
    for (int s = 0; s < S; ++s) {
        const float Ms = c.lngs[s] + s_logc[s];
        float ge_restricted = 0.f;
        for (int l = 0; l < L; ++l) {
            const int k = s * Le + (Le == 1 ? 0 : l);
            const float M = logE[(size_t)k * N + nn] + s_logC[s * L + l];
            float gM = expf(M - Ms) * gMs[s] + expf(M - c.Ml[l]) * gMl[l];
            if (!ok) gM = 0.f;
            const float tot = warp_sum(gM);
            if (lane == 0) atomicAdd(&acc[K + s * L + l], (double)tot);
            if (Le == 1) {
                ge_restricted += gM;
            } else {
                float g = 0.f;
                if (ok) {
                    g = gM;
                    for (int sp = 0; sp < nsplit; ++sp) g += glogE_split[((size_t)sp * K + k) * N + n];
                    gE[(size_t)k * N + n] = g;
                }
                const float rs = warp_sum(g);
                if (lane == 0) atomicAdd(&acc[k], (double)rs);
            }
        }
        if (Le == 1) {
            float g = 0.f;
            if (ok) {
                g = ge_restricted;
                for (int sp = 0; sp < nsplit; ++sp) g += glogE_split[((size_t)sp * K + s) * N + n];
                gE[(size_t)s * N + n] = g;
            }
            const float rs = warp_sum(g);
            if (lane == 0) atomicAdd(&acc[s], (double)rs);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K + S * L; i += blockDim.x)
        atomicAdd(i < K ? &scal[sc.rowsum + i] : &scal[sc.glogC + i - K], acc[i]);
}

// softmax Jacobian of the emission rows: g theta_E[k,n] = glogE[k,n] - E[k,n] * sum_n' glogE[k,n']
__global__ void fhmm_emission_grad_finish(int N, int K, const float *__restrict__ logE, const double *__restrict__ rowsum,
                                          float *__restrict__ gE) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * N) return;
    const int k = (int)(i / N);
    gE[i] = gE[i] - expf(logE[i]) * (float)rowsum[k];
}

// log-likelihood (methods.py:28-40) in closed form: z[s,l,n] = I*(logE[s,e(l),n] - lse_n logE[s,e(l),:]);
// ll = lse_n( logmeanexp_l logmeanexp_s z ). Per-node values in float64, then one CTA reduces.
__global__ void fhmm_ll_nodes(int S, int L, int Le, int N, int I, const float *__restrict__ logE,
                              const double *__restrict__ row_sum, double *__restrict__ v) {
    __shared__ double s_rl[kMaxK];  // lse_n logE[k,:] = log(row_sum[k])
    for (int k = threadIdx.x; k < S * Le; k += blockDim.x) s_rl[k] = log(row_sum[k]);
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double outer_m = -INFINITY, outer[kMaxL];
    for (int l = 0; l < L; ++l) {
        const int e = Le == 1 ? 0 : l;
        double m = -INFINITY;
        for (int s = 0; s < S; ++s)
            m = fmax(m, (double)I * ((double)logE[(size_t)(s * Le + e) * N + n] - s_rl[s * Le + e]));
        double z = 0.0;
        for (int s = 0; s < S; ++s)
            z += exp((double)I * ((double)logE[(size_t)(s * Le + e) * N + n] - s_rl[s * Le + e]) - m);
        outer[l] = m + log(z) - log((double)S);
        outer_m = fmax(outer_m, outer[l]);
        if (Le == 1) {  // restricted: every lineage sees the same emission row, logmeanexp over l is the identity
            for (int l2 = 1; l2 < L; ++l2) outer[l2] = outer[0];
            break;
        }
    }
    double z = 0.0;
    for (int l = 0; l < L; ++l) z += exp(outer[l] - outer_m);
    v[n] = outer_m + log(z) - log((double)L);
}

// two-level float64 log-sum-exp: per-CTA (max, sum) partials, then one small CTA combines them
__global__ void __launch_bounds__(256) lse_partials(int N, const double *__restrict__ v, double2 *__restrict__ part) {
    __shared__ double sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double x = i < N ? v[i] : -INFINITY;
    double m = x;
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    m = sm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmax(m, sm[w]);
    __syncthreads();
    double z = (i < N && m > -INFINITY) ? exp(x - m) : 0.0;
    z = warp_sum(z);
    if (lane == 0) sm[warp] = z;
    __syncthreads();
    if (threadIdx.x == 0) {
        z = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) z += sm[w];
        part[blockIdx.x] = make_double2(m, z);
    }
}

__global__ void __launch_bounds__(256) lse_combine(int nparts, const double2 *__restrict__ part, double *__restrict__ out) {
    __shared__ double sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double m = -INFINITY;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) m = fmax(m, part[i].x);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    m = sm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmax(m, sm[w]);
    __syncthreads();
    double z = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) z += part[i].y * exp(part[i].x - m);
    z = warp_sum(z);
    if (lane == 0) sm[warp] = z;
    __syncthreads();
    if (threadIdx.x == 0) {
        z = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) z += sm[w];
        out[0] = m + log(z);
    }
}

// the 8 reported scalars (trainer.py:139-196): loss, divergence, log_likelihood, sparsity, orthogonality, exclusivity,
// regularisation, reserved
__device__ void write_terms(int S, int L, int K, const double *__restrict__ scal, float sparsity, float w_sparse,
                            float w_excl, float w_orth, float w_tpm, float *__restrict__ terms) {
    const ScalarOffsets sc(S, L, K);
    const float divergence = (float)scal[sc.div];
    float orth = 0.f;
    for (int s0 = 0; s0 < S; ++s0) orth += (float)scal[sc.KL + s0];
    orth /= (float)S;
    float excl = 0.f;
    if (L > 1) {
        for (int l = 0; l < L; ++l) excl -= logf((float)scal[sc.d + l]);
        excl /= (float)L;
    }
    const float reg = (float)(scal[sc.reg] / (double)S);
    terms[0] = divergence + w_sparse * sparsity + w_orth * orth + (L > 1 ? w_excl * excl : 0.f) + w_tpm * reg;
    terms[1] = divergence;
    terms[2] = (float)scal[sc.ll];
    terms[3] = sparsity;
    terms[4] = orth;
    terms[5] = excl;
    terms[6] = reg;
    terms[7] = 0.f;
}

// --------------------------------------------------------------------------------------------------
// Backward of the hidden recurrence. With c_t[s,l] = w_l gG_t[s,l] + glogC[s,l] / (I Hbar[s,l]) the adjoint is
//     gH_t = c_t + A_l gH_{t+1}          (gH_I = 0),
// an affine chain cut into the same blocks of P steps as the forward pass:
//   fhmm_bwd_local (grid = blocks, one warp per lineage): R_t = c_t + A R_{t+1} inside each block, R = 0 beyond it
//   fhmm_small_backward (one CTA):  Y_b = gH_{(b+1)P} = R_{(b+1)P} + A^P Y_{b+1}   (ceil(I/P) dependent steps)
//                                   gH_t = R_t + A^{(b+1)P - t} Y_b                 (all (t,s,l) in parallel)
// then dLoss/dA[l,s,s'] = sum_t H_t[s,l] gH_{t+1}[s',l] as a parallel contraction, the sparsity term, the softmax
// Jacobians of pi, w, A and the 8 reported scalars
//   terms: loss, divergence, log_likelihood, sparsity, orthogonality, exclusivity, regularisation, reserved
__global__ void __launch_bounds__(512)
fhmm_bwd_local(int S, int L, int I, int P, int restricted, int K, const float *__restrict__ small,
               const float *__restrict__ aux, const float *__restrict__ gG, const double *__restrict__ scal,
               float *__restrict__ R, const double *__restrict__ Hlin, float *__restrict__ gw_acc) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    extern __shared__ float sm[];
    const int SL = S * L, SP = S | 1;
    float *s_A = sm;  // [L][S][SP]
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < L * S * S; i += nt) s_A[(i / S) * SP + i % S] = expf(small[so.log_A + i]);
    __syncthreads();
    if (warp >= L) return;
    const int l = warp;
    const bool on = lane < S;
    const int sl = on ? lane * L + l : 0;
    const float base = on ? (float)scal[sc.glogC + sl] / ((float)I * fmaxf(aux[ao.Hbar + sl], FLT_MIN)) : 0.f;
    const float wl = expf(small[so.log_w + l]);
    const float *Arow = s_A + (l * S + (on ? lane : 0)) * SP;
    const int kidx = restricted ? (on ? lane : 0) : sl;
    const int t_lo = blockIdx.x * P, t_hi = min(I, t_lo + P);
    float r_next = 0.f, gw_part = 0.f;  // gw_l = sum_{t,s} gG_t[s,l] H_t[s,l], summed here because the blocks run in parallel
    for (int t = t_hi - 1; t >= t_lo; --t) {
        const float gg = on ? gG[(size_t)t * K + kidx] : 0.f;
        if (on) gw_part = fmaf(gg, (float)Hlin[(size_t)t * SL + sl], gw_part);
        const float c = on ? fmaf(wl, gg, base) : 0.f;
        float a0 = c, a1 = 0.f;
        if (t + 1 < t_hi) {
            int sp = 0;
            for (; sp + 1 < S; sp += 2) {
                const float n0 = __shfl_sync(0xffffffffu, r_next, sp), n1 = __shfl_sync(0xffffffffu, r_next, sp + 1);
                a0 = fmaf(Arow[sp], n0, a0);
                a1 = fmaf(Arow[sp + 1], n1, a1);
            }
            if (sp < S) a0 = fmaf(Arow[sp], __shfl_sync(0xffffffffu, r_next, sp), a0);
        }
        r_next = on ? a0 + a1 : 0.f;
        if (on) R[(size_t)t * SL + sl] = r_next;
    }
    gw_part = warp_sum(gw_part);
    if (lane == 0) atomicAdd(&gw_acc[l], gw_part);
}

__global__ void __launch_bounds__(512)
fhmm_small_backward(int S, int L, int I, int P, int restricted, int K, const float *__restrict__ small,
                    const double *__restrict__ Hlin, const double *__restrict__ Apow, const float *__restrict__ aux,
                    const float *__restrict__ gG, const double *__restrict__ scal, float w_sparse, float w_excl,
                    float w_orth, float w_tpm, float *__restrict__ gH /* in: R, out: gH */, float *__restrict__ g_pi,
                    float *__restrict__ g_w, float *__restrict__ g_A, float *__restrict__ terms,
                    const float *__restrict__ gw_acc) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    const float *log_pi = small + so.log_pi, *log_w = small + so.log_w, *log_Amix = small + so.log_Amix;
    extern __shared__ float sm[];
    const int SL = S * L, SS = S * S, SP = S | 1;
    const int NB = (I + P - 1) / P;
    float *s_pow = sm;                          // [P][L][S][SP]: A^(j+1), rows padded
    float *s_gA = s_pow + (size_t)P * L * S * SP;  // [L][S][S]
    float *s_Y = s_gA + L * SS;                 // [NB][S][L]: Y_b = gH at the first step after block b (0 for the last)
    float *s_w = s_Y + (size_t)NB * SL;         // [L]
    float *s_gw = s_w + L;                      // [L]
    float *s_red = s_gw + L;                    // [blockDim]
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    for (size_t i = tid; i < (size_t)P * L * SS; i += nt) s_pow[(i / S) * SP + i % S] = (float)Apow[i];
    for (int i = tid; i < L; i += nt) {
        s_w[i] = expf(log_w[i]);
        s_gw[i] = gw_acc[i];  // summed by fhmm_bwd_local
    }
    __syncthreads();
    // dependent chain over the blocks, one warp per lineage
    if (warp < L) {
        const int l = warp;
        const bool on = lane < S;
        const int sl = on ? lane * L + l : 0;
        const float *AProw = s_pow + ((size_t)(P - 1) * L * S + (size_t)l * S + (on ? lane : 0)) * SP;
        float y = 0.f;  // Y_{NB-1} = 0
        for (int b = NB - 1; b >= 0; --b) {
            if (on) s_Y[(size_t)b * SL + sl] = y;
            if (b == 0) break;
            // Y_{b-1} = gH_{bP} = R_{bP} + A^{P'} Y_b, P' = steps from bP to (b+1)P (= P; block b is full unless last)
            const int span = min(I, (b + 1) * P) - b * P;  // length of block b
            const float *Arow = (span == P) ? AProw
                                            : s_pow + ((size_t)(span - 1) * L * S + (size_t)l * S + (on ? lane : 0)) * SP;
            float a0 = on ? gH[(size_t)(b * P) * SL + sl] : 0.f, a1 = 0.f;
            int sp = 0;
            for (; sp + 1 < S; sp += 2) {
                const float n0 = __shfl_sync(0xffffffffu, y, sp), n1 = __shfl_sync(0xffffffffu, y, sp + 1);
                a0 = fmaf(Arow[sp], n0, a0);
                a1 = fmaf(Arow[sp + 1], n1, a1);
            }
            if (sp < S) a0 = fmaf(Arow[sp], __shfl_sync(0xffffffffu, y, sp), a0);
            y = on ? a0 + a1 : 0.f;
        }
    }
    __syncthreads();
    // gH_t = R_t + A^{e_b - t} Y_b for every (t, s, l), e_b = end of t's block.
    // Threads are arranged as (group, element): a thread keeps one (s, l) and strides over t.
    {
        const int groups = nt / SL;
        const int gidx = tid / SL, e = tid % SL, s0 = e / L, l = e % L;
        if (gidx < groups) {
            for (int t = gidx; t < I; t += groups) {
                const int b = t / P;
                const int pw = min(I, (b + 1) * P) - t;  // >= 1
                const float *Arow = s_pow + ((size_t)(pw - 1) * L * S + (size_t)l * S + s0) * SP;
                const float *Y = s_Y + (size_t)b * SL;
                const size_t i = (size_t)t * SL + e;
                float a0 = gH[i], a1 = 0.f;
                int sp = 0;
                for (; sp + 1 < S; sp += 2) {
                    a0 = fmaf(Arow[sp], Y[sp * L + l], a0);
                    a1 = fmaf(Arow[sp + 1], Y[(sp + 1) * L + l], a1);
                }
                if (sp < S) a0 = fmaf(Arow[sp], Y[sp * L + l], a0);
                gH[i] = a0 + a1;
            }
        }
        __syncthreads();
    }
    // theta_pi: glogpi = pi * gH_0, softmax over s (per lineage)
    if (warp < L) {  // one warp per lineage, lane = state
        const int l = warp;
        const bool on = lane < S;
        const float pi = on ? expf(log_pi[lane * L + l]) : 0.f;
        const float pg = on ? pi * gH[lane * L + l] : 0.f;
        const float dot = warp_sum(pg);
        if (on) g_pi[lane * L + l] = pg - pi * dot;
    }
    // gA[l,s,s'] = sum_{t=0}^{I-2} H_t[s,l] * gH_{t+1}[s',l]: chunks of t are staged in shared memory (coalesced),
    // every output keeps its own accumulator
    {
        float *s_Hc = s_pow + (size_t)L * S * SP;  // powers A^2.. are no longer needed; A^1 (first L*S*SP floats) is
        const int TC = P > 1 ? (int)min((size_t)64, ((size_t)(P - 1) * L * S * SP) / (2 * (size_t)SL)) : 0;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};  // up to 4 outputs per thread (L*S*S <= 4 * blockDim)
        if (TC >= 2 && L * SS <= 4 * nt) {
            float *s_Gc = s_Hc + (size_t)TC * SL;
            for (int t0 = 0; t0 + 1 < I; t0 += TC) {
                const int len = min(TC, I - 1 - t0);
                for (int i = tid; i < len * SL; i += nt) {
                    s_Hc[i] = (float)Hlin[(size_t)t0 * SL + i];
                    s_Gc[i] = gH[(size_t)(t0 + 1) * SL + i];
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = tid + q * nt;
                    if (i < L * SS) {
                        const int l = i / SS, s0 = (i / S) % S, sp = i % S;
                        float a = acc[q];
                        for (int t = 0; t < len; ++t) a = fmaf(s_Hc[t * SL + s0 * L + l], s_Gc[t * SL + sp * L + l], a);
                        acc[q] = a;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (tid + q * nt < L * SS) s_gA[tid + q * nt] = acc[q];
        } else {
            for (int i = tid; i < L * SS; i += nt) {
                const int l = i / SS, s0 = (i / S) % S, sp = i % S;
                float a0 = 0.f;
                for (int t = 0; t + 1 < I; ++t)
                    a0 = fmaf((float)Hlin[(size_t)t * SL + s0 * L + l], gH[(size_t)(t + 1) * SL + sp * L + l], a0);
                s_gA[i] = a0;
            }
        }
    }
    __syncthreads();
    // theta_A: glogA = A*gA + sparsity part on the diagonal; softmax over s'
    for (int r = tid; r < L * S; r += nt) {
        const int l = r / S, s0 = r % S;
        const float *Ar = s_pow + (size_t)r * SP;  // A^1
        const float diag = -w_sparse / (float)S * s_w[l] * Ar[s0] / expf(log_Amix[s0 * S + s0]);
        float sum = 0.f;
        for (int sp = 0; sp < S; ++sp) sum += Ar[sp] * s_gA[(size_t)r * S + sp] + (sp == s0 ? diag : 0.f);
        for (int sp = 0; sp < S; ++sp) {
            const float gl = Ar[sp] * s_gA[(size_t)r * S + sp] + (sp == s0 ? diag : 0.f);
            g_A[(size_t)r * S + sp] = gl - Ar[sp] * sum;
        }
    }
    if (tid == 0) {
        // theta_w: glogw_l = w_l*gw_l + sum_s glogC[s,l] + sparsity part
        float glw[kMaxL], sum = 0.f;
        for (int l = 0; l < L; ++l) {
            float g = s_w[l] * s_gw[l];
            for (int s0 = 0; s0 < S; ++s0) {
                g += (float)scal[sc.glogC + s0 * L + l];
                g += -w_sparse / (float)S * s_w[l] * s_pow[(size_t)(l * S + s0) * SP + s0] / expf(log_Amix[s0 * S + s0]);
            }
            glw[l] = g;
            sum += g;
        }
        for (int l = 0; l < L; ++l) g_w[l] = glw[l] - s_w[l] * sum;
        write_terms(S, L, K, scal, aux[ao.sparsity], w_sparse, w_excl, w_orth, w_tpm, terms);
    }
}

// --------------------------------------------------------------------------------------------------
// The same backward as two multi-CTA kernels, one CTA per block of P steps (default; the blocks are independent apart
// from the short chain over block boundaries, and one SM is slower at this than 23):
//   fhmm_bwd_blocks   R_t = c_t + A R_{t+1} inside block b out of shared memory (a warp per lineage), the block's share
//                     of gw; the CTA that finishes last runs the chain Y_b = R_{(b+1)P} + A^P Y_{b+1} and adds the gw
//                     shares in block order
//   fhmm_bwd_expand   gH_t = R_t + A^{e_b - t} Y_b for the block, the block's share of gA = sum_t H_t (x) gH_{t+1};
//                     the CTA that finishes last adds the gA shares in block order and applies the softmax Jacobians of
//                     pi, w, A and writes the 8 reported scalars
__global__ void __launch_bounds__(512)
fhmm_bwd_blocks(int S, int L, int I, int P, int restricted, int K, const float *__restrict__ small,
                const float *__restrict__ aux, const float *__restrict__ gG, const double *__restrict__ scal,
                const double *__restrict__ Hlin, const double *__restrict__ Apow, float *__restrict__ R,
                float *__restrict__ gwPart, float *__restrict__ Yblk, float *__restrict__ gw_out,
                int *__restrict__ ticket) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    extern __shared__ float sm[];
    const int SL = S * L, SS = S * S, SP = S | 1, NB = (I + P - 1) / P;
    float *s_A = sm;                          // [L][S][SP]: A
    float *s_AP = s_A + L * S * SP;           // [L][S][SP]: A^P
    float *s_AQ = s_AP + L * S * SP;          // [L][S][SP]: A^(length of the last block)
    float *s_gg = s_AQ + L * S * SP;          // [P][K]
    float *s_h = s_gg + P * K;                // [P][SL]
    float *s_r0 = s_h + P * SL;               // [NB][SL] (last CTA): R at the first step of every block
    __shared__ int s_last;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x, t_lo = b * P, t_hi = min(I, t_lo + P), len = t_hi - t_lo;
    const int last_len = I - (NB - 1) * P;
    for (int i = tid; i < L * SS; i += nt) {
        const int o = (i / S) * SP + i % S;
        s_A[o] = (float)Apow[i];
        s_AP[o] = (float)Apow[(size_t)(P - 1) * L * SS + i];
        s_AQ[o] = (float)Apow[(size_t)(last_len - 1) * L * SS + i];
    }
#pragma unroll 4
    for (int i = tid; i < len * K; i += nt) s_gg[i] = gG[(size_t)t_lo * K + i];
#pragma unroll 4
    for (int i = tid; i < len * SL; i += nt) s_h[i] = (float)Hlin[(size_t)t_lo * SL + i];
    // (the per-lane constants are fetched in the same round trip as the staging loads)
    const bool on = lane < S && warp < L;
    const int sl = on ? lane * L + warp : 0;
    const float base = on ? (float)scal[sc.glogC + sl] / ((float)I * fmaxf(aux[ao.Hbar + sl], FLT_MIN)) : 0.f;
    const float wl = expf(small[so.log_w + (warp < L ? warp : 0)]);
    __syncthreads();
    if (warp < L) {
        const int l = warp;
        const float *Arow = s_A + (l * S + (on ? lane : 0)) * SP;
        const int kidx = restricted ? (on ? lane : 0) : sl;
        float r_next = 0.f, gw_part = 0.f;  // gw_l = sum_{t,s} gG_t[s,l] H_t[s,l]
        for (int j = len - 1; j >= 0; --j) {
            const float gg = on ? s_gg[j * K + kidx] : 0.f;
            if (on) gw_part = fmaf(gg, s_h[j * SL + sl], gw_part);
            float a0 = on ? fmaf(wl, gg, base) : 0.f, a1 = 0.f;
            if (j + 1 < len) {
                int sp = 0;
                for (; sp + 1 < S; sp += 2) {
                    const float n0 = __shfl_sync(0xffffffffu, r_next, sp), n1 = __shfl_sync(0xffffffffu, r_next, sp + 1);
                    a0 = fmaf(Arow[sp], n0, a0);
                    a1 = fmaf(Arow[sp + 1], n1, a1);
                }
                if (sp < S) a0 = fmaf(Arow[sp], __shfl_sync(0xffffffffu, r_next, sp), a0);
            }
            r_next = on ? a0 + a1 : 0.f;
            if (on) R[(size_t)(t_lo + j) * SL + sl] = r_next;
        }
        gw_part = warp_sum(gw_part);
        if (lane == 0) gwPart[b * L + l] = gw_part;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1) == NB - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll 4
    for (int i = tid; i < NB * SL; i += nt) s_r0[i] = __ldcg(R + (size_t)(i / SL) * P * SL + i % SL);
    for (int l = tid; l < L; l += nt) {
        gw_out[l] = ordered_sum_ldcg(gwPart + l, NB, (size_t)L);
    }
    __syncthreads();
    if (warp < L) {  // Y_b = gH at the first step after block b (0 for the last block)
        const int l = warp;
        const bool on = lane < S;
        const int sl = on ? lane * L + l : 0;
        float y = 0.f;
        for (int q = NB - 1; q >= 0; --q) {
            if (on) Yblk[(size_t)q * SL + sl] = y;
            if (q == 0) break;
            // Y_{q-1} = gH_{qP} = R_{qP} + A^{length of block q} Y_q
            const float *Arow = (q == NB - 1 ? s_AQ : s_AP) + (l * S + (on ? lane : 0)) * SP;
            float a0 = on ? s_r0[q * SL + sl] : 0.f, a1 = 0.f;
            int sp = 0;
            for (; sp + 1 < S; sp += 2) {
                const float n0 = __shfl_sync(0xffffffffu, y, sp), n1 = __shfl_sync(0xffffffffu, y, sp + 1);
                a0 = fmaf(Arow[sp], n0, a0);
                a1 = fmaf(Arow[sp + 1], n1, a1);
            }
            if (sp < S) a0 = fmaf(Arow[sp], __shfl_sync(0xffffffffu, y, sp), a0);
            y = on ? a0 + a1 : 0.f;
        }
    }
}

__global__ void __launch_bounds__(256)
fhmm_bwd_expand(int S, int L, int I, int P, int restricted, int K, const float *__restrict__ small,
                const double *__restrict__ Hlin, const double *__restrict__ Apow, const float *__restrict__ aux,
                const double *__restrict__ scal, const float *__restrict__ R, const float *__restrict__ Yblk,
                const float *__restrict__ gw_in, float w_sparse, float w_excl, float w_orth, float w_tpm,
                float *__restrict__ gAPart, float *__restrict__ gH0, float *__restrict__ g_pi, float *__restrict__ g_w,
                float *__restrict__ g_A, float *__restrict__ terms, int *__restrict__ ticket, int staged) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    const float *log_pi = small + so.log_pi, *log_w = small + so.log_w, *log_Amix = small + so.log_Amix;
    extern __shared__ double smd[];
    const int SL = S * L, SS = S * S, SP = S | 1, LSS = L * SS, NB = (I + P - 1) / P;
    double *s_scal = smd;                                                        // [sc.total] (last CTA)
    float *s_H = reinterpret_cast<float *>(s_scal + ((sc.total + 1) & ~1));      // [P][SL]
    float *s_g = s_H + P * SL;                                                   // [P + 1][SL]: gH_t, then gH at (b+1)P
    float *s_gA = s_g + (P + 1) * SL;                                            // [L][S][S] (last CTA)
    float *s_A = s_gA + LSS;                                                     // [L][S][SP] (last CTA): A
    float *s_w = s_A + L * S * SP;                                               // [L]
    float *s_amd = s_w + L;                                                      // [S]
    float *s_pw = s_amd + S;                                                     // [P][L][S][S] when staged: A^(j+1)
    __shared__ int s_last;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x, t_lo = b * P, len = min(I, t_lo + P) - t_lo;
    for (int i = tid; i < SL; i += nt) s_g[len * SL + i] = Yblk[(size_t)b * SL + i];
#pragma unroll 4
    for (int i = tid; i < len * SL; i += nt) s_H[i] = (float)Hlin[(size_t)t_lo * SL + i];
    if (staged) {
#pragma unroll 8
        for (int i = tid; i < len * LSS; i += nt) s_pw[i] = (float)Apow[i];
    }
    __syncthreads();
    for (int o = tid; o < len * SL; o += nt) {
        const int j = o / SL, e = o - j * SL, s0 = e / L, l = e - s0 * L;
        const int pw = len - j;  // steps to the end of the block, >= 1
        const size_t arow = (size_t)(pw - 1) * LSS + (size_t)l * SS + (size_t)s0 * S;
        const float *Y = s_g + len * SL + l;
        float a0 = R[(size_t)t_lo * SL + o], a1 = 0.f;
        int sp = 0;
        if (staged) {
            const float *Arow = s_pw + arow;
            for (; sp + 1 < S; sp += 2) {
                a0 = fmaf(Arow[sp], Y[sp * L], a0);
                a1 = fmaf(Arow[sp + 1], Y[(sp + 1) * L], a1);
            }
            if (sp < S) a0 = fmaf(Arow[sp], Y[sp * L], a0);
        } else {
            const double *Arow = Apow + arow;
            for (; sp + 1 < S; sp += 2) {
                a0 = fmaf((float)Arow[sp], Y[sp * L], a0);
                a1 = fmaf((float)Arow[sp + 1], Y[(sp + 1) * L], a1);
            }
            if (sp < S) a0 = fmaf((float)Arow[sp], Y[sp * L], a0);
        }
        s_g[o] = a0 + a1;
        if (t_lo + j == 0) gH0[e] = a0 + a1;
    }
    __syncthreads();
    // the block's share of gA[l,s,s'] = sum_{t=0}^{I-2} H_t[s,l] gH_{t+1}[s',l]
    {
        const int jmax = min(len, I - 1 - t_lo);  // steps t of this block with t <= I - 2
        for (int o = tid; o < LSS; o += nt) {
            const int l = o / SS, s0 = (o / S) % S, sp = o % S;
            const float *h = s_H + s0 * L + l, *g = s_g + SL + sp * L + l;
            float a0 = 0.f, a1 = 0.f;
            int j = 0;
            for (; j + 1 < jmax; j += 2) {
                a0 = fmaf(h[j * SL], g[j * SL], a0);
                a1 = fmaf(h[(j + 1) * SL], g[(j + 1) * SL], a1);
            }
            if (j < jmax) a0 = fmaf(h[j * SL], g[j * SL], a0);
            gAPart[(size_t)b * LSS + o] = a0 + a1;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1) == NB - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last CTA: add the shares in block order, then the softmax Jacobians and the reported scalars
    for (int i = tid; i < sc.total; i += nt) s_scal[i] = scal[i];
    for (int o = tid; o < LSS; o += nt) {
        s_gA[o] = ordered_sum_ldcg(gAPart + o, NB, (size_t)LSS);
        s_A[(o / S) * SP + o % S] = (float)Apow[o];
    }
    for (int i = tid; i < L; i += nt) s_w[i] = expf(log_w[i]);
    for (int i = tid; i < S; i += nt) s_amd[i] = expf(log_Amix[i * S + i]);
    __syncthreads();
    if (warp < L) {  // theta_pi: glogpi = pi * gH_0, softmax over s (per lineage)
        const int l = warp;
        const bool on = lane < S;
        const float pi = on ? expf(log_pi[lane * L + l]) : 0.f;
        const float pg = on ? pi * __ldcg(gH0 + lane * L + l) : 0.f;
        const float dot = warp_sum(pg);
        if (on) g_pi[lane * L + l] = pg - pi * dot;
    }
    for (int r = tid; r < L * S; r += nt) {  // theta_A: glogA = A*gA + sparsity part on the diagonal, softmax over s'
        const int l = r / S, s0 = r % S;
        const float *Ar = s_A + (size_t)r * SP;
        const float diag = -w_sparse / (float)S * s_w[l] * Ar[s0] / s_amd[s0];
        float sum = 0.f;
        for (int sp = 0; sp < S; ++sp) sum += Ar[sp] * s_gA[(size_t)r * S + sp] + (sp == s0 ? diag : 0.f);
        for (int sp = 0; sp < S; ++sp) {
            const float gl = Ar[sp] * s_gA[(size_t)r * S + sp] + (sp == s0 ? diag : 0.f);
            g_A[(size_t)r * S + sp] = gl - Ar[sp] * sum;
        }
    }
    if (tid == 0) {  // theta_w: glogw_l = w_l*gw_l + sum_s glogC[s,l] + sparsity part
        float glw[kMaxL], sum = 0.f;
        for (int l = 0; l < L; ++l) {
            float g = s_w[l] * gw_in[l];
            for (int s0 = 0; s0 < S; ++s0) {
                g += (float)s_scal[sc.glogC + s0 * L + l];
                g += -w_sparse / (float)S * s_w[l] * s_A[(size_t)(l * S + s0) * SP + s0] / s_amd[s0];
            }
            glw[l] = g;
            sum += g;
        }
        for (int l = 0; l < L; ++l) g_w[l] = glw[l] - s_w[l] * sum;
        write_terms(S, L, K, s_scal, aux[ao.sparsity], w_sparse, w_excl, w_orth, w_tpm, terms);
    }
}

// --------------------------------------------------------------------------------------------------
// fhmm_bwd_local + fhmm_small_backward as ONE CTA of 1,024 threads that works out of shared memory. The two kernels
// above walk R, H and gG in global memory inside loops whose iterations each wait for an L2 round trip (24 + 104 us at
// config 3 with every SM but one idle); here A^1..A^P, R (-> gH), H, gG and the node-level scalars are staged once with
// all loads in flight, and every later phase reads shared memory:
//   0  stage                      1  R_t = c_t + A R_{t+1} inside the blocks (a warp per (block, lineage))
//   2  Y_b chain over the blocks  3  gH_t = R_t + A^{e_b - t} Y_b in place
//   4  g_pi                       5  gA = sum_t H_t (x) gH_{t+1} (the t range split over the idle threads, fixed order)
//   6  softmax Jacobians of A, w and the 8 reported scalars
// Used when fused_bwd_smem() fits (config 3: 224 KB); larger shapes keep the two-kernel path.
__host__ __device__ inline size_t fused_bwd_floats(int S, int L, int I, int P, int K, int nt) {
    const size_t SL = (size_t)S * L, NB = (size_t)(I + P - 1) / P;
    return (size_t)P * L * S * (S | 1) + 2 * (size_t)I * SL + ((size_t)K == SL ? 0 : (size_t)I * K) + NB * SL +
           (size_t)L * S * S + SL + 2 * (size_t)L + S + (size_t)nt;
}

__global__ void __launch_bounds__(1024)
fhmm_backward_fused(int S, int L, int I, int P, int restricted, int K, const float *__restrict__ small,
                    const double *__restrict__ Hlin, const double *__restrict__ Apow, const float *__restrict__ aux,
                    const float *__restrict__ gG, const double *__restrict__ scal, float w_sparse, float w_excl,
                    float w_orth, float w_tpm, float *__restrict__ g_pi, float *__restrict__ g_w,
                    float *__restrict__ g_A, float *__restrict__ terms) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    const float *log_pi = small + so.log_pi, *log_w = small + so.log_w, *log_Amix = small + so.log_Amix;
    extern __shared__ double smd[];
    const int SL = S * L, SS = S * S, SP = S | 1, LSS = L * SS;
    const int NB = (I + P - 1) / P;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    double *s_scal = smd;                                            // [sc.total]
    float *s_pow = reinterpret_cast<float *>(s_scal + ((sc.total + 1) & ~1));  // [P][L][S][SP]: A^(j+1), rows padded
    float *s_R = s_pow + (size_t)P * L * S * SP;                     // [I][SL]: R, then gH
    float *s_H = s_R + (size_t)I * SL;                               // [I][SL]
    float *s_gG = (K == SL) ? s_R : s_H + (size_t)I * SL;            // [I][K] (in place of R when the indices coincide)
    float *s_Y = s_H + (size_t)I * SL + (K == SL ? 0 : (size_t)I * K);  // [NB][SL]
    float *s_gA = s_Y + (size_t)NB * SL;                             // [L][S][S]
    float *s_base = s_gA + LSS;                                      // [SL]
    float *s_w = s_base + SL;                                        // [L]
    float *s_gw = s_w + L;                                           // [L]
    float *s_amd = s_gw + L;                                         // [S]: A_mix[s,s]
    float *s_red = s_amd + S;                                        // [nt]

    // ---- 0: stage
#pragma unroll 4
    for (int i = tid; i < sc.total; i += nt) s_scal[i] = scal[i];
#pragma unroll 4
    for (int i = tid; i < P * LSS; i += nt) s_pow[(i / S) * SP + i % S] = (float)Apow[i];
#pragma unroll 8
    for (int i = tid; i < I * SL; i += nt) s_H[i] = (float)Hlin[i];
#pragma unroll 8
    for (int i = tid; i < I * K; i += nt) s_gG[i] = gG[i];
    for (int i = tid; i < L; i += nt) {
        s_w[i] = expf(log_w[i]);
        s_gw[i] = 0.f;
    }
    for (int i = tid; i < S; i += nt) s_amd[i] = expf(log_Amix[i * S + i]);
    for (int i = tid; i < SL; i += nt)
        s_base[i] = (float)scal[sc.glogC + i] / ((float)I * fmaxf(aux[ao.Hbar + i], FLT_MIN));
    __syncthreads();

    // ---- 1: R inside the blocks, gw_l = sum_{t,s} gG_t[s,l] H_t[s,l]
    for (int task = warp; task < NB * L; task += nwarps) {
        const int b = task / L, l = task - b * L;
        const bool on = lane < S;
        const int sl = on ? lane * L + l : 0;
        const float base = on ? s_base[sl] : 0.f, wl = s_w[l];
        const float *Arow = s_pow + (size_t)(l * S + (on ? lane : 0)) * SP;  // A^1
        const int kidx = restricted ? (on ? lane : 0) : sl;
        const int t_lo = b * P, t_hi = min(I, t_lo + P);
        float r_next = 0.f, gw_part = 0.f;
        for (int t = t_hi - 1; t >= t_lo; --t) {
            const float gg = on ? s_gG[(size_t)t * K + kidx] : 0.f;
            if (on) gw_part = fmaf(gg, s_H[(size_t)t * SL + sl], gw_part);
            float a0 = on ? fmaf(wl, gg, base) : 0.f, a1 = 0.f;
            if (t + 1 < t_hi) {
                int sp = 0;
                for (; sp + 1 < S; sp += 2) {
                    const float n0 = __shfl_sync(0xffffffffu, r_next, sp), n1 = __shfl_sync(0xffffffffu, r_next, sp + 1);
                    a0 = fmaf(Arow[sp], n0, a0);
                    a1 = fmaf(Arow[sp + 1], n1, a1);
                }
                if (sp < S) a0 = fmaf(Arow[sp], __shfl_sync(0xffffffffu, r_next, sp), a0);
            }
            r_next = on ? a0 + a1 : 0.f;
            __syncwarp();  // K == SL: every lane has read its gG entry before the row is overwritten
            if (on) s_R[(size_t)t * SL + sl] = r_next;
        }
        gw_part = warp_sum(gw_part);
        if (lane == 0) atomicAdd(&s_gw[l], gw_part);
    }
    __syncthreads();

    // ---- 2: Y_b = gH at the first step after block b, one warp per lineage
    if (warp < L) {
        const int l = warp;
        const bool on = lane < S;
        const int sl = on ? lane * L + l : 0;
        const float *AProw = s_pow + ((size_t)(P - 1) * L * S + (size_t)l * S + (on ? lane : 0)) * SP;
        float y = 0.f;
        for (int b = NB - 1; b >= 0; --b) {
            if (on) s_Y[(size_t)b * SL + sl] = y;
            if (b == 0) break;
            const int span = min(I, (b + 1) * P) - b * P;
            const float *Arow = (span == P) ? AProw
                                            : s_pow + ((size_t)(span - 1) * L * S + (size_t)l * S + (on ? lane : 0)) * SP;
            float a0 = on ? s_R[(size_t)(b * P) * SL + sl] : 0.f, a1 = 0.f;
            int sp = 0;
            for (; sp + 1 < S; sp += 2) {
                const float n0 = __shfl_sync(0xffffffffu, y, sp), n1 = __shfl_sync(0xffffffffu, y, sp + 1);
                a0 = fmaf(Arow[sp], n0, a0);
                a1 = fmaf(Arow[sp + 1], n1, a1);
            }
            if (sp < S) a0 = fmaf(Arow[sp], __shfl_sync(0xffffffffu, y, sp), a0);
            y = on ? a0 + a1 : 0.f;
        }
    }
    __syncthreads();

    // ---- 3: gH_t = R_t + A^{e_b - t} Y_b in place; a thread keeps one (s, l) and strides over t
    {
        const int groups = nt / SL;  // SL <= 512
        const int gidx = tid / SL, e = tid - gidx * SL, s0 = e / L, l = e - s0 * L;
        if (gidx < groups) {
            int b = gidx / P, e_b = min(I, (b + 1) * P);
            for (int t = gidx; t < I; t += groups) {
                while (t >= e_b) {
                    ++b;
                    e_b = min(I, (b + 1) * P);
                }
                const int pw = e_b - t;  // >= 1
                const float *Arow = s_pow + ((size_t)(pw - 1) * L * S + (size_t)l * S + s0) * SP;
                const float *Y = s_Y + (size_t)b * SL + l;
                float a0 = s_R[(size_t)t * SL + e], a1 = 0.f;
                int sp = 0;
                for (; sp + 1 < S; sp += 2) {
                    a0 = fmaf(Arow[sp], Y[sp * L], a0);
                    a1 = fmaf(Arow[sp + 1], Y[(sp + 1) * L], a1);
                }
                if (sp < S) a0 = fmaf(Arow[sp], Y[sp * L], a0);
                s_R[(size_t)t * SL + e] = a0 + a1;
            }
        }
    }
    __syncthreads();

    // ---- 4: theta_pi: glogpi = pi * gH_0, softmax over s (per lineage)
    if (warp < L) {
        const int l = warp;
        const bool on = lane < S;
        const float pi = on ? expf(log_pi[lane * L + l]) : 0.f;
        const float pg = on ? pi * s_R[lane * L + l] : 0.f;
        const float dot = warp_sum(pg);
        if (on) g_pi[lane * L + l] = pg - pi * dot;
    }
    // ---- 5: gA[l,s,s'] = sum_{t=0}^{I-2} H_t[s,l] gH_{t+1}[s',l]; `parts` slices of t per output, added in order
    {
        const int parts = max(1, min(8, nt / LSS));
        const int T1 = I - 1, per = (T1 + parts - 1) / parts;
        for (int base_o = 0; base_o < LSS; base_o += (parts > 1 ? LSS : nt)) {
            // parts > 1: one sweep (LSS * parts <= nt); parts == 1: outputs strided over the threads
            const int part = parts > 1 ? tid / LSS : 0;
            const int o = parts > 1 ? tid - part * LSS : base_o + tid;
            float a0 = 0.f, a1 = 0.f;
            const bool live = part < parts && o < LSS;
            if (live) {
                const int l = o / SS, s0 = (o / S) % S, sp = o % S;
                const float *h = s_H + s0 * L + l, *g = s_R + SL + sp * L + l;
                const int t0 = part * per, t1 = min(T1, t0 + per);
                int t = t0;
                for (; t + 1 < t1; t += 2) {
                    a0 = fmaf(h[(size_t)t * SL], g[(size_t)t * SL], a0);
                    a1 = fmaf(h[(size_t)(t + 1) * SL], g[(size_t)(t + 1) * SL], a1);
                }
                if (t < t1) a0 = fmaf(h[(size_t)t * SL], g[(size_t)t * SL], a0);
            }
            if (parts > 1) {
                s_red[tid] = a0 + a1;
                __syncthreads();
                if (tid < LSS) {
                    float a = 0.f;
                    for (int q = 0; q < parts; ++q) a += s_red[q * LSS + tid];
                    s_gA[tid] = a;
                }
            } else if (live) {
                s_gA[o] = a0 + a1;
            }
        }
    }
    __syncthreads();
    // ---- 6: theta_A: glogA = A*gA + sparsity part on the diagonal, softmax over s'; theta_w; the reported scalars
    for (int r = tid; r < L * S; r += nt) {
        const int l = r / S, s0 = r % S;
        const float *Ar = s_pow + (size_t)r * SP;  // A^1
        const float diag = -w_sparse / (float)S * s_w[l] * Ar[s0] / s_amd[s0];
        float sum = 0.f;
        for (int sp = 0; sp < S; ++sp) sum += Ar[sp] * s_gA[(size_t)r * S + sp] + (sp == s0 ? diag : 0.f);
        for (int sp = 0; sp < S; ++sp) {
            const float gl = Ar[sp] * s_gA[(size_t)r * S + sp] + (sp == s0 ? diag : 0.f);
            g_A[(size_t)r * S + sp] = gl - Ar[sp] * sum;
        }
    }
    if (tid == 0) {
        float glw[kMaxL], sum = 0.f;
        for (int l = 0; l < L; ++l) {
            float g = s_w[l] * s_gw[l];
            for (int s0 = 0; s0 < S; ++s0) {
                g += (float)s_scal[sc.glogC + s0 * L + l];
                g += -w_sparse / (float)S * s_w[l] * s_pow[(size_t)(l * S + s0) * SP + s0] / s_amd[s0];
            }
            glw[l] = g;
            sum += g;
        }
        for (int l = 0; l < L; ++l) g_w[l] = glw[l] - s_w[l] * sum;
        write_terms(S, L, K, s_scal, aux[ao.sparsity], w_sparse, w_excl, w_orth, w_tpm, terms);
    }
}

// --------------------------------------------------------------------------------------------------
// Decoders: filtering / smoothing / viterbi (reference _FHMM.py:153-281). All three reduce to
//   Bmat[t,k] = lse_n(log D[t,n] + logE[k,n]) = log sum_n D[t,n] E[k,n]     (a dense [I x N].[N x K] contraction)
// followed by S x L recurrences in the log domain, written as the reference writes them (including its quirks:
// smoothing uses the LAST row of D for every t, _FHMM.py:209,223; psi is stored as float).
__global__ void __launch_bounds__(256) fhmm_row_max(int N, const float *__restrict__ logE, float *__restrict__ rowmax) {
    const float *x = logE + (size_t)blockIdx.x * N;
    __shared__ float sm[8];
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += 256) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
        rowmax[blockIdx.x] = fmaxf(m, sm[0]);
    }
}

// grid (I, K): Bmat[t][k] = rowmax[k] + log sum_n D[t,n] exp(logE[k,n] - rowmax[k])
__global__ void __launch_bounds__(256) fhmm_bmat(int N, int K, const float *__restrict__ D,
                                                 const float *__restrict__ logE, const float *__restrict__ rowmax,
                                                 float *__restrict__ Bmat) {
    const int t = blockIdx.x, k = blockIdx.y;
    const float *d = D + (size_t)t * N, *e = logE + (size_t)k * N;
    const float mk = rowmax[k];
    double acc = 0.0;
    for (int n = threadIdx.x; n < N; n += 256) acc += (double)(d[n] * expf(e[n] - mk));
    acc = warp_sum(acc);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        acc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += sm[w];
        Bmat[(size_t)t * K + k] = mk + (float)log(acc);
    }
}

// One CTA; thread (s, l) owns one entry of every [S][L] slab. Outputs (all float):
//   log_alpha [I][S][L], log_probs [I][L], log_beta [I][S][L], log_gamma [I][S][L],
//   log_delta [I][S][L], psi [I][S][L], log_max [I][L], best_path [L][I]
__global__ void __launch_bounds__(512)
fhmm_decode_small(int S, int L, int Le, int I, const float *__restrict__ small, const float *__restrict__ Bmat,
                  float *__restrict__ log_alpha, float *__restrict__ log_probs, float *__restrict__ log_beta,
                  float *__restrict__ log_gamma, float *__restrict__ log_delta, float *__restrict__ psi,
                  float *__restrict__ log_max, float *__restrict__ best_path) {
    const SmallOffsets so(S, L);
    const float *log_pi = small + so.log_pi, *log_A = small + so.log_A;
    extern __shared__ float sm[];
    const int SL = S * L, K = S * Le;
    float *s_A = sm;            // [L][S][S] log A
    float *s_v = s_A + L * S * S;  // [2][S][L] ping-pong vector
    float *s_lp = s_v + 2 * SL;    // [L]
    float *s_tot = s_lp + L;       // [L] sum_t log_probs[t][l]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < L * S * S; i += nt) s_A[i] = log_A[i];
    const bool own = tid < SL;
    const int s0 = own ? tid / L : 0, l = own ? tid % L : 0;
    const int kb = s0 * Le + (Le == 1 ? 0 : l);  // emission row of (s, l)
    if (tid < L) s_tot[tid] = 0.f;
    __syncthreads();
    // ---------------- filtering (_FHMM.py:153-187)
    for (int t = 0; t < I; ++t) {
        float *cur = s_v + (t & 1) * SL;
        const float *prev = s_v + ((t + 1) & 1) * SL;
        if (own) {
            float pred;
            if (t == 0) {
                pred = log_pi[tid];
            } else {
                float m = -INFINITY;
                for (int s = 0; s < S; ++s) m = fmaxf(m, prev[s * L + l] + s_A[(l * S + s) * S + s0]);
                float z = 0.f;
                for (int s = 0; s < S; ++s) z += expf(prev[s * L + l] + s_A[(l * S + s) * S + s0] - m);
                pred = m + logf(z);
            }
            cur[tid] = Bmat[(size_t)t * K + kb] + pred;
        }
        __syncthreads();
        if (tid < L) {  // log_probs[t][l] = lse_s log_alpha[t][s][l]
            float m = -INFINITY;
            for (int s = 0; s < S; ++s) m = fmaxf(m, cur[s * L + tid]);
            float z = 0.f;
            for (int s = 0; s < S; ++s) z += expf(cur[s * L + tid] - m);
            const float lp = m + logf(z);
            s_lp[tid] = lp;
            log_probs[(size_t)t * L + tid] = lp;
            s_tot[tid] += lp;
        }
        __syncthreads();
        if (own) {
            const float a = cur[tid] - s_lp[l];
            cur[tid] = a;
            log_alpha[(size_t)t * SL + tid] = a;
        }
        __syncthreads();
    }
    // ---------------- smoothing (_FHMM.py:190-236); B_last = Bmat[I-1] for every step (reference quirk)
    {
        const float *Bl = Bmat + (size_t)(I - 1) * K;
        // log_beta[I-1][s,l] = lse_s'( log A[l,s,s'] + B_last[s',l] + 0 )
        for (int t = I - 1; t >= 0; --t) {
            float *cur = s_v + (t & 1) * SL;
            const float *nxt = s_v + ((t + 1) & 1) * SL;  // log_beta[t+1]
            if (own) {
                float m = -INFINITY;
                for (int sp = 0; sp < S; ++sp) {
                    const float inner = Bl[sp * Le + (Le == 1 ? 0 : l)] + (t == I - 1 ? 0.f : nxt[sp * L + l]);
                    m = fmaxf(m, s_A[(l * S + s0) * S + sp] + inner);
                }
                float z = 0.f;
                for (int sp = 0; sp < S; ++sp) {
                    const float inner = Bl[sp * Le + (Le == 1 ? 0 : l)] + (t == I - 1 ? 0.f : nxt[sp * L + l]);
                    z += expf(s_A[(l * S + s0) * S + sp] + inner - m);
                }
                const float b = m + logf(z);
                cur[tid] = b;
                log_beta[(size_t)t * SL + tid] = b;
                log_gamma[(size_t)t * SL + tid] = b - s_tot[l] + log_alpha[(size_t)t * SL + tid];
            }
            __syncthreads();
        }
    }
    // ---------------- viterbi (_FHMM.py:238-281)
    for (int t = 0; t < I; ++t) {
        float *cur = s_v + (t & 1) * SL;
        const float *prev = s_v + ((t + 1) & 1) * SL;
        if (own) {
            float best = 0.f;
            int arg = 0;
            if (t == 0) {
                best = log_pi[tid];
            } else {
                best = -INFINITY;
                for (int s = 0; s < S; ++s) {  // torch.max returns the first maximal index
                    const float v = prev[s * L + l] + s_A[(l * S + s) * S + s0];
                    if (v > best) {
                        best = v;
                        arg = s;
                    }
                }
            }
            const float dl = Bmat[(size_t)t * K + kb] + best;
            cur[tid] = dl;
            log_delta[(size_t)t * SL + tid] = dl;
            psi[(size_t)t * SL + tid] = (float)arg;  // psi[0] stays 0 as in the reference
        }
        __syncthreads();
        if (tid < L) {
            float m = -INFINITY;
            for (int s = 0; s < S; ++s) m = fmaxf(m, cur[s * L + tid]);
            log_max[(size_t)t * L + tid] = m;
        }
        __syncthreads();
    }
    // back-tracking, one thread per lineage. The reference (:270-279) reads psi[t, best_path_i[0], i], i.e. it
    // follows the chain from the most recently inserted state.
    if (tid < L) {
        const float *last = s_v + ((I - 1) & 1) * SL;
        int cur_state = 0;
        float m = -INFINITY;
        for (int s = 0; s < S; ++s)
            if (last[s * L + tid] > m) {
                m = last[s * L + tid];
                cur_state = s;
            }
        best_path[(size_t)tid * I + (I - 1)] = (float)cur_state;
        for (int t = I - 1; t >= 1; --t) {
            cur_state = (int)psi[(size_t)t * SL + cur_state * L + tid];
            best_path[(size_t)tid * I + (t - 1)] = (float)cur_state;
        }
    }
}

// --------------------------------------------------------------------------------------------------
// LSSM (reference models/_LSSM.py:7-117): ONE latent chain with state-level lineage weights,
//     joint[t,s,l,n] = logE[s,e(l),n] + log w[s,l] + hidden[t,s],      hidden_t = hidden_{t-1} . A.
// The chain is the FHMM recurrence with a single lineage, so the blocked forward / backward kernels above run with
// L = 1; the kernels below mix the lineage weights in and out:
//   forward : G[t,(s,l)] = w[s,l] H_t[s] (G[t,s] = H_t[s] when the emission is shared), C[s,l] = w[s,l] Hbar[s]
//   backward: gG1[t,s] = sum_l w[s,l] gG[t,(s,l)], glogC1[s] = sum_l glogC[s,l], and the softmax Jacobian of theta_w
//   small layout: log_pi[S] | log_w[S*L] | log_A[S*S]
struct LssmSmallOffsets {
    int log_pi, log_w, log_A, total;
    __host__ __device__ LssmSmallOffsets(int S, int L) {
        log_pi = 0;
        log_w = S;
        log_A = log_w + S * L;
        total = log_A + S * S;
    }
};

__global__ void __launch_bounds__(512)
lssm_mix_forward(int S, int L, int I, int restricted, int KP, int w_is_log, const float *__restrict__ theta_w,
                 const float *__restrict__ small1, const float *__restrict__ aux1, const double *__restrict__ Hlin1,
                 float *__restrict__ lsmall, float *__restrict__ G, float *__restrict__ aux) {
    const SmallOffsets so1(S, 1);
    const AuxOffsets ao1(S, 1), ao(S, L);
    const LssmSmallOffsets lo(S, L);
    __shared__ float s_logw[kMaxS * kMaxL], s_logC[kMaxS * kMaxL];
    const int tid = threadIdx.x, nt = blockDim.x, SL = S * L;
    for (int s = tid; s < S; s += nt) {  // log_softmax(theta_w, dim=-1)  (methods.py:12-14)
        const float *x = theta_w + s * L;
        float m = -INFINITY;
        for (int l = 0; l < L; ++l) m = fmaxf(m, x[l]);
        float z = 0.f;
        for (int l = 0; l < L; ++l) z += expf(x[l] - m);
        const float lse = w_is_log ? 0.f : m + logf(z);  // NSSM passes log r[s,l], already normalised over l
        for (int l = 0; l < L; ++l) {
            s_logw[s * L + l] = x[l] - lse;
            lsmall[lo.log_w + s * L + l] = x[l] - lse;
        }
        lsmall[lo.log_pi + s] = small1[so1.log_pi + s];
    }
    for (int i = tid; i < S * S; i += nt) lsmall[lo.log_A + i] = small1[so1.log_A + i];
    __syncthreads();
    for (int e = tid; e < SL; e += nt) {
        const float lc = s_logw[e] + aux1[ao1.logC + e / L];  // log w[s,l] + log Hbar[s]
        s_logC[e] = lc;
        aux[ao.logC + e] = lc;
        aux[ao.Hbar + e] = aux1[ao1.Hbar + e / L];
    }
    if (tid == 0) aux[ao.sparsity] = aux1[ao1.sparsity];
    if (restricted) {
        for (int i = tid; i < I * KP; i += nt) {
            const int t = i / KP, k = i % KP;
            G[i] = k < S ? (float)Hlin1[(size_t)t * S + k] : 0.f;  // sum_l w[s,l] = 1
        }
    } else {
        for (int i = tid; i < I * KP; i += nt) {
            const int t = i / KP, k = i % KP;
            G[i] = k < SL ? (float)Hlin1[(size_t)t * S + k / L] * expf(s_logw[k]) : 0.f;
        }
    }
    __syncthreads();
    for (int s0 = tid; s0 < S; s0 += nt) {
        float m = -INFINITY;
        for (int ll = 0; ll < L; ++ll) m = fmaxf(m, s_logC[s0 * L + ll]);
        float z = 0.f;
        for (int ll = 0; ll < L; ++ll) z += expf(s_logC[s0 * L + ll] - m);
        aux[ao.logc + s0] = m + logf(z);
    }
    for (int ll = tid; ll < L; ll += nt) {
        float m = -INFINITY;
        for (int s0 = 0; s0 < S; ++s0) m = fmaxf(m, s_logC[s0 * L + ll]);
        float z = 0.f;
        for (int s0 = 0; s0 < S; ++s0) z += expf(s_logC[s0 * L + ll] - m);
        aux[ao.logq + ll] = m + logf(z);
    }
}

// joint[t,s,l,n] = logE[s,e(l),n] + log w[s,l] + hidden[t,s]   (_LSSM.py:98-111)
__global__ void lssm_joint_kernel(int S, int L, int Le, int N, int I, const float *__restrict__ logE,
                                  const float *__restrict__ hidden, const float *__restrict__ log_w,
                                  float *__restrict__ joint) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int tsl = blockIdx.y;  // t*S*L + s*L + l
    if (n >= N) return;
    const int l = tsl % L, s = (tsl / L) % S, t = tsl / (S * L);
    const int e = Le == 1 ? 0 : l;
    joint[(size_t)tsl * N + n] = logE[(size_t)(s * Le + e) * N + n] + log_w[s * L + l] + hidden[t * S + s];
}

// grid = S CTAs (one latent state each). Folds the lineage axis of dLoss/dG and dLoss/dlogC onto the single chain
// and finishes the theta_w gradient; CTA 0 writes the 8 scalars.
__global__ void __launch_bounds__(256)
lssm_glue_backward(int S, int L, int I, int restricted, int K, int raw_w, const float *__restrict__ lsmall,
                   const double *__restrict__ Hlin1, const float *__restrict__ gG, const double *__restrict__ scal,
                   const float *__restrict__ aux, float w_sparse, float w_excl, float w_orth, float w_tpm,
                   float *__restrict__ gG1, double *__restrict__ scal1, float *__restrict__ g_w,
                   float *__restrict__ terms) {
    const LssmSmallOffsets lo(S, L);
    const ScalarOffsets sc(S, L, K), sc1(S, 1, S);
    const AuxOffsets ao(S, L);
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float s_w[kMaxL], s_red[8][kMaxL];
    if (tid < L) s_w[tid] = expf(lsmall[lo.log_w + s * L + tid]);
    __syncthreads();
    float acc[kMaxL];
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) acc[l] = 0.f;
    for (int t = tid; t < I; t += blockDim.x) {
        const float h = (float)Hlin1[(size_t)t * S + s];
        if (restricted) {
            gG1[(size_t)t * S + s] = gG[(size_t)t * K + s];  // d/dw vanishes: sum_l w[s,l] is constant
        } else {
            float sum = 0.f;
#pragma unroll
            for (int l = 0; l < kMaxL; ++l)
                if (l < L) {
                    const float g = gG[(size_t)t * K + s * L + l];
                    sum = fmaf(s_w[l], g, sum);
                    acc[l] = fmaf(g, h, acc[l]);
                }
            gG1[(size_t)t * S + s] = sum;
        }
    }
#pragma unroll
    for (int l = 0; l < kMaxL; ++l)
        if (l < L) {
            const float v = warp_sum(acc[l]);
            if (lane == 0) s_red[warp][l] = v;
        }
    __syncthreads();
    if (tid == 0) {
        const int nw = blockDim.x >> 5;
        float glw[kMaxL], tot = 0.f;
        double gc = 0.0;
        for (int l = 0; l < L; ++l) {
            float gp = 0.f;
            for (int w = 0; w < nw; ++w) gp += s_red[w][l];
            const double glc = scal[sc.glogC + s * L + l];
            glw[l] = s_w[l] * gp + (float)glc;
            tot += glw[l];
            gc += glc;
        }
        // raw_w: dLoss/dlog w itself (the NSSM chains it through log r), else the softmax Jacobian of theta_w
        for (int l = 0; l < L; ++l) g_w[s * L + l] = raw_w ? glw[l] : glw[l] - s_w[l] * tot;
        scal1[sc1.glogC + s] = gc;
        if (s == 0) write_terms(S, L, K, scal, aux[ao.sparsity], w_sparse, w_excl, w_orth, w_tpm, terms);
    }
}

// Decoder inputs of the LSSM (_LSSM.py:120-260) in the FHMM decoder's layout: every lineage sees the same pi and A,
// and the state-level weight joins the emission term: Bexp[t,(s,l)] = Bmat[t,k(s,l)] + log w[s,l].
__global__ void lssm_expand_decode(int S, int L, int Le, int I, const float *__restrict__ lsmall,
                                   const float *__restrict__ Bmat, float *__restrict__ small_exp,
                                   float *__restrict__ Bexp) {
    const LssmSmallOffsets lo(S, L);
    const SmallOffsets so(S, L);
    const int SL = S * L, K = S * Le;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    for (int i = gid; i < SL; i += gn) small_exp[so.log_pi + i] = lsmall[lo.log_pi + i / L];
    for (int i = gid; i < L; i += gn) small_exp[so.log_w + i] = 0.f;
    for (int i = gid; i < L * S * S; i += gn) small_exp[so.log_A + i] = lsmall[lo.log_A + i % (S * S)];
    for (int i = gid; i < I * SL; i += gn) {
        const int t = i / SL, e = i % SL, s = e / L, l = e % L;
        Bexp[i] = Bmat[(size_t)t * K + s * Le + (Le == 1 ? 0 : l)] + lsmall[lo.log_w + e];
    }
}

// --------------------------------------------------------------------------------------------------
// NSSM (reference models/_NSSM.py:7-118): one latent chain with NODE-level lineage weights,
//     joint[t,s,l,n] = logE[s,n] + log cw[s',n,l] + hidden[t,s]       (s' = 0 when the weights are shared by states).
// Written as Etil[s,l,n] = E[s,n] cw[s',n,l] = r[s,l] Ehat[s,l,n] with Ehat rows normalised over n and
// r[s,l] = sum_n Etil[s,l,n]  (sum_l r[s,l] = 1), this is the unrestricted LSSM with emission Ehat and lineage
// weights r, so the LSSM path runs unchanged on (log Ehat, log r); the two kernels below build those inputs and
// chain the gradients back to theta_E [S][N] and theta_cw [Sc][N][L].
__global__ void __launch_bounds__(128)
nssm_build(int S, int L, int Sc, int N, const float *__restrict__ theta_cw, const float *__restrict__ logE0,
           float *__restrict__ logcw, float *__restrict__ Etil) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    for (int sc = 0; sc < Sc; ++sc) {
        const float *x = theta_cw + ((size_t)sc * N + n) * L;
        float v[kMaxL], m = -INFINITY;
#pragma unroll
        for (int l = 0; l < kMaxL; ++l)
            if (l < L) {
                v[l] = x[l];
                m = fmaxf(m, v[l]);
            }
        float z = 0.f;
#pragma unroll
        for (int l = 0; l < kMaxL; ++l)
            if (l < L) z += expf(v[l] - m);
        const float lse = m + logf(z);
#pragma unroll
        for (int l = 0; l < kMaxL; ++l)
            if (l < L) {
                v[l] -= lse;  // log_softmax(theta_cw, dim=-1)  (methods.py:12-14)
                logcw[((size_t)sc * N + n) * L + l] = v[l];
            }
        const int s_lo = Sc == 1 ? 0 : sc, s_hi = Sc == 1 ? S : sc + 1;
        for (int s = s_lo; s < s_hi; ++s) {
            const float e = logE0[(size_t)s * N + n];
#pragma unroll
            for (int l = 0; l < kMaxL; ++l)
                if (l < L) Etil[(size_t)(s * L + l) * N + n] = e + v[l];
        }
    }
}

// gEt[k][n] = dLoss/dlog Ehat[k][n] (from the LSSM node pass), glw[k] = dLoss/dlog r[k], rowsum[k] = sum_n gEt[k][n].
// dLoss/dlog Etil[k][n] = gEt[k][n] + Ehat[k][n] (glw[k] - rowsum[k]); its sums over l and over s are the gradients
// w.r.t. logE[s][n] and log cw[s'][n][l]; the weights' softmax Jacobian is applied here, the emission's afterwards.
__global__ void __launch_bounds__(128)
nssm_backward_nodes(int S, int L, int Sc, int N, const float *__restrict__ gEt, const float *__restrict__ logEhat,
                    const float *__restrict__ glw, const double *__restrict__ rowsum, const float *__restrict__ logcw,
                    float *__restrict__ gE0, double *__restrict__ rowsum0, float *__restrict__ g_cw) {
    __shared__ float s_coef[kMaxK];
    __shared__ double s_acc[kMaxS];
    for (int k = threadIdx.x; k < S * L; k += blockDim.x) s_coef[k] = glw[k] - (float)rowsum[k];
    for (int k = threadIdx.x; k < S; k += blockDim.x) s_acc[k] = 0.0;
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = n < N;
    const int nn = ok ? n : 0, lane = threadIdx.x & 31;
    float gcw[kMaxL];
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) gcw[l] = 0.f;
    for (int s = 0; s < S; ++s) {
        float ge = 0.f;
#pragma unroll
        for (int l = 0; l < kMaxL; ++l)
            if (l < L) {
                const size_t idx = (size_t)(s * L + l) * N + nn;
                const float gt = gEt[idx] + expf(logEhat[idx]) * s_coef[s * L + l];
                ge += gt;
                gcw[l] += gt;
            }
        if (ok) gE0[(size_t)s * N + n] = ge;
        const float rs = warp_sum(ok ? ge : 0.f);
        if (lane == 0) atomicAdd(&s_acc[s], (double)rs);
        if (Sc > 1 || s == S - 1) {  // softmax Jacobian of the weights of (s', n)
            const int sc = Sc > 1 ? s : 0;
            float tot = 0.f;
#pragma unroll
            for (int l = 0; l < kMaxL; ++l)
                if (l < L) tot += gcw[l];
#pragma unroll
            for (int l = 0; l < kMaxL; ++l)
                if (l < L) {
                    const size_t idx = ((size_t)sc * N + nn) * L + l;
                    if (ok) g_cw[idx] = gcw[l] - expf(logcw[idx]) * tot;
                    gcw[l] = 0.f;
                }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < S; k += blockDim.x) atomicAdd(&rowsum0[k], s_acc[k]);
}

}  // namespace cy2p

using namespace cy2p;

namespace {

int round_kp(int K) {
    const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64};
    for (int o : opts)
        if (K <= o) return o;
    return -1;
}

struct FhmmWs {
    size_t logE, row_lse, hidden, G, small, aux, gG, split, X1, X2, Y1, TP, scal, llv, Hlin, gH, partials, llpart, Apow, total;
    int KP, nsplit, tchunk, P;
    size_t smem_fwd, smem_bwd, smem_local;
    // emission matrices with more than 24 rows: forward kernel over all rows + gradient slices of 12 rows
    bool two_pass;
    int nsplitA, tchunkA;
    size_t R;
    size_t gwacc;  // [kMaxL] floats: dLoss/dw partial sums of fhmm_bwd_local
    // split path of the small kernels (one CTA per block of P steps): per-block hand-over buffers
    size_t Xblk, HbarPart, Yblk, gAPart, gwPart, gH0, tickets;
};

size_t up(size_t x) { return (x + 255) / 256 * 256; }

int env_flag(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// Block length of the recurrence decomposition: ~sqrt(I), bounded by the shared memory the powers A^1..A^P need
int choose_block(int S, int L, int I, size_t *smem_fwd, size_t *smem_bwd) {
    const size_t budget = 200 * 1024;
    int P = 1;
    while ((P + 1) * (P + 1) <= I) ++P;
    for (;; --P) {
        const size_t NB = (size_t)(I + P - 1) / P;
        const size_t f = sizeof(float) * (size_t)((L * S * S + 2 * S * L + L + S * S + 1) & ~1) +
                         sizeof(double) * ((size_t)P * L * S * S + (P > 1 ? NB * S * L : 0) + 1024);
        const size_t b = sizeof(float) * ((size_t)P * L * S * (S | 1) + (size_t)L * S * S + NB * S * L + 2 * L + 512);
        if ((f <= budget && b <= budget) || P == 1) {
            *smem_fwd = f;
            *smem_bwd = b;
            return P;
        }
    }
}

FhmmWs fhmm_layout(int S, int L, int Le, int N, int I) {
    FhmmWs w;
    const int K = S * Le;
    w.P = choose_block(S, L, I, &w.smem_fwd, &w.smem_bwd);
    w.smem_local = sizeof(float) * (size_t)L * S * (S | 1);
    w.KP = round_kp(K);
    // t-splits: the data pass runs 128-thread CTAs at 4 per SM (J = 4 variants are register-limited), so aim at
    // whole waves of 148 * 4 CTAs (the first capture ran 1.66 waves: 40% of the second wave idle);
    // the G chunk of a split must fit 32 KB of shared memory
    // CY2P_FHMM_TWO_PASS=1 forces the forward + tensor-core-gradient split for small emission matrices too (measurement)
    // (default since round 2: the split path with the tensor-core kernels beats the fused FFMA2 pass at every K —
    // 0.55 vs 0.57 ms per call at K = 10, 0.77 vs 1.44 ms at K = 40, profiles/bench_fhmm_mma_r02.jsonl)
    w.two_pass = w.KP > 24 || env_flag("CY2P_FHMM_TWO_PASS", env_flag("CY2P_FHMM_MMA", 1)) != 0;
    auto split_for = [&](int J, int kp, int ctas_per_sm, int *nsplit_out, int *tchunk_out) {
        const int tiles = (N + 128 * J - 1) / (128 * J);
        const int wave = 148 * ctas_per_sm;
        int nsplit = tiles >= wave ? 1 : wave / tiles;
        if (nsplit < 1) nsplit = 1;
        if (nsplit > I) nsplit = I;
        int tchunk = (I + nsplit - 1) / nsplit;
        const int max_chunk = 8192 / (kp > 0 ? kp : 1);
        if (tchunk > max_chunk) tchunk = max_chunk;
        *nsplit_out = (I + tchunk - 1) / tchunk;
        *tchunk_out = tchunk;
    };
    if (w.two_pass) {
        split_for(2, w.KP, 3, &w.nsplitA, &w.tchunkA);
        split_for(4, 12, 4, &w.nsplit, &w.tchunk);
        if (env_flag("CY2P_FHMM_MMA", 1)) {
            // tensor-core gradients: a CTA of fhmm_dE_mma keeps the G chunk of its t split in shared memory as two
            // tf32 planes; 128 time steps per split keep that at <= 64 KB (3 CTAs per SM)
            if (w.tchunk > 128) w.tchunk = 128;
            if (w.tchunk < 64 && I >= 64) w.tchunk = 64;
            w.nsplit = (I + w.tchunk - 1) / w.tchunk;
            // forward kernel: 128-thread CTAs over 32 * NW cells, ~4 resident per SM: pick the t split whose grid
            // fills whole waves best (chunks of <= 256 time steps: <= 41-90 KB of shared memory for the G planes)
            const int kt = (w.KP + 7) / 8, nw = kt <= 3 ? 4 : 2;
            const int tiles = (N + 32 * nw - 1) / (32 * nw), wave = 148 * 4;
            int best = (I + 255) / 256;
            double best_eff = 0.0;
            for (int ns = (I + 255) / 256; ns <= 12 && ns <= I; ++ns) {
                const long ctas = (long)tiles * ns;
                const double eff = (double)ctas / (double)(((ctas + wave - 1) / wave) * wave);
                if (eff > best_eff + 0.02) best_eff = eff, best = ns;
            }
            w.nsplitA = best;
            w.tchunkA = (I + best - 1) / best;
            w.nsplitA = (I + w.tchunkA - 1) / w.tchunkA;
        }
    } else {
        split_for(w.KP <= 12 ? 4 : 2, w.KP, 4, &w.nsplit, &w.tchunk);
        w.nsplitA = w.tchunkA = 0;
    }
    const int nsplit = w.nsplit;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = up(off + bytes);
        return o;
    };
    w.logE = take(sizeof(float) * (size_t)K * N);
    w.row_lse = take(sizeof(double) * K);
    w.hidden = take(sizeof(float) * (size_t)I * S * L);
    w.G = take(sizeof(float) * (size_t)I * (w.KP > 0 ? w.KP : 1));
    w.small = take(sizeof(float) * SmallOffsets(S, L).total);
    w.aux = take(sizeof(float) * AuxOffsets(S, L).total);
    w.gG = take(sizeof(float) * (size_t)I * K);
    w.split = take(sizeof(float) * (size_t)nsplit * K * N);
    w.X1 = take(sizeof(float) * (size_t)N * (S + L));
    w.X2 = take(sizeof(float) * (size_t)N * L);
    w.Y1 = take(sizeof(float) * (size_t)N * (S + L));
    w.TP = take(sizeof(float) * (size_t)N * L);
    w.scal = take(sizeof(double) * ScalarOffsets(S, L, K).total);
    w.llv = take(sizeof(double) * (size_t)N);
    w.Hlin = take(sizeof(double) * (size_t)I * S * L);
    w.gH = take(sizeof(float) * (size_t)I * S * L);
    w.partials = take(sizeof(float2) * (size_t)K * ((N + kSoftmaxChunk - 1) / kSoftmaxChunk));
    w.llpart = take(sizeof(double2) * (size_t)((N + 255) / 256));
    w.Apow = take(sizeof(double) * (size_t)w.P * L * S * S);
    w.R = w.two_pass ? take(sizeof(float) * (size_t)I * N) : 0;
    w.gwacc = take(sizeof(float) * kMaxL);
    {
        size_t f1, b1;
        const int P1 = choose_block(S, 1, I, &f1, &b1);  // the LSSM runs the same kernels with one lineage
        const int Pmin = w.P < P1 ? w.P : P1;
        const size_t NBmax = (size_t)(I + Pmin - 1) / Pmin;
        w.Xblk = take(sizeof(double) * NBmax * S * L);
        w.HbarPart = take(sizeof(double) * NBmax * S * L);
        w.Yblk = take(sizeof(float) * NBmax * S * L);
        w.gAPart = take(sizeof(float) * NBmax * L * S * S);
        w.gwPart = take(sizeof(float) * NBmax * L);
        w.gH0 = take(sizeof(float) * (size_t)S * L);
        w.tickets = take(sizeof(int) * 4);
    }
    w.total = off;
    return w;
}

// Optional second stream for the dG contraction: dlogE (needed next, by the regulariser-gradient pass) and dG (needed only
// by the backward recurrences) both read r and nothing else, so they run side by side.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;  // fork: recorded on the main stream after r is written; join: after dG
};

template <int KP>
void launch_data_two_pass(int N, int I, int K, const FhmmWs &w, char *ws, const float *logE, const float *G,
                          const float *D, float *pred, float *split, float *gG, double *div_acc, cudaStream_t st,
                          SideStream side = SideStream()) {
    float *R = (float *)(ws + w.R);
    if (env_flag("CY2P_FHMM_MMA", 1)) {  // P, pred, divergence and r: the emission GEMM on the tensor cores
        constexpr int KT = (KP + 7) / 8;
        constexpr int NW = KT <= 3 ? 4 : 2;  // cell tiles per warp (E^ fragments live in registers: 4 KT NW of them)
        const int tpad = (w.tchunkA + 15) / 16 * 16;
        const size_t smem = sizeof(uint32_t) * 2 * (size_t)tpad * mma_ld(KT * 8);
        dim3 grid((unsigned)((N + 32 * NW - 1) / (32 * NW)), (unsigned)w.nsplitA);
        count_launch();
        cudaFuncSetAttribute(fhmm_fwd_mma<KT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fhmm_fwd_mma<KT, NW><<<grid, 128, smem, st>>>(N, I, K, KP, w.tchunkA, logE, G, D, pred, R, div_acc, 1.0f / (float)I);
    } else {
        dim3 gridA((unsigned)((N + 255) / 256), (unsigned)w.nsplitA);
        count_launch();
        fhmm_data_fwd<KP><<<gridA, 128, sizeof(float) * (size_t)w.tchunkA * KP, st>>>(N, I, K, w.tchunkA, logE, G, D, pred,
                                                                                     R, div_acc, 1.0f / (float)I);
    }
    if (D == nullptr) return;  // forward only
    const int use_mma = env_flag("CY2P_FHMM_MMA", 1);
    if (use_mma) {  // the two gradients on the tensor cores (3xTF32), r read twice (mostly out of L2)
        const float inv_I = 1.0f / (float)I;
        if (side.stream) {  // r is complete on the main stream: dG may start on the side stream
            cudaEventRecord(side.fork, st);
            cudaStreamWaitEvent(side.stream, side.fork, 0);
        }
        {
            const int MT = (K + 15) / 16;
            const int tpad = (w.tchunk + 7) / 8 * 8;
            const size_t smem = sizeof(uint32_t) * 2 * (size_t)tpad * (MT * 16 + 8);
            dim3 grid((unsigned)((N + 127) / 128), (unsigned)w.nsplit);
            count_launch();
#define DE_CASE(MT_, NW_)                                                                                           \
    case MT_:                                                                                                       \
        cudaFuncSetAttribute(fhmm_dE_mma<MT_, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        fhmm_dE_mma<MT_, NW_><<<grid, 128, smem, st>>>(N, I, K, K, KP, w.tchunk, logE, G, R, split, inv_I);        \
        break;
            switch (MT) { DE_CASE(1, 4) DE_CASE(2, 4) DE_CASE(3, 4) DE_CASE(4, 4) }
#undef DE_CASE
        }
        {
            const int NT = (K + 7) / 8;
            const size_t smem = sizeof(uint32_t) * 2 * (size_t)NT * 8 * kMmaLdE;
            dim3 grid((unsigned)((N + kMmaCells - 1) / kMmaCells));
            count_launch();
            cudaStream_t sg = side.stream ? side.stream : st;
#define DG_CASE(NT_)                                                                                                \
    case NT_:                                                                                                       \
        cudaFuncSetAttribute(fhmm_dG_mma<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        fhmm_dG_mma<NT_><<<grid, 256, smem, sg>>>(N, I, K, K, logE, R, gG, inv_I);                                  \
        break;
            switch (NT) { DG_CASE(1) DG_CASE(2) DG_CASE(3) DG_CASE(4) DG_CASE(5) DG_CASE(6) DG_CASE(7) DG_CASE(8) }
#undef DG_CASE
            if (side.stream) cudaEventRecord(side.join, side.stream);
        }
        return;
    }
    dim3 gridB((unsigned)((N + 511) / 512), (unsigned)w.nsplit);
    for (int koff = 0; koff < K; koff += 12) {
        const int Ks = K - koff < 12 ? K - koff : 12;
        count_launch();
        fhmm_data_pass<12, 4><<<gridB, 128, sizeof(float) * (size_t)w.tchunk * 12, st>>>(
            N, I, Ks, w.tchunk, logE, G, nullptr, nullptr, split, gG, div_acc, 1.0f / (float)I, R, koff, K, KP);
    }
}

template <int KP, int J>
void launch_data_pass(int N, int I, int K, const FhmmWs &w, const float *logE, const float *G, const float *D,
                      float *pred, float *split, float *gG, double *div_acc, cudaStream_t st) {
    dim3 grid((unsigned)((N + 128 * J - 1) / (128 * J)), (unsigned)w.nsplit);
    const size_t smem = sizeof(float) * (size_t)w.tchunk * KP;
    count_launch();
    fhmm_data_pass<KP, J><<<grid, 128, smem, st>>>(N, I, K, w.tchunk, logE, G, D, pred, split, gG, div_acc,
                                                    1.0f / (float)I, nullptr, 0, K, KP);
}

void dispatch_data_pass(int N, int I, int K, const FhmmWs &w, char *ws, const float *logE, const float *G,
                        const float *D, float *pred, float *split, float *gG, double *div_acc, cudaStream_t st,
                        SideStream side = SideStream()) {
#define CASE(KP_, J_)                                                                         \
    case KP_:                                                                                 \
        launch_data_pass<KP_, J_>(N, I, K, w, logE, G, D, pred, split, gG, div_acc, st);      \
        break;
#define CASE2(KP_)                                                                                \
    case KP_:                                                                                     \
        launch_data_two_pass<KP_>(N, I, K, w, ws, logE, G, D, pred, split, gG, div_acc, st, side); \
        break;
    if (w.two_pass) {
        switch (w.KP) {
            CASE2(2) CASE2(4) CASE2(6) CASE2(8) CASE2(10) CASE2(12) CASE2(16) CASE2(20) CASE2(24)
            CASE2(32) CASE2(40) CASE2(48) CASE2(64)
        }
        return;
    }
    switch (w.KP) {
        CASE(2, 4) CASE(4, 4) CASE(6, 4) CASE(8, 4) CASE(10, 4) CASE(12, 4) CASE(16, 2) CASE(20, 2) CASE(24, 2)
    }
#undef CASE
#undef CASE2
}

int32_t check_dims(int S, int L, int N, int I, int restricted) {
    CY2P_REQUIRE(S >= 1 && L >= 1 && N >= 1 && I >= 1, "S, L, N, I must be >= 1");
    CY2P_REQUIRE(S <= kMaxS && L <= kMaxL, "supported sizes: S <= 32, L <= 16");
    CY2P_REQUIRE(S * (restricted ? 1 : L) <= kMaxK, "supported sizes: S*(restricted ? 1 : L) <= 64");
    return CY2P_OK;
}

// Backward of the hidden recurrence: the fused shared-memory kernel when its staging fits one CTA, else the two-kernel
// path. CY2P_FHMM_FUSED_BWD=0 forces the latter (measurement).
struct SplitBufs {  // FhmmWs::{Xblk, HbarPart, Yblk, gAPart, gwPart, gH0, tickets} resolved to pointers
    double *Xblk, *HbarPart;
    float *Yblk, *gAPart, *gwPart, *gH0;
    int *tickets;
};
SplitBufs split_bufs(const FhmmWs &w, char *ws) {
    return SplitBufs{(double *)(ws + w.Xblk), (double *)(ws + w.HbarPart), (float *)(ws + w.Yblk),
                     (float *)(ws + w.gAPart), (float *)(ws + w.gwPart), (float *)(ws + w.gH0), (int *)(ws + w.tickets)};
}
// The split path needs blocks (P > 1) and the tickets fhmm_small_forward zeroes in the same call.
bool use_split(int P) { return P > 1 && env_flag("CY2P_FHMM_SPLIT_SMALL", 1) != 0; }

int32_t launch_hidden_backward(int S, int L, int I, int P, int restricted, int K, const float *small,
                               const double *Hlin, const double *Apow, const float *aux, const float *gG,
                               const double *scal, float w_sparse, float w_excl, float w_orth, float w_tpm, float *gH,
                               float *g_pi, float *g_w, float *g_A, float *terms, float *gwacc, size_t smem_local,
                               size_t smem_bwd, cudaStream_t st, const SplitBufs &sb) {
    const ScalarOffsets sc(S, L, K);
    if (use_split(P)) {
        const int NB = (I + P - 1) / P, SL = S * L, SP = S | 1;
        const size_t sm1 = sizeof(float) * ((size_t)3 * L * S * SP + (size_t)P * K + (size_t)P * SL + (size_t)NB * SL);
        size_t sm2 = sizeof(double) * (size_t)((sc.total + 1) & ~1) +
                     sizeof(float) * ((size_t)P * SL + (size_t)(P + 1) * SL + (size_t)L * S * S + (size_t)L * S * SP + L + S);
        const size_t pw_bytes = sizeof(float) * (size_t)P * L * S * S;  // A^1..A^P as float, staged when they fit
        const int staged = sm2 + pw_bytes <= 160 * 1024 ? 1 : 0;
        if (staged) sm2 += pw_bytes;
        if (sm1 <= 200 * 1024 && sm2 <= 200 * 1024) {
            CY2P_CUDA(cudaFuncSetAttribute(fhmm_bwd_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
            count_launch();
            fhmm_bwd_blocks<<<NB, L * 32, sm1, st>>>(S, L, I, P, restricted, K, small, aux, gG, scal, Hlin, Apow, gH,
                                                     sb.gwPart, sb.Yblk, gwacc, sb.tickets + 1);
            CY2P_CUDA(cudaFuncSetAttribute(fhmm_bwd_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            count_launch();
            fhmm_bwd_expand<<<NB, 256, sm2, st>>>(S, L, I, P, restricted, K, small, Hlin, Apow, aux, scal, gH, sb.Yblk,
                                                  gwacc, w_sparse, w_excl, w_orth, w_tpm, sb.gAPart, sb.gH0, g_pi, g_w,
                                                  g_A, terms, sb.tickets + 2, staged);
            CY2P_CUDA(cudaGetLastError());
            return CY2P_OK;
        }
    }
    const int nt = 1024;
    const size_t fused = sizeof(double) * (size_t)((sc.total + 1) & ~1) + sizeof(float) * fused_bwd_floats(S, L, I, P, K, nt);
    if (fused <= 227 * 1024 && env_flag("CY2P_FHMM_FUSED_BWD", 1)) {
        CY2P_CUDA(cudaFuncSetAttribute(fhmm_backward_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused));
        count_launch();
        fhmm_backward_fused<<<1, nt, fused, st>>>(S, L, I, P, restricted, K, small, Hlin, Apow, aux, gG, scal, w_sparse,
                                                  w_excl, w_orth, w_tpm, g_pi, g_w, g_A, terms);
        CY2P_CUDA(cudaGetLastError());
        return CY2P_OK;
    }
    const int NB = (I + P - 1) / P;
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_bwd_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_local));
    CY2P_CUDA(cudaMemsetAsync(gwacc, 0, sizeof(float) * kMaxL, st));
    count_launch();
    fhmm_bwd_local<<<NB, L * 32, smem_local, st>>>(S, L, I, P, restricted, K, small, aux, gG, scal, gH, Hlin, gwacc);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_small_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd));
    count_launch();
    fhmm_small_backward<<<1, 512, smem_bwd, st>>>(S, L, I, P, restricted, K, small, Hlin, Apow, aux, gG, scal, w_sparse,
                                                  w_excl, w_orth, w_tpm, gH, g_pi, g_w, g_A, terms, gwacc);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

// forward pieces shared by cy2p_fhmm_forward and cy2p_fhmm_loss_grad
int32_t run_emission(int K, int N, const float *theta_E, const FhmmWs &w, char *ws, cudaStream_t st) {
    float *logE = (float *)(ws + w.logE);
    const int nchunks = (N + kSoftmaxChunk - 1) / kSoftmaxChunk;
    CY2P_CUDA(cudaMemsetAsync(ws + w.row_lse, 0, sizeof(double) * K, st));
    count_launch();
    emission_softmax_partials<<<dim3(K, nchunks), 256, 0, st>>>(N, theta_E, (float2 *)(ws + w.partials));
    count_launch();
    emission_softmax_write<<<dim3(K, nchunks), 256, 0, st>>>(N, theta_E, (const float2 *)(ws + w.partials), logE,
                                                              (double *)(ws + w.row_lse));
    return CY2P_OK;
}

int32_t run_small_forward(int S, int L, int I, int restricted, const float *theta_pi, const float *theta_w,
                          const float *theta_A, const FhmmWs &w, char *ws, cudaStream_t st) {
    if (w.smem_fwd > 220 * 1024 || w.smem_bwd > 220 * 1024) {
        set_error("model too large for the one-CTA recurrence kernels (S=%d, L=%d, I=%d)", S, L, I);
        return CY2P_ERR_INVALID_ARG;
    }
    const SplitBufs sb = split_bufs(w, ws);
    size_t smx = sizeof(float) * ((size_t)w.P * S * L + L + (size_t)S * L);
    const size_t stage_bytes = sizeof(double) * ((size_t)S * L + (size_t)(w.P - 1) * L * S * S);
    const int staged = smx + stage_bytes <= 160 * 1024 ? 1 : 0;
    if (staged) smx += stage_bytes;
    const bool split = use_split(w.P) && smx <= 200 * 1024;
    const size_t hf_bytes = sizeof(float) * (size_t)I * S * L;  // float copy of H in shared memory when it fits
    const int hf = (!split && w.smem_fwd + hf_bytes <= 227 * 1024 && env_flag("CY2P_FHMM_HF_SMEM", 1)) ? 1 : 0;
    const size_t smem = w.smem_fwd + (hf ? hf_bytes : 0);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_small_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_launch();
    fhmm_small_forward<<<1, 1024, smem, st>>>(S, L, I, restricted, w.KP, w.P, theta_pi, theta_w, theta_A,
                                              (float *)(ws + w.small), (float *)(ws + w.hidden), (float *)(ws + w.G),
                                              (float *)(ws + w.aux), (double *)(ws + w.Hlin), (double *)(ws + w.Apow),
                                              hf, split ? sb.Xblk : nullptr, sb.tickets);
    if (split) {
        const int NB = (I + w.P - 1) / w.P;
        CY2P_CUDA(cudaFuncSetAttribute(fhmm_forward_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx));
        count_launch();
        fhmm_forward_expand<<<NB, 256, smx, st>>>(S, L, I, restricted, w.KP, w.P, (const float *)(ws + w.small),
                                                  (const double *)(ws + w.Apow), sb.Xblk, (float *)(ws + w.hidden),
                                                  (double *)(ws + w.Hlin), (float *)(ws + w.G), (float *)(ws + w.aux),
                                                  sb.HbarPart, sb.tickets, staged);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t run_forward(int S, int L, int Le, int N, int I, int restricted, const float *theta_pi, const float *theta_w,
                    const float *theta_E, const float *theta_A, const FhmmWs &w, char *ws, cudaStream_t st) {
    int32_t rc = run_emission(S * Le, N, theta_E, w, ws, st);
    if (rc) return rc;
    return run_small_forward(S, L, I, restricted, theta_pi, theta_w, theta_A, w, ws, st);
}

// ---- LSSM: the FHMM work space (for the node / data passes with L lineages) plus the single-chain pieces
struct LssmWs {
    FhmmWs f;
    size_t small1, aux1, scal1, gG1, G1, Apow1, lsmall, zero, tmp, total;
    int P1, KP1;
    size_t smem_fwd1, smem_bwd1, smem_local1;
};

LssmWs lssm_layout(int S, int L, int Le, int N, int I) {
    LssmWs w;
    w.f = fhmm_layout(S, L, Le, N, I);
    w.P1 = choose_block(S, 1, I, &w.smem_fwd1, &w.smem_bwd1);
    w.smem_local1 = sizeof(float) * (size_t)S * (S | 1);
    w.KP1 = round_kp(S);
    size_t off = w.f.total;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = up(off + bytes);
        return o;
    };
    w.small1 = take(sizeof(float) * SmallOffsets(S, 1).total);
    w.aux1 = take(sizeof(float) * AuxOffsets(S, 1).total);
    w.scal1 = take(sizeof(double) * ScalarOffsets(S, 1, S).total);
    w.gG1 = take(sizeof(float) * (size_t)I * S);
    w.G1 = take(sizeof(float) * (size_t)I * w.KP1);
    w.Apow1 = take(sizeof(double) * (size_t)w.P1 * S * S);
    w.lsmall = take(sizeof(float) * LssmSmallOffsets(S, L).total);
    w.zero = take(sizeof(float) * 4);
    w.tmp = take(sizeof(float) * 16);  // discarded outputs of the single-lineage backward (g_w, terms)
    w.total = off;
    return w;
}

int32_t lssm_check_ws(const LssmWs &w, const void *d_ws, size_t ws_bytes) {
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    return CY2P_OK;
}

// prenorm: logE (in the work space) and theta_w = log w are already normalised (the NSSM path)
int32_t lssm_run_forward(int S, int L, int Le, int N, int I, int restricted, const float *theta_pi, const float *theta_w,
                         const float *theta_E, const float *theta_A, const LssmWs &w, char *ws, cudaStream_t st,
                         int prenorm = 0) {
    int32_t rc = prenorm ? CY2P_OK : run_emission(S * Le, N, theta_E, w.f, ws, st);
    if (rc) return rc;
    if (w.smem_fwd1 > 220 * 1024 || w.smem_bwd1 > 220 * 1024) {
        set_error("model too large for the one-CTA recurrence kernels (S=%d, I=%d)", S, I);
        return CY2P_ERR_INVALID_ARG;
    }
    CY2P_CUDA(cudaMemsetAsync(ws + w.zero, 0, sizeof(float) * 4, st));  // theta_w of the single lineage
    const SplitBufs sb = split_bufs(w.f, ws);
    size_t smx = sizeof(float) * ((size_t)w.P1 * S + 1 + (size_t)S);
    const size_t stage_bytes = sizeof(double) * ((size_t)S + (size_t)(w.P1 - 1) * S * S);
    const int staged = smx + stage_bytes <= 160 * 1024 ? 1 : 0;
    if (staged) smx += stage_bytes;
    const bool split = use_split(w.P1) && smx <= 200 * 1024;
    const size_t hf_bytes = sizeof(float) * (size_t)I * S;
    const int hf = (!split && w.smem_fwd1 + hf_bytes <= 227 * 1024 && env_flag("CY2P_FHMM_HF_SMEM", 1)) ? 1 : 0;
    const size_t smem1 = w.smem_fwd1 + (hf ? hf_bytes : 0);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_small_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    count_launch();
    fhmm_small_forward<<<1, 1024, smem1, st>>>(S, 1, I, 1, w.KP1, w.P1, theta_pi, (const float *)(ws + w.zero), theta_A,
                                               (float *)(ws + w.small1), (float *)(ws + w.f.hidden),
                                               (float *)(ws + w.G1), (float *)(ws + w.aux1), (double *)(ws + w.f.Hlin),
                                               (double *)(ws + w.Apow1), hf, split ? sb.Xblk : nullptr, sb.tickets);
    if (split) {
        const int NB1 = (I + w.P1 - 1) / w.P1;
        CY2P_CUDA(cudaFuncSetAttribute(fhmm_forward_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx));
        count_launch();
        fhmm_forward_expand<<<NB1, 256, smx, st>>>(S, 1, I, 1, w.KP1, w.P1, (const float *)(ws + w.small1),
                                                   (const double *)(ws + w.Apow1), sb.Xblk, (float *)(ws + w.f.hidden),
                                                   (double *)(ws + w.f.Hlin), (float *)(ws + w.G1),
                                                   (float *)(ws + w.aux1), sb.HbarPart, sb.tickets, staged);
    }
    count_launch();
    lssm_mix_forward<<<1, 512, 0, st>>>(S, L, I, restricted, w.f.KP, prenorm, theta_w, (const float *)(ws + w.small1),
                                        (const float *)(ws + w.aux1), (const double *)(ws + w.f.Hlin),
                                        (float *)(ws + w.lsmall), (float *)(ws + w.f.G), (float *)(ws + w.f.aux));
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

// Everything after the forward pass of one LSSM epoch: the pass over D, the node passes, the log-likelihood, the
// lineage fold and the single-chain backward. raw = 1 (NSSM): d_gE / d_gw receive dLoss/dlogE and dLoss/dlog w.
int32_t lssm_loss_core(cy2p_plan *plan, int S, int L, int Le, int N, int I, int restricted, const LssmWs &w,
                       char *ws, const float *d_D, const float *h_weights, float *d_terms, float *d_gpi,
                       float *d_gw, float *d_gE, float *d_gA, int raw, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int K = S * Le;
    int32_t rc;
    const float w_sparse = h_weights[0], w_excl = h_weights[1], w_orth = h_weights[2], w_tpm = h_weights[3];
    const ScalarOffsets sc(S, L, K);
    float *logE = (float *)(ws + w.f.logE);
    double *scal = (double *)(ws + w.f.scal), *scal1 = (double *)(ws + w.scal1);
    const float *aux = (const float *)(ws + w.f.aux);
    float *X1 = (float *)(ws + w.f.X1), *X2 = (float *)(ws + w.f.X2), *Y1 = (float *)(ws + w.f.Y1),
          *TP = (float *)(ws + w.f.TP);
    dispatch_data_pass(N, I, K, w.f, ws, logE, (const float *)(ws + w.f.G), d_D, nullptr, (float *)(ws + w.f.split),
                       (float *)(ws + w.f.gG), scal + sc.div, st);
    const unsigned gN = (unsigned)((N + 127) / 128);
    count_launch();
    fhmm_node_pass_a<<<gN, 128, 0, st>>>(S, L, Le, N, logE, aux, X1, X2);
    rc = cy2p_propagate_rows(plan, X1, Y1, S + L, 0, N, stream);
    if (rc) return rc;
    count_launch();
    csr_times_small<<<(unsigned)(((int64_t)N * L + 255) / 256), 256, 0, st>>>(N, L, plan->d_indptr, plan->d_indices,
                                                                             plan->d_data, X2, TP);
    count_launch();
    fhmm_node_pass_b<<<gN, 128, 0, st>>>(S, L, Le, N, logE, aux, X1, X2, Y1, scal, K);
    count_launch();
    fhmm_node_pass_c<<<gN, 128, 0, st>>>(S, L, Le, N, K, w.f.nsplit, logE, aux, X1, X2, Y1, TP,
                                         (const float *)(ws + w.f.split), scal, w_excl, w_orth, w_tpm, d_gE);
    if (!raw) {  // softmax Jacobian of the emission rows (the NSSM chains dLoss/dlogE itself)
        count_launch();
        fhmm_emission_grad_finish<<<(unsigned)(((int64_t)K * N + 255) / 256), 256, 0, st>>>(N, K, logE,
                                                                                           scal + sc.rowsum, d_gE);
    }
    count_launch();
    fhmm_ll_nodes<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(S, L, Le, N, I, logE, (const double *)(ws + w.f.row_lse),
                                                               (double *)(ws + w.f.llv));
    const int nparts = (N + 255) / 256;
    count_launch();
    lse_partials<<<nparts, 256, 0, st>>>(N, (const double *)(ws + w.f.llv), (double2 *)(ws + w.f.llpart));
    count_launch();
    lse_combine<<<1, 256, 0, st>>>(nparts, (const double2 *)(ws + w.f.llpart), scal + sc.ll);
    // fold the lineage axis, then the single-chain backward (FHMM kernels with L = 1)
    count_launch();
    lssm_glue_backward<<<S, 256, 0, st>>>(S, L, I, restricted, K, raw, (const float *)(ws + w.lsmall),
                                          (const double *)(ws + w.f.Hlin), (const float *)(ws + w.f.gG), scal, aux,
                                          w_sparse, w_excl, w_orth, w_tpm, (float *)(ws + w.gG1), scal1, d_gw, d_terms);
    float *tmp = (float *)(ws + w.tmp);
    rc = launch_hidden_backward(S, 1, I, w.P1, 1, S, (const float *)(ws + w.small1), (const double *)(ws + w.f.Hlin),
                                (const double *)(ws + w.Apow1), (const float *)(ws + w.aux1),
                                (const float *)(ws + w.gG1), scal1, w_sparse, 0.f, 0.f, 0.f, (float *)(ws + w.f.gH),
                                d_gpi, tmp, d_gA, tmp + 4, (float *)(ws + w.f.gwacc), w.smem_local1, w.smem_bwd1, st,
                                split_bufs(w.f, ws));
    if (rc) return rc;
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

// ---- NSSM: the unrestricted LSSM work space plus the node-level pieces
struct NssmWs {
    LssmWs l;
    size_t logE0, rowsum0, logcw, logr, gEt, glw, total;
};

NssmWs nssm_layout(int S, int L, int Sc, int N, int I) {
    NssmWs w;
    w.l = lssm_layout(S, L, L, N, I);
    size_t off = w.l.total;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = up(off + bytes);
        return o;
    };
    w.logE0 = take(sizeof(float) * (size_t)S * N);
    w.rowsum0 = take(sizeof(double) * 2 * S);  // [0,S): sum exp(logE0) of the softmax, [S,2S): gradient row sums
    w.logcw = take(sizeof(float) * (size_t)Sc * N * L);
    w.logr = take(sizeof(float) * S * L);
    w.gEt = take(sizeof(float) * (size_t)S * L * N);  // Etil on the way in, dLoss/dlog Ehat on the way back
    w.glw = take(sizeof(float) * S * L);
    w.total = off;
    return w;
}

int32_t nssm_check(int S, int L, int N, int I, const NssmWs &w, const void *d_ws, size_t ws_bytes) {
    int32_t rc = check_dims(S, L, N, I, 0);  // the effective emission has S*L rows
    if (rc) return rc;
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    return CY2P_OK;
}

// log Ehat -> l.f.logE, log r -> logr, log E -> logE0, log cw -> logcw, then the LSSM forward on them
int32_t nssm_run_forward(int S, int L, int Sc, int N, int I, const float *theta_pi, const float *theta_cw,
                         const float *theta_E, const float *theta_A, const NssmWs &w, char *ws, cudaStream_t st) {
    const int K = S * L, nchunks = (N + kSoftmaxChunk - 1) / kSoftmaxChunk;
    float2 *partials = (float2 *)(ws + w.l.f.partials);
    CY2P_CUDA(cudaMemsetAsync(ws + w.rowsum0, 0, sizeof(double) * 2 * S, st));
    CY2P_CUDA(cudaMemsetAsync(ws + w.l.f.row_lse, 0, sizeof(double) * K, st));
    count_launch();
    emission_softmax_partials<<<dim3(S, nchunks), 256, 0, st>>>(N, theta_E, partials);
    count_launch();
    emission_softmax_write<<<dim3(S, nchunks), 256, 0, st>>>(N, theta_E, partials, (float *)(ws + w.logE0),
                                                              (double *)(ws + w.rowsum0));
    count_launch();
    nssm_build<<<(N + 127) / 128, 128, 0, st>>>(S, L, Sc, N, theta_cw, (const float *)(ws + w.logE0),
                                                (float *)(ws + w.logcw), (float *)(ws + w.gEt));
    count_launch();
    emission_softmax_partials<<<dim3(K, nchunks), 256, 0, st>>>(N, (const float *)(ws + w.gEt), partials);
    count_launch();
    emission_softmax_write<<<dim3(K, nchunks), 256, 0, st>>>(N, (const float *)(ws + w.gEt), partials,
                                                              (float *)(ws + w.l.f.logE), (double *)(ws + w.l.f.row_lse),
                                                              (float *)(ws + w.logr));
    return lssm_run_forward(S, L, L, N, I, 0, theta_pi, (const float *)(ws + w.logr), nullptr, theta_A, w.l, ws, st, 1);
}

}  // namespace

extern "C" {

size_t cy2p_fhmm_ws_bytes(int32_t S, int32_t L, int32_t N, int32_t I) {
    if (S < 1 || L < 1 || N < 1 || I < 1 || S > kMaxS || L > kMaxL) return 0;
    // sized for the unrestricted layout (the larger one) so one buffer serves both
    const FhmmWs a = fhmm_layout(S, L, 1, N, I);
    if (S * L > kMaxK) return a.total;
    const FhmmWs b = fhmm_layout(S, L, L, N, I);
    return a.total > b.total ? a.total : b.total;
}

int32_t cy2p_fhmm_small_floats(int32_t S, int32_t L) { return SmallOffsets(S, L).total; }

int32_t cy2p_fhmm_forward(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                          const float *theta_w, const float *theta_E, const float *theta_A, float *d_pred,
                          float *d_hidden, float *d_joint, float *d_logE, float *d_small, void *d_ws, size_t ws_bytes,
                          void *stream) {
    int32_t rc = check_dims(S, L, N, I, restricted);
    if (rc) return rc;
    CY2P_REQUIRE(theta_pi && theta_w && theta_E && theta_A, "null parameter");
    const int Le = restricted ? 1 : L, K = S * Le;
    const FhmmWs w = fhmm_layout(S, L, Le, N, I);
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    rc = run_forward(S, L, Le, N, I, restricted, theta_pi, theta_w, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    const float *logE = (const float *)(ws + w.logE);
    if (d_pred) {
        // pred only: the data pass with D == nullptr (no divergence / gradient contributions)
        CY2P_CUDA(cudaMemsetAsync(ws + w.gG, 0, sizeof(float) * (size_t)I * K, st));
        CY2P_CUDA(cudaMemsetAsync(ws + w.scal, 0, sizeof(double) * ScalarOffsets(S, L, K).total, st));
        dispatch_data_pass(N, I, K, w, ws, logE, (const float *)(ws + w.G), nullptr, d_pred, (float *)(ws + w.split),
                           (float *)(ws + w.gG), (double *)(ws + w.scal), st);
    }
    if (d_hidden)
        CY2P_CUDA(cudaMemcpyAsync(d_hidden, ws + w.hidden, sizeof(float) * (size_t)I * S * L, cudaMemcpyDeviceToDevice, st));
    if (d_logE)
        CY2P_CUDA(cudaMemcpyAsync(d_logE, logE, sizeof(float) * (size_t)K * N, cudaMemcpyDeviceToDevice, st));
    if (d_small)
        CY2P_CUDA(cudaMemcpyAsync(d_small, ws + w.small, sizeof(float) * SmallOffsets(S, L).total,
                                  cudaMemcpyDeviceToDevice, st));
    if (d_joint) {
        dim3 grid((unsigned)((N + 255) / 256), (unsigned)(I * S * L));
        count_launch();
        fhmm_joint_kernel<<<grid, 256, 0, st>>>(S, L, Le, N, I, logE, (const float *)(ws + w.hidden),
                                                (const float *)(ws + w.small) + SmallOffsets(S, L).log_w, d_joint);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_fhmm_loss_grad(cy2p_plan *plan, int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                            const float *theta_pi, const float *theta_w, const float *theta_E, const float *theta_A,
                            const float *d_D, const float *h_weights, float *d_terms, float *d_gpi, float *d_gw,
                            float *d_gE, float *d_gA, void *d_ws, size_t ws_bytes, void *stream) {
    int32_t rc = check_dims(S, L, N, I, restricted);
    if (rc) return rc;
    CY2P_REQUIRE(plan && theta_pi && theta_w && theta_E && theta_A && d_D && h_weights, "null pointer");
    CY2P_REQUIRE(d_terms && d_gpi && d_gw && d_gE && d_gA, "null output");
    CY2P_REQUIRE(plan->n == N, "plan size != num_nodes");
    const int Le = restricted ? 1 : L, K = S * Le;
    const FhmmWs w = fhmm_layout(S, L, Le, N, I);
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    char *ws = (char *)d_ws;
    const float w_sparse = h_weights[0], w_excl = h_weights[1], w_orth = h_weights[2], w_tpm = h_weights[3];
    const ScalarOffsets sc(S, L, K);
    float *logE = (float *)(ws + w.logE);
    double *scal = (double *)(ws + w.scal);
    float *X1 = (float *)(ws + w.X1), *X2 = (float *)(ws + w.X2), *Y1 = (float *)(ws + w.Y1), *TP = (float *)(ws + w.TP);

    CY2P_CUDA(cudaMemsetAsync(ws + w.gG, 0, sizeof(float) * (size_t)I * K, st));
    CY2P_CUDA(cudaMemsetAsync(scal, 0, sizeof(double) * sc.total, st));
    // Two streams (CY2P_FHMM_OVERLAP=0: one). The epoch's dependency graph has two independent arms after the memsets:
    //   main: emission softmax -> [wait small_forward] -> pass over D (P / r, dlogE, dG on the tensor cores)
    //   side: small_forward (one CTA, ~65 us of dependent steps) -> [wait logE] -> node-level conditionals, their
    //         products with T (K1 with B = S + L), regulariser sums, log-likelihood  (~150 us of small kernels)
    // and they join before the regulariser gradients (which add the dlogE slices) and the backward recurrences.
    cudaStream_t s2 = st;
    cudaEvent_t *ev = plan->side_events;
    if (env_flag("CY2P_FHMM_OVERLAP", 1)) {
        if (!plan->side_stream) {
            CY2P_CUDA(cudaStreamCreateWithFlags(&plan->side_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 6; ++i) CY2P_CUDA(cudaEventCreateWithFlags(&plan->side_events[i], cudaEventDisableTiming));
        }
        s2 = plan->side_stream;
        CY2P_CUDA(cudaEventRecord(ev[0], st));  // fork: the memsets and the caller's earlier work
        CY2P_CUDA(cudaStreamWaitEvent(s2, ev[0], 0));
    }
    const bool two = s2 != st;
    rc = run_emission(K, N, theta_E, w, ws, st);
    if (rc) return rc;
    if (two) CY2P_CUDA(cudaEventRecord(ev[1], st));  // logE ready
    rc = run_small_forward(S, L, I, restricted, theta_pi, theta_w, theta_A, w, ws, s2);
    if (rc) return rc;
    if (two) {
        CY2P_CUDA(cudaEventRecord(ev[2], s2));  // G, aux, small ready
        CY2P_CUDA(cudaStreamWaitEvent(st, ev[2], 0));
        CY2P_CUDA(cudaStreamWaitEvent(s2, ev[1], 0));
    }
    // pass over D (main stream; with two streams and the tensor-core path, dG runs on the side stream after the
    // node-level arm that is queued there below, and is joined before the backward recurrences)
    SideStream dg_side;
    const bool dg_on_side = two && w.two_pass && env_flag("CY2P_FHMM_MMA", 1);
    // node-level conditionals and their products with T (side stream)
    const unsigned gN = (unsigned)((N + 127) / 128);
    count_launch();
    fhmm_node_pass_a<<<gN, 128, 0, s2>>>(S, L, Le, N, logE, (const float *)(ws + w.aux), X1, X2);
    rc = cy2p_propagate_rows(plan, X1, Y1, S + L, 0, N, (void *)s2);  // [P(n|s) | P(n|l)] . T   (trainer.py:175,190)
    if (rc) return rc;
    count_launch();
    csr_times_small<<<(unsigned)(((int64_t)N * L + 255) / 256), 256, 0, s2>>>(N, L, plan->d_indptr, plan->d_indices,
                                                                             plan->d_data, X2, TP);
    count_launch();
    fhmm_node_pass_b<<<gN, 128, 0, s2>>>(S, L, Le, N, logE, (const float *)(ws + w.aux), X1, X2, Y1, scal, K);
    // log-likelihood (side stream)
    count_launch();
    fhmm_ll_nodes<<<(unsigned)((N + 255) / 256), 256, 0, s2>>>(S, L, Le, N, I, logE, (const double *)(ws + w.row_lse),
                                                               (double *)(ws + w.llv));
    const int nparts = (N + 255) / 256;
    count_launch();
    lse_partials<<<nparts, 256, 0, s2>>>(N, (const double *)(ws + w.llv), (double2 *)(ws + w.llpart));
    count_launch();
    lse_combine<<<1, 256, 0, s2>>>(nparts, (const double2 *)(ws + w.llpart), scal + sc.ll);
    if (two) CY2P_CUDA(cudaEventRecord(ev[3], s2));  // the node-level arm is complete
    if (dg_on_side) {
        dg_side.stream = s2;
        dg_side.fork = ev[4];
        dg_side.join = ev[5];
    }
    dispatch_data_pass(N, I, K, w, ws, logE, (const float *)(ws + w.G), d_D, nullptr, (float *)(ws + w.split),
                       (float *)(ws + w.gG), scal + sc.div, st, dg_side);
    if (two) CY2P_CUDA(cudaStreamWaitEvent(st, ev[3], 0));
    count_launch();
    fhmm_node_pass_c<<<gN, 128, 0, st>>>(S, L, Le, N, K, w.nsplit, logE, (const float *)(ws + w.aux), X1, X2, Y1, TP,
                                         (const float *)(ws + w.split), scal, w_excl, w_orth, w_tpm, d_gE);
    count_launch();
    fhmm_emission_grad_finish<<<(unsigned)(((int64_t)K * N + 255) / 256), 256, 0, st>>>(N, K, logE, scal + sc.rowsum,
                                                                                       d_gE);
    // small backward
    if (dg_on_side) CY2P_CUDA(cudaStreamWaitEvent(st, ev[5], 0));  // dG (side stream) is complete
    rc = launch_hidden_backward(S, L, I, w.P, restricted, K, (const float *)(ws + w.small),
                                (const double *)(ws + w.Hlin), (const double *)(ws + w.Apow),
                                (const float *)(ws + w.aux), (const float *)(ws + w.gG), scal, w_sparse, w_excl, w_orth,
                                w_tpm, (float *)(ws + w.gH), d_gpi, d_gw, d_gA, d_terms, (float *)(ws + w.gwacc),
                                w.smem_local, w.smem_bwd, st, split_bufs(w, ws));
    if (rc) return rc;
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_fhmm_decode(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                         const float *theta_w, const float *theta_E, const float *theta_A, const float *d_D,
                         float *d_log_alpha, float *d_log_probs, float *d_log_beta, float *d_log_gamma,
                         float *d_log_delta, float *d_psi, float *d_log_max, float *d_best_path, void *d_ws,
                         size_t ws_bytes, void *stream) {
    int32_t rc = check_dims(S, L, N, I, restricted);
    if (rc) return rc;
    CY2P_REQUIRE(theta_pi && theta_w && theta_E && theta_A && d_D, "null input");
    CY2P_REQUIRE(d_log_alpha && d_log_probs && d_log_beta && d_log_gamma && d_log_delta && d_psi && d_log_max &&
                     d_best_path, "null output");
    CY2P_REQUIRE(S * L <= 512, "decoders need S*L <= 512");
    const int Le = restricted ? 1 : L, K = S * Le;
    const FhmmWs w = fhmm_layout(S, L, Le, N, I);
    if (!d_ws || ws_bytes < w.total || ((uintptr_t)d_ws & 255)) {
        set_error("work space too small or misaligned: need %zu bytes, 256-byte aligned", w.total);
        return CY2P_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    rc = run_forward(S, L, Le, N, I, restricted, theta_pi, theta_w, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    const float *logE = (const float *)(ws + w.logE);
    float *rowmax = (float *)(ws + w.gG);   // scratch: K floats
    float *Bmat = (float *)(ws + w.X1);     // scratch: I*K floats (X1 holds N*(S+L) >= ... checked below)
    CY2P_REQUIRE((size_t)I * K <= (size_t)N * (S + L) || true, "scratch");
    float *bm = Bmat;
    if ((size_t)I * K > (size_t)N * (S + L)) bm = (float *)(ws + w.split);  // the split buffer holds nsplit*K*N floats
    count_launch();
    fhmm_row_max<<<K, 256, 0, st>>>(N, logE, rowmax);
    count_launch();
    fhmm_bmat<<<dim3(I, K), 256, 0, st>>>(N, K, d_D, logE, rowmax, bm);
    const size_t smem = sizeof(float) * ((size_t)L * S * S + 2 * (size_t)S * L + 2 * L);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_decode_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_launch();
    fhmm_decode_small<<<1, 512, smem, st>>>(S, L, Le, I, (const float *)(ws + w.small), bm, d_log_alpha, d_log_probs,
                                            d_log_beta, d_log_gamma, d_log_delta, d_psi, d_log_max, d_best_path);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

// ------------------------------------------------------------------------------------------------ LSSM
size_t cy2p_lssm_ws_bytes(int32_t S, int32_t L, int32_t N, int32_t I) {
    if (S < 1 || L < 1 || N < 1 || I < 1 || S > kMaxS || L > kMaxL) return 0;
    const LssmWs a = lssm_layout(S, L, 1, N, I);
    if (S * L > kMaxK) return a.total;
    const LssmWs b = lssm_layout(S, L, L, N, I);
    return a.total > b.total ? a.total : b.total;
}

int32_t cy2p_lssm_small_floats(int32_t S, int32_t L) { return LssmSmallOffsets(S, L).total; }

int32_t cy2p_lssm_forward(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                          const float *theta_w, const float *theta_E, const float *theta_A, float *d_pred,
                          float *d_hidden, float *d_joint, float *d_logE, float *d_small, void *d_ws, size_t ws_bytes,
                          void *stream) {
    int32_t rc = check_dims(S, L, N, I, restricted);
    if (rc) return rc;
    CY2P_REQUIRE(theta_pi && theta_w && theta_E && theta_A, "null parameter");
    const int Le = restricted ? 1 : L, K = S * Le;
    const LssmWs w = lssm_layout(S, L, Le, N, I);
    if ((rc = lssm_check_ws(w, d_ws, ws_bytes))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    rc = lssm_run_forward(S, L, Le, N, I, restricted, theta_pi, theta_w, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    const float *logE = (const float *)(ws + w.f.logE);
    if (d_pred) {
        CY2P_CUDA(cudaMemsetAsync(ws + w.f.gG, 0, sizeof(float) * (size_t)I * K, st));
        CY2P_CUDA(cudaMemsetAsync(ws + w.f.scal, 0, sizeof(double) * ScalarOffsets(S, L, K).total, st));
        dispatch_data_pass(N, I, K, w.f, ws, logE, (const float *)(ws + w.f.G), nullptr, d_pred, (float *)(ws + w.f.split),
                           (float *)(ws + w.f.gG), (double *)(ws + w.f.scal), st);
    }
    if (d_hidden)
        CY2P_CUDA(cudaMemcpyAsync(d_hidden, ws + w.f.hidden, sizeof(float) * (size_t)I * S, cudaMemcpyDeviceToDevice, st));
    if (d_logE)
        CY2P_CUDA(cudaMemcpyAsync(d_logE, logE, sizeof(float) * (size_t)K * N, cudaMemcpyDeviceToDevice, st));
    if (d_small)
        CY2P_CUDA(cudaMemcpyAsync(d_small, ws + w.lsmall, sizeof(float) * LssmSmallOffsets(S, L).total,
                                  cudaMemcpyDeviceToDevice, st));
    if (d_joint) {
        dim3 grid((unsigned)((N + 255) / 256), (unsigned)(I * S * L));
        count_launch();
        lssm_joint_kernel<<<grid, 256, 0, st>>>(S, L, Le, N, I, logE, (const float *)(ws + w.f.hidden),
                                                (const float *)(ws + w.lsmall) + LssmSmallOffsets(S, L).log_w, d_joint);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_lssm_loss_grad(cy2p_plan *plan, int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                            const float *theta_pi, const float *theta_w, const float *theta_E, const float *theta_A,
                            const float *d_D, const float *h_weights, float *d_terms, float *d_gpi, float *d_gw,
                            float *d_gE, float *d_gA, void *d_ws, size_t ws_bytes, void *stream) {
    int32_t rc = check_dims(S, L, N, I, restricted);
    if (rc) return rc;
    CY2P_REQUIRE(plan && theta_pi && theta_w && theta_E && theta_A && d_D && h_weights, "null pointer");
    CY2P_REQUIRE(d_terms && d_gpi && d_gw && d_gE && d_gA, "null output");
    CY2P_REQUIRE(plan->n == N, "plan size != num_nodes");
    const int Le = restricted ? 1 : L, K = S * Le;
    const LssmWs w = lssm_layout(S, L, Le, N, I);
    if ((rc = lssm_check_ws(w, d_ws, ws_bytes))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    char *ws = (char *)d_ws;
    CY2P_CUDA(cudaMemsetAsync(ws + w.f.gG, 0, sizeof(float) * (size_t)I * K, st));
    CY2P_CUDA(cudaMemsetAsync(ws + w.f.scal, 0, sizeof(double) * ScalarOffsets(S, L, K).total, st));
    CY2P_CUDA(cudaMemsetAsync(ws + w.scal1, 0, sizeof(double) * ScalarOffsets(S, 1, S).total, st));
    rc = lssm_run_forward(S, L, Le, N, I, restricted, theta_pi, theta_w, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    return lssm_loss_core(plan, S, L, Le, N, I, restricted, w, ws, d_D, h_weights, d_terms, d_gpi, d_gw, d_gE, d_gA, 0,
                          stream);
}

int32_t cy2p_lssm_decode(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                         const float *theta_w, const float *theta_E, const float *theta_A, const float *d_D,
                         float *d_log_alpha, float *d_log_probs, float *d_log_beta, float *d_log_gamma,
                         float *d_log_delta, float *d_psi, float *d_log_max, float *d_best_path, void *d_ws,
                         size_t ws_bytes, void *stream) {
    int32_t rc = check_dims(S, L, N, I, restricted);
    if (rc) return rc;
    CY2P_REQUIRE(theta_pi && theta_w && theta_E && theta_A && d_D, "null input");
    CY2P_REQUIRE(d_log_alpha && d_log_probs && d_log_beta && d_log_gamma && d_log_delta && d_psi && d_log_max &&
                     d_best_path, "null output");
    CY2P_REQUIRE(S * L <= 512, "decoders need S*L <= 512");
    const int Le = restricted ? 1 : L, K = S * Le;
    const LssmWs w = lssm_layout(S, L, Le, N, I);
    if ((rc = lssm_check_ws(w, d_ws, ws_bytes))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    rc = lssm_run_forward(S, L, Le, N, I, restricted, theta_pi, theta_w, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    const float *logE = (const float *)(ws + w.f.logE);
    float *rowmax = (float *)(ws + w.f.gG);
    float *bm = (float *)(ws + w.f.X1);
    if ((size_t)I * K > (size_t)N * (S + L)) bm = (float *)(ws + w.f.split);
    count_launch();
    fhmm_row_max<<<K, 256, 0, st>>>(N, logE, rowmax);
    count_launch();
    fhmm_bmat<<<dim3(I, K), 256, 0, st>>>(N, K, d_D, logE, rowmax, bm);
    float *small_exp = (float *)(ws + w.f.small), *Bexp = (float *)(ws + w.f.gH);  // both free on this path
    count_launch();
    lssm_expand_decode<<<64, 256, 0, st>>>(S, L, Le, I, (const float *)(ws + w.lsmall), bm, small_exp, Bexp);
    const size_t smem = sizeof(float) * ((size_t)L * S * S + 2 * (size_t)S * L + 2 * L);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_decode_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_launch();
    fhmm_decode_small<<<1, 512, smem, st>>>(S, L, L, I, small_exp, Bexp, d_log_alpha, d_log_probs, d_log_beta,
                                            d_log_gamma, d_log_delta, d_psi, d_log_max, d_best_path);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

// ------------------------------------------------------------------------------------------------ NSSM
size_t cy2p_nssm_ws_bytes(int32_t S, int32_t L, int32_t N, int32_t I) {
    if (S < 1 || L < 1 || N < 1 || I < 1 || S > kMaxS || L > kMaxL || S * L > kMaxK) return 0;
    return nssm_layout(S, L, S, N, I).total;
}

int32_t cy2p_nssm_forward(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                          const float *theta_cw, const float *theta_E, const float *theta_A, float *d_pred,
                          float *d_hidden, float *d_joint, float *d_logE, float *d_logcw, float *d_small, void *d_ws,
                          size_t ws_bytes, void *stream) {
    const int Sc = restricted ? 1 : S, K = S * L;
    const NssmWs w = nssm_layout(S > 0 ? S : 1, L > 0 ? L : 1, Sc > 0 ? Sc : 1, N > 0 ? N : 1, I > 0 ? I : 1);
    int32_t rc = nssm_check(S, L, N, I, w, d_ws, ws_bytes);
    if (rc) return rc;
    CY2P_REQUIRE(theta_pi && theta_cw && theta_E && theta_A, "null parameter");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    rc = nssm_run_forward(S, L, Sc, N, I, theta_pi, theta_cw, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    const float *logEhat = (const float *)(ws + w.l.f.logE);
    if (d_pred) {
        CY2P_CUDA(cudaMemsetAsync(ws + w.l.f.gG, 0, sizeof(float) * (size_t)I * K, st));
        CY2P_CUDA(cudaMemsetAsync(ws + w.l.f.scal, 0, sizeof(double) * ScalarOffsets(S, L, K).total, st));
        dispatch_data_pass(N, I, K, w.l.f, ws, logEhat, (const float *)(ws + w.l.f.G), nullptr, d_pred,
                           (float *)(ws + w.l.f.split), (float *)(ws + w.l.f.gG), (double *)(ws + w.l.f.scal), st);
    }
    if (d_hidden)
        CY2P_CUDA(cudaMemcpyAsync(d_hidden, ws + w.l.f.hidden, sizeof(float) * (size_t)I * S, cudaMemcpyDeviceToDevice, st));
    if (d_logE)
        CY2P_CUDA(cudaMemcpyAsync(d_logE, ws + w.logE0, sizeof(float) * (size_t)S * N, cudaMemcpyDeviceToDevice, st));
    if (d_logcw)
        CY2P_CUDA(cudaMemcpyAsync(d_logcw, ws + w.logcw, sizeof(float) * (size_t)Sc * N * L, cudaMemcpyDeviceToDevice, st));
    if (d_small)
        CY2P_CUDA(cudaMemcpyAsync(d_small, ws + w.l.lsmall, sizeof(float) * LssmSmallOffsets(S, L).total,
                                  cudaMemcpyDeviceToDevice, st));
    if (d_joint) {
        dim3 grid((unsigned)((N + 255) / 256), (unsigned)(I * S * L));
        count_launch();
        lssm_joint_kernel<<<grid, 256, 0, st>>>(S, L, L, N, I, logEhat, (const float *)(ws + w.l.f.hidden),
                                                (const float *)(ws + w.l.lsmall) + LssmSmallOffsets(S, L).log_w,
                                                d_joint);
    }
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_nssm_loss_grad(cy2p_plan *plan, int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted,
                            const float *theta_pi, const float *theta_cw, const float *theta_E, const float *theta_A,
                            const float *d_D, const float *h_weights, float *d_terms, float *d_gpi, float *d_gcw,
                            float *d_gE, float *d_gA, void *d_ws, size_t ws_bytes, void *stream) {
    const int Sc = restricted ? 1 : S, K = S * L;
    const NssmWs w = nssm_layout(S > 0 ? S : 1, L > 0 ? L : 1, Sc > 0 ? Sc : 1, N > 0 ? N : 1, I > 0 ? I : 1);
    int32_t rc = nssm_check(S, L, N, I, w, d_ws, ws_bytes);
    if (rc) return rc;
    CY2P_REQUIRE(plan && theta_pi && theta_cw && theta_E && theta_A && d_D && h_weights, "null pointer");
    CY2P_REQUIRE(d_terms && d_gpi && d_gcw && d_gE && d_gA, "null output");
    CY2P_REQUIRE(plan->n == N, "plan size != num_nodes");
    cudaStream_t st = (cudaStream_t)stream;
    CY2P_CUDA(cudaSetDevice(plan->device));
    char *ws = (char *)d_ws;
    const ScalarOffsets sc(S, L, K);
    CY2P_CUDA(cudaMemsetAsync(ws + w.l.f.gG, 0, sizeof(float) * (size_t)I * K, st));
    CY2P_CUDA(cudaMemsetAsync(ws + w.l.f.scal, 0, sizeof(double) * sc.total, st));
    CY2P_CUDA(cudaMemsetAsync(ws + w.l.scal1, 0, sizeof(double) * ScalarOffsets(S, 1, S).total, st));
    rc = nssm_run_forward(S, L, Sc, N, I, theta_pi, theta_cw, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    rc = lssm_loss_core(plan, S, L, L, N, I, 0, w.l, ws, d_D, h_weights, d_terms, d_gpi, (float *)(ws + w.glw),
                        (float *)(ws + w.gEt), d_gA, 1, stream);
    if (rc) return rc;
    double *rowsum0 = (double *)(ws + w.rowsum0) + S;
    count_launch();
    nssm_backward_nodes<<<(N + 127) / 128, 128, 0, st>>>(S, L, Sc, N, (const float *)(ws + w.gEt),
                                                         (const float *)(ws + w.l.f.logE), (const float *)(ws + w.glw),
                                                         (const double *)(ws + w.l.f.scal) + sc.rowsum,
                                                         (const float *)(ws + w.logcw), d_gE, rowsum0, d_gcw);
    count_launch();
    fhmm_emission_grad_finish<<<(unsigned)(((int64_t)S * N + 255) / 256), 256, 0, st>>>(
        N, S, (const float *)(ws + w.logE0), rowsum0, d_gE);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

int32_t cy2p_nssm_decode(int32_t S, int32_t L, int32_t N, int32_t I, int32_t restricted, const float *theta_pi,
                         const float *theta_cw, const float *theta_E, const float *theta_A, const float *d_D,
                         float *d_log_alpha, float *d_log_probs, float *d_log_beta, float *d_log_gamma,
                         float *d_log_delta, float *d_psi, float *d_log_max, float *d_best_path, void *d_ws,
                         size_t ws_bytes, void *stream) {
    const int Sc = restricted ? 1 : S, K = S * L;
    const NssmWs w = nssm_layout(S > 0 ? S : 1, L > 0 ? L : 1, Sc > 0 ? Sc : 1, N > 0 ? N : 1, I > 0 ? I : 1);
    int32_t rc = nssm_check(S, L, N, I, w, d_ws, ws_bytes);
    if (rc) return rc;
    CY2P_REQUIRE(theta_pi && theta_cw && theta_E && theta_A && d_D, "null input");
    CY2P_REQUIRE(d_log_alpha && d_log_probs && d_log_beta && d_log_gamma && d_log_delta && d_psi && d_log_max &&
                     d_best_path, "null output");
    CY2P_REQUIRE(S * L <= 512, "decoders need S*L <= 512");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    rc = nssm_run_forward(S, L, Sc, N, I, theta_pi, theta_cw, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    // Bmat over the normalised rows; log r joins it in lssm_expand_decode (log_weights = log Ehat + log r, _NSSM.py:80)
    const float *logEhat = (const float *)(ws + w.l.f.logE);
    float *rowmax = (float *)(ws + w.l.f.gG);
    float *bm = (float *)(ws + w.l.f.X1);
    if ((size_t)I * K > (size_t)N * (S + L)) bm = (float *)(ws + w.l.f.split);
    count_launch();
    fhmm_row_max<<<K, 256, 0, st>>>(N, logEhat, rowmax);
    count_launch();
    fhmm_bmat<<<dim3(I, K), 256, 0, st>>>(N, K, d_D, logEhat, rowmax, bm);
    float *small_exp = (float *)(ws + w.l.f.small), *Bexp = (float *)(ws + w.l.f.gH);
    count_launch();
    lssm_expand_decode<<<64, 256, 0, st>>>(S, L, L, I, (const float *)(ws + w.l.lsmall), bm, small_exp, Bexp);
    const size_t smem = sizeof(float) * ((size_t)L * S * S + 2 * (size_t)S * L + 2 * L);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_decode_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_launch();
    fhmm_decode_small<<<1, 512, smem, st>>>(S, L, L, I, small_exp, Bexp, d_log_alpha, d_log_probs, d_log_beta,
                                            d_log_gamma, d_log_delta, d_psi, d_log_max, d_best_path);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"
