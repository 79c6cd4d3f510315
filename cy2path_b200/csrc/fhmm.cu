// K3 — factorial latent dynamic model (FHMM): forward pass, training loss and ANALYTIC gradients.
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

constexpr int kMaxS = 64, kMaxL = 16, kMaxK = 64;

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
// log_softmax over the N nodes of every emission row (methods.py:17-19). One CTA per row.
// row_lse[row] (double) = logsumexp of the written row (the ~1e-7 residual the reference's
// compute_log_likelihood subtracts, methods.py:33-34).
__global__ void __launch_bounds__(1024) emission_logsoftmax(int N, const float *__restrict__ theta,
                                                            float *__restrict__ logE, double *__restrict__ row_lse) {
    const float *x = theta + (size_t)blockIdx.x * N;
    float *y = logE + (size_t)blockIdx.x * N;
    __shared__ float smax[32];
    __shared__ double ssum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if (lane == 0) smax[warp] = m;
    __syncthreads();
    m = smax[0];
    for (int w = 1; w < nw; ++w) m = fmaxf(m, smax[w]);
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += (double)expf(x[i] - m);
    s = warp_sum(s);
    if (lane == 0) ssum[warp] = s;
    __syncthreads();
    s = 0.0;
    for (int w = 0; w < nw; ++w) s += ssum[w];
    const float lse = m + (float)log(s);
    double r = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float v = x[i] - lse;
        y[i] = v;
        r += exp((double)v);
    }
    __syncthreads();
    r = warp_sum(r);
    if (lane == 0) ssum[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        r = 0.0;
        for (int w = 0; w < nw; ++w) r += ssum[w];
        row_lse[blockIdx.x] = log(r);
    }
}

// --------------------------------------------------------------------------------------------------
// Small forward (one CTA): the three small log_softmaxes, the mixed TPM, the log-domain hidden
// recurrence exactly as the reference writes it (lse over the source state), G and C.
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

__global__ void __launch_bounds__(256) fhmm_small_forward(int S, int L, int I, int restricted, int KP,
                                                          const float *__restrict__ theta_pi,
                                                          const float *__restrict__ theta_w,
                                                          const float *__restrict__ theta_A, float *__restrict__ small,
                                                          float *__restrict__ hidden, float *__restrict__ G,
                                                          float *__restrict__ aux) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    float *log_pi = small + so.log_pi, *log_w = small + so.log_w, *log_A = small + so.log_A,
          *log_Amix = small + so.log_Amix;
    extern __shared__ float sm[];
    float *s_logA = sm;                   // [L][S][S]
    float *s_h = s_logA + L * S * S;      // [2][S][L]
    float *s_logw = s_h + 2 * S * L;      // [L]
    const int tid = threadIdx.x, nt = blockDim.x;

    // log_softmax(theta_pi, dim=0): per lineage l over states
    for (int l = tid; l < L; l += nt) {
        float m = -INFINITY;
        for (int s = 0; s < S; ++s) m = fmaxf(m, theta_pi[s * L + l]);
        float z = 0.f;
        for (int s = 0; s < S; ++s) z += expf(theta_pi[s * L + l] - m);
        const float lse = m + logf(z);
        for (int s = 0; s < S; ++s) log_pi[s * L + l] = theta_pi[s * L + l] - lse;
    }
    if (tid == 0) {
        float m = -INFINITY;
        for (int l = 0; l < L; ++l) m = fmaxf(m, theta_w[l]);
        float z = 0.f;
        for (int l = 0; l < L; ++l) z += expf(theta_w[l] - m);
        const float lse = m + logf(z);
        for (int l = 0; l < L; ++l) {
            log_w[l] = theta_w[l] - lse;
            s_logw[l] = theta_w[l] - lse;
        }
    }
    for (int r = tid; r < L * S; r += nt) {  // rows (l, s) over destination state
        const float *x = theta_A + (size_t)r * S;
        float m = -INFINITY;
        for (int j = 0; j < S; ++j) m = fmaxf(m, x[j]);
        float z = 0.f;
        for (int j = 0; j < S; ++j) z += expf(x[j] - m);
        const float lse = m + logf(z);
        for (int j = 0; j < S; ++j) {
            log_A[(size_t)r * S + j] = x[j] - lse;
            s_logA[r * S + j] = x[j] - lse;
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
    }
    __syncthreads();
    if (tid == 0) {
        float d = 0.f;
        for (int s = 0; s < S; ++s) d += log_Amix[s * S + s];
        aux[ao.sparsity] = -d / (float)S;
    }

    // hidden recurrence. Thread (s', l) owns one entry; S*L <= blockDim.
    const int SL = S * L;
    const bool own = tid < SL;
    const int sp = own ? tid / L : 0, l = own ? tid % L : 0;
    float hsum = 0.f;
    if (own) {
        const float h0 = log_pi[tid];
        s_h[tid] = h0;
        hidden[tid] = h0;
        hsum = expf(h0);
    }
    __syncthreads();
    for (int t = 0; t < I; ++t) {
        const float *cur = s_h + (t & 1) * SL;
        float *nxt = s_h + ((t + 1) & 1) * SL;
        // G[t][k] from the current hidden state
        if (restricted) {
            if (tid < S) {
                float g = 0.f;
                for (int ll = 0; ll < L; ++ll) g += expf(cur[tid * L + ll] + s_logw[ll]);
                G[(size_t)t * KP + tid] = g;
            } else if (tid < KP) {
                G[(size_t)t * KP + tid] = 0.f;
            }
        } else {
            if (own) G[(size_t)t * KP + tid] = expf(cur[tid] + s_logw[l]);
            else if (tid < KP) G[(size_t)t * KP + tid] = 0.f;
        }
        if (t + 1 < I && own) {
            // hidden[t+1][s', l] = lse_s(hidden[t][s, l] + log A[l, s, s'])   (_FHMM.py:116-124)
            float m = -INFINITY;
            for (int s = 0; s < S; ++s) m = fmaxf(m, cur[s * L + l] + s_logA[(l * S + s) * S + sp]);
            float z = 0.f;
            for (int s = 0; s < S; ++s) z += expf(cur[s * L + l] + s_logA[(l * S + s) * S + sp] - m);
            const float h = m + logf(z);
            nxt[tid] = h;
            hidden[(size_t)(t + 1) * SL + tid] = h;
            hsum += expf(h);
        }
        __syncthreads();
    }
    // time averages: Hbar, log C = log w + log Hbar, log c_s = log sum_l C, log q_l = log sum_s C
    float *s_logC = s_h;  // reuse
    if (own) {
        const float hb = hsum / (float)I;
        aux[ao.Hbar + tid] = hb;
        const float lc = logf(hb) + s_logw[l];
        aux[ao.logC + tid] = lc;
        s_logC[tid] = lc;
    }
    __syncthreads();
    for (int s = tid; s < S; s += nt) {
        float m = -INFINITY;
        for (int ll = 0; ll < L; ++ll) m = fmaxf(m, s_logC[s * L + ll]);
        float z = 0.f;
        for (int ll = 0; ll < L; ++ll) z += expf(s_logC[s * L + ll] - m);
        aux[ao.logc + s] = m + logf(z);
    }
    for (int ll = tid; ll < L; ll += nt) {
        float m = -INFINITY;
        for (int s = 0; s < S; ++s) m = fmaxf(m, s_logC[s * L + ll]);
        float z = 0.f;
        for (int s = 0; s < S; ++s) z += expf(s_logC[s * L + ll] - m);
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
template <int KP, int J>
__global__ void __launch_bounds__(128)
fhmm_data_pass(int N, int I, int K, int tchunk, const float *__restrict__ logE, const float *__restrict__ G,
               const float *__restrict__ D, float *__restrict__ pred, float *__restrict__ glogE_split,
               float *__restrict__ gG, double *__restrict__ div_acc, float inv_I) {
    extern __shared__ float sG[];  // [tchunk][KP]
    const int t0 = blockIdx.y * tchunk, t1 = min(I, t0 + tchunk);
    for (int i = threadIdx.x; i < (t1 - t0) * KP; i += blockDim.x) sG[i] = G[(size_t)t0 * KP + i];
    const int lane = threadIdx.x & 31;
    const int base = blockIdx.x * blockDim.x * J + threadIdx.x;
    float Eh[J][KP], dE[J][KP], mcol[J];
    bool ok[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int n = base + j * blockDim.x;
        ok[j] = n < N;
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            Eh[j][k] = (ok[j] && k < K) ? logE[(size_t)k * N + n] : -INFINITY;
            m = fmaxf(m, Eh[j][k]);
        }
        if (!ok[j]) m = 0.f;
        mcol[j] = m;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            Eh[j][k] = (ok[j] && k < K) ? expf(Eh[j][k] - m) : 0.f;
            dE[j][k] = 0.f;
        }
    }
    __syncthreads();
    double div = 0.0;
    for (int t = t0; t < t1; ++t) {
        const float *g = sG + (t - t0) * KP;
        float part[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) part[k] = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int n = base + j * blockDim.x;
            const float d = (ok[j] && D != nullptr) ? __ldg(D + (size_t)t * N + n) : 0.f;
            float p = 0.f;
#pragma unroll
            for (int k = 0; k < KP; ++k) p = fmaf(g[k], Eh[j][k], p);
            p = fmaxf(p, FLT_MIN);
            const float lp = logf(p);
            if (pred != nullptr && ok[j]) pred[(size_t)t * N + n] = mcol[j] + lp;
            float r = 0.f;
            if (d > 0.f) {
                r = d / p;
                div += (double)(d * (logf(d) - mcol[j] - lp));  // xlogy(D,D) - D*pred  (KLDivLoss, trainer.py:139)
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                dE[j][k] = fmaf(r, g[k], dE[j][k]);
                part[k] = fmaf(r, Eh[j][k], part[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float s = warp_sum(part[k]);
            if (lane == 0 && k < K && s != 0.f) atomicAdd(&gG[(size_t)t * K + k], -inv_I * s);
        }
    }
    float *out = glogE_split + (size_t)blockIdx.y * K * N;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int n = base + j * blockDim.x;
        if (!ok[j]) continue;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) out[(size_t)k * N + n] = -inv_I * Eh[j][k] * dE[j][k];
    }
    div = warp_sum(div);
    __shared__ double sdiv[4];
    if (lane == 0) sdiv[threadIdx.x >> 5] = div;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(div_acc, (sdiv[0] + sdiv[1] + sdiv[2] + sdiv[3]) * (double)inv_I);
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
    for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p) acc = fmaf(data[p], X[(size_t)indices[p] * B + b], acc);
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
                              const double *__restrict__ row_lse, double *__restrict__ v) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double outer_m = -INFINITY, outer[kMaxL];
    for (int l = 0; l < L; ++l) {
        const int e = Le == 1 ? 0 : l;
        double m = -INFINITY;
        for (int s = 0; s < S; ++s)
            m = fmax(m, (double)I * ((double)logE[(size_t)(s * Le + e) * N + n] - row_lse[s * Le + e]));
        double z = 0.0;
        for (int s = 0; s < S; ++s)
            z += exp((double)I * ((double)logE[(size_t)(s * Le + e) * N + n] - row_lse[s * Le + e]) - m);
        outer[l] = m + log(z) - log((double)S);
        outer_m = fmax(outer_m, outer[l]);
    }
    double z = 0.0;
    for (int l = 0; l < L; ++l) z += exp(outer[l] - outer_m);
    v[n] = outer_m + log(z) - log((double)L);
}

__global__ void __launch_bounds__(1024) lse_double(int N, const double *__restrict__ v, double *__restrict__ out) {
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmax(m, v[i]);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    m = sm[0];
    for (int w = 1; w < nw; ++w) m = fmax(m, sm[w]);
    __syncthreads();
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += exp(v[i] - m);
    s = warp_sum(s);
    if (lane == 0) sm[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0;
        for (int w = 0; w < nw; ++w) s += sm[w];
        out[0] = m + log(s);
    }
}

// --------------------------------------------------------------------------------------------------
// Small backward (one CTA): dLoss/dG and dLoss/dlogC -> w, H_t; reverse recurrence -> A, pi; the sparsity term;
// softmax Jacobians; assembles the 8 reported scalars.
//   terms: loss, divergence, log_likelihood, sparsity, orthogonality, exclusivity, regularisation, reserved
__global__ void __launch_bounds__(1024)
fhmm_small_backward(int S, int L, int I, int restricted, int K, const float *__restrict__ small,
                    const float *__restrict__ hidden, const float *__restrict__ aux, const float *__restrict__ gG,
                    const double *__restrict__ scal, float w_sparse, float w_excl, float w_orth, float w_tpm,
                    float *__restrict__ g_pi, float *__restrict__ g_w, float *__restrict__ g_A,
                    float *__restrict__ terms) {
    const SmallOffsets so(S, L);
    const AuxOffsets ao(S, L);
    const ScalarOffsets sc(S, L, K);
    const float *log_pi = small + so.log_pi, *log_w = small + so.log_w, *log_A = small + so.log_A,
                *log_Amix = small + so.log_Amix;
    extern __shared__ float sm[];
    const int SL = S * L, SSL = S * S * L;
    float *s_A = sm;              // [L][S][S] probabilities
    float *s_gA = s_A + SSL;      // [L][S][S]
    float *s_gH = s_gA + SSL;     // [2][S][L]
    float *s_H = s_gH + 2 * SL;   // [S][L]  H_{t-1}
    float *s_w = s_H + SL;        // [L]
    float *s_gw = s_w + L;        // [L] d/d(prob w)
    float *s_base = s_gw + L;     // [S][L] glogC / (I * Hbar)
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < SSL; i += nt) {
        s_A[i] = expf(log_A[i]);
        s_gA[i] = 0.f;
    }
    for (int i = tid; i < L; i += nt) {
        s_w[i] = expf(log_w[i]);
        s_gw[i] = 0.f;
    }
    for (int i = tid; i < SL; i += nt) {
        const float hb = fmaxf(aux[ao.Hbar + i], FLT_MIN);
        s_base[i] = (float)scal[sc.glogC + i] / ((float)I * hb);
    }
    __syncthreads();
    // gH_t[s,l] = w_l * gG_t[s,l] + base[s,l] + sum_s' A[l,s,s'] gH_{t+1}[s',l]
    // gw_l += sum_s gG_t[s,l] H_t[s,l];  gA[l,s,s'] += H_{t-1}[s,l] gH_t[s',l]
    float gw_part = 0.f;  // thread (s,l) partial of sum_t gG*H
    for (int t = I - 1; t >= 0; --t) {
        float *cur = s_gH + (t & 1) * SL;
        const float *nxt = s_gH + ((t + 1) & 1) * SL;
        if (tid < SL) {
            const int s = tid / L, l = tid % L;
            const float gg = gG[(size_t)t * K + (restricted ? s : tid)];
            const float h = expf(hidden[(size_t)t * SL + tid]);
            gw_part = fmaf(gg, h, gw_part);
            float g = s_w[l] * gg + s_base[tid];
            if (t + 1 < I)
                for (int sp = 0; sp < S; ++sp) g = fmaf(s_A[(l * S + s) * S + sp], nxt[sp * L + l], g);
            cur[tid] = g;
            s_H[tid] = h;  // H_t, used as H_{t-1} by the gA update of step t+1 ... see below
        }
        __syncthreads();
        // gA uses H_{t}[s,l] * gH_{t+1}[s',l]: at this point s_H = H_t and nxt = gH_{t+1}
        if (t + 1 < I)
            for (int i = tid; i < SSL; i += nt) {
                const int l = i / (S * S), s = (i / S) % S, sp = i % S;
                s_gA[i] = fmaf(s_H[s * L + l], nxt[sp * L + l], s_gA[i]);
            }
        __syncthreads();
    }
    // reduce gw over s
    if (tid < SL) atomicAdd(&s_gw[tid % L], gw_part);
    __syncthreads();
    const float *gH0 = s_gH;  // t = 0 wrote buffer 0
    // theta_pi: softmax over s for each l
    for (int l = tid; l < L; l += nt) {
        float dot = 0.f;  // sum_s glogpi[s,l], glogpi = pi * gH_0
        for (int s = 0; s < S; ++s) dot = fmaf(expf(log_pi[s * L + l]), gH0[s * L + l], dot);
        for (int s = 0; s < S; ++s) {
            const float p = expf(log_pi[s * L + l]);
            g_pi[s * L + l] = p * gH0[s * L + l] - p * dot;
        }
    }
    // theta_A: glogA = A*gA + sparsity part on the diagonal; softmax over s'
    for (int r = tid; r < L * S; r += nt) {
        const int l = r / S, s = r % S;
        const float diag = -w_sparse / (float)S * s_w[l] * s_A[(size_t)r * S + s] / expf(log_Amix[s * S + s]);
        float sum = 0.f;
        for (int sp = 0; sp < S; ++sp) {
            float gl = s_A[(size_t)r * S + sp] * s_gA[(size_t)r * S + sp];
            if (sp == s) gl += diag;
            sum += gl;
        }
        for (int sp = 0; sp < S; ++sp) {
            float gl = s_A[(size_t)r * S + sp] * s_gA[(size_t)r * S + sp];
            if (sp == s) gl += diag;
            g_A[(size_t)r * S + sp] = gl - s_A[(size_t)r * S + sp] * sum;
        }
    }
    if (tid == 0) {
        // theta_w: glogw_l = w_l*gw_l + sum_s glogC[s,l] + sparsity part
        float glw[kMaxL], sum = 0.f;
        for (int l = 0; l < L; ++l) {
            float g = s_w[l] * s_gw[l];
            for (int s = 0; s < S; ++s) {
                g += (float)scal[sc.glogC + s * L + l];
                g += -w_sparse / (float)S * s_w[l] * s_A[(l * S + s) * S + s] / expf(log_Amix[s * S + s]);
            }
            glw[l] = g;
            sum += g;
        }
        for (int l = 0; l < L; ++l) g_w[l] = glw[l] - s_w[l] * sum;
        // scalar terms
        const float divergence = (float)scal[sc.div];
        const float sparsity = aux[ao.sparsity];
        float orth = 0.f;
        for (int s = 0; s < S; ++s) orth += (float)scal[sc.KL + s];
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
    size_t logE, row_lse, hidden, G, small, aux, gG, split, X1, X2, Y1, TP, scal, llv, total;
    int KP, nsplit, tchunk;
};

size_t up(size_t x) { return (x + 255) / 256 * 256; }

FhmmWs fhmm_layout(int S, int L, int Le, int N, int I) {
    FhmmWs w;
    const int K = S * Le;
    w.KP = round_kp(K);
    // t-splits: enough CTAs to fill the machine, G chunk <= 32 KB of shared memory
    const int J = w.KP <= 12 ? 4 : (w.KP <= 24 ? 2 : 1);
    const int tiles = (N + 128 * J - 1) / (128 * J);
    int nsplit = (148 * 6 + tiles - 1) / tiles;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > I) nsplit = I;
    int tchunk = (I + nsplit - 1) / nsplit;
    const int max_chunk = 8192 / (w.KP > 0 ? w.KP : 1);
    if (tchunk > max_chunk) tchunk = max_chunk;
    nsplit = (I + tchunk - 1) / tchunk;
    w.nsplit = nsplit;
    w.tchunk = tchunk;
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
    w.total = off;
    return w;
}

template <int KP, int J>
void launch_data_pass(int N, int I, int K, const FhmmWs &w, const float *logE, const float *G, const float *D,
                      float *pred, float *split, float *gG, double *div_acc, cudaStream_t st) {
    dim3 grid((unsigned)((N + 128 * J - 1) / (128 * J)), (unsigned)w.nsplit);
    const size_t smem = sizeof(float) * (size_t)w.tchunk * KP;
    count_launch();
    fhmm_data_pass<KP, J><<<grid, 128, smem, st>>>(N, I, K, w.tchunk, logE, G, D, pred, split, gG, div_acc,
                                                    1.0f / (float)I);
}

void dispatch_data_pass(int N, int I, int K, const FhmmWs &w, const float *logE, const float *G, const float *D,
                        float *pred, float *split, float *gG, double *div_acc, cudaStream_t st) {
#define CASE(KP_, J_)                                                                         \
    case KP_:                                                                                 \
        launch_data_pass<KP_, J_>(N, I, K, w, logE, G, D, pred, split, gG, div_acc, st);      \
        break;
    switch (w.KP) {
        CASE(2, 4) CASE(4, 4) CASE(6, 4) CASE(8, 4) CASE(10, 4) CASE(12, 4) CASE(16, 2) CASE(20, 2) CASE(24, 2)
        CASE(32, 1) CASE(40, 1) CASE(48, 1) CASE(64, 1)
    }
#undef CASE
}

int32_t check_dims(int S, int L, int N, int I, int restricted) {
    CY2P_REQUIRE(S >= 1 && L >= 1 && N >= 1 && I >= 1, "S, L, N, I must be >= 1");
    CY2P_REQUIRE(S <= kMaxS && L <= kMaxL && S * L <= 256, "supported sizes: S <= 64, L <= 16, S*L <= 256");
    CY2P_REQUIRE(S * (restricted ? 1 : L) <= kMaxK, "supported sizes: S*(restricted ? 1 : L) <= 64");
    return CY2P_OK;
}

// forward pieces shared by cy2p_fhmm_forward and cy2p_fhmm_loss_grad
int32_t run_forward(int S, int L, int Le, int N, int I, int restricted, const float *theta_pi, const float *theta_w,
                    const float *theta_E, const float *theta_A, const FhmmWs &w, char *ws, cudaStream_t st) {
    const int K = S * Le;
    float *logE = (float *)(ws + w.logE);
    count_launch();
    emission_logsoftmax<<<K, 1024, 0, st>>>(N, theta_E, logE, (double *)(ws + w.row_lse));
    const size_t smem = sizeof(float) * ((size_t)L * S * S + 2 * (size_t)S * L + L);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_small_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_launch();
    fhmm_small_forward<<<1, 256, smem, st>>>(S, L, I, restricted, w.KP, theta_pi, theta_w, theta_A,
                                             (float *)(ws + w.small), (float *)(ws + w.hidden), (float *)(ws + w.G),
                                             (float *)(ws + w.aux));
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
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
        dispatch_data_pass(N, I, K, w, logE, (const float *)(ws + w.G), nullptr, d_pred, (float *)(ws + w.split),
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
    rc = run_forward(S, L, Le, N, I, restricted, theta_pi, theta_w, theta_E, theta_A, w, ws, st);
    if (rc) return rc;
    // pass over D
    dispatch_data_pass(N, I, K, w, logE, (const float *)(ws + w.G), d_D, nullptr, (float *)(ws + w.split),
                       (float *)(ws + w.gG), scal + sc.div, st);
    // node-level conditionals and their products with T
    const unsigned gN = (unsigned)((N + 127) / 128);
    count_launch();
    fhmm_node_pass_a<<<gN, 128, 0, st>>>(S, L, Le, N, logE, (const float *)(ws + w.aux), X1, X2);
    rc = cy2p_propagate_rows(plan, X1, Y1, S + L, 0, N, stream);  // [P(n|s) | P(n|l)] . T   (trainer.py:175,190)
    if (rc) return rc;
    count_launch();
    csr_times_small<<<(unsigned)(((int64_t)N * L + 255) / 256), 256, 0, st>>>(N, L, plan->d_indptr, plan->d_indices,
                                                                             plan->d_data, X2, TP);
    count_launch();
    fhmm_node_pass_b<<<gN, 128, 0, st>>>(S, L, Le, N, logE, (const float *)(ws + w.aux), X1, X2, Y1, scal, K);
    count_launch();
    fhmm_node_pass_c<<<gN, 128, 0, st>>>(S, L, Le, N, K, w.nsplit, logE, (const float *)(ws + w.aux), X1, X2, Y1, TP,
                                         (const float *)(ws + w.split), scal, w_excl, w_orth, w_tpm, d_gE);
    count_launch();
    fhmm_emission_grad_finish<<<(unsigned)(((int64_t)K * N + 255) / 256), 256, 0, st>>>(N, K, logE, scal + sc.rowsum,
                                                                                       d_gE);
    // log-likelihood
    count_launch();
    fhmm_ll_nodes<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(S, L, Le, N, I, logE, (const double *)(ws + w.row_lse),
                                                               (double *)(ws + w.llv));
    count_launch();
    lse_double<<<1, 1024, 0, st>>>(N, (const double *)(ws + w.llv), scal + sc.ll);
    // small backward
    const size_t smem = sizeof(float) * (2 * (size_t)S * S * L + 4 * (size_t)S * L + 2 * L);
    CY2P_CUDA(cudaFuncSetAttribute(fhmm_small_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_launch();
    fhmm_small_backward<<<1, 1024, smem, st>>>(S, L, I, restricted, K, (const float *)(ws + w.small),
                                               (const float *)(ws + w.hidden), (const float *)(ws + w.aux),
                                               (const float *)(ws + w.gG), scal, w_sparse, w_excl, w_orth, w_tpm, d_gpi,
                                               d_gw, d_gA, d_terms);
    CY2P_CUDA(cudaGetLastError());
    return CY2P_OK;
}

}  // extern "C"
