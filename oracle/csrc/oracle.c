/* TEST INFRASTRUCTURE — plain-C restatement of the reference's CPU arithmetic.
 *
 * Not product code. Built by oracle/build.py (gcc) into oracle/_build/liboracle.so and
 * loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * Each function cites the reference line it follows; the sparse product itself lives in
 * scipy (_sparsetools csc_matvec / csc_matvecs, scipy 1.18.1, unpinned by the reference),
 * whose published loop is restated here. Pinned by tests/test_oracle.py against scipy
 * (bitwise) and against the golden vectors produced by the unmodified reference.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* y = x @ T for CSR T (float32 data), x and y float64.
 * reference: cy2path/sampling.py:40-42 -> scipy csc_matvec on T.T:
 *   for i in rows(T): for p in [indptr[i], indptr[i+1]): y[indices[p]] += data[p] * x[i]   */
void orc_spmv_left(int n, const int32_t *indptr, const int32_t *indices, const float *data,
                   const double *x, double *y) {
    memset(y, 0, sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const double xi = x[i];
        for (int p = indptr[i]; p < indptr[i + 1]; ++p) y[indices[p]] += (double)data[p] * xi;
    }
}

/* Batched Y[b,:] = X[b,:] @ T, X and Y row-major [B][n] float64 — the same scatter per row b;
 * scipy evaluates the batched product column by column of X^T with identical arithmetic
 * (bitwise equal to the per-vector call). */
void orc_spmm_left(int n, int B, const int32_t *indptr, const int32_t *indices, const float *data,
                   const double *X, double *Y) {
    for (int b = 0; b < B; ++b)
        orc_spmv_left(n, indptr, indices, data, X + (size_t)b * n, Y + (size_t)b * n);
}

/* scipy.spatial.distance.cosine: 1 - u.v/sqrt(u.u*v.v) clipped to [0,2]. reference: sampling.py:44 */
double orc_cosine(int n, const double *u, const double *v) {
    double uv = 0, uu = 0, vv = 0;
    for (int i = 0; i < n; ++i) { uv += u[i] * v[i]; uu += u[i] * u[i]; vv += v[i] * v[i]; }
    double d = 1.0 - uv / sqrt(uu * vv);
    return d < 0 ? 0 : (d > 2 ? 2 : d);
}

/* One trajectory. reference: cy2path/simulation.py:74-85.
 * j = first index with r <= (double)cum[state][j]; returns -1-k on CDF overshoot at step k
 * (the reference raises IndexError there). prob = sum of T[prev, next] over stored duplicates. */
int orc_chain(int W, const int32_t *idx, const float *cum, const int32_t *indptr,
              const int32_t *indices, const float *data, int init_state, int steps,
              const double *uniforms, int32_t *states, float *probs) {
    states[0] = init_state;
    for (int k = 1; k <= steps; ++k) {
        const int prev = states[k - 1];
        const double r = uniforms[k - 1];
        int j = -1;
        for (int c = 0; c < W; ++c)
            if (r <= (double)cum[(size_t)prev * W + c]) { j = c; break; }
        if (j < 0) return -k;
        const int nxt = idx[(size_t)prev * W + j];
        states[k] = nxt;
        float acc = 0.f;
        for (int p = indptr[prev]; p < indptr[prev + 1]; ++p)
            if (indices[p] == nxt) acc += data[p];
        probs[k - 1] = acc;
    }
    return 0;
}
