"""Host-only emulation of candidate iterate storage formats for the batched propagation (K1).

Question: which storage / accumulation format keeps the batched update inside
rel-L1 1e-6 of the reference's float64 iterate (`sampling.py:33,40-42`) after 2,000 iterations?

Formats (all on the float32 T, sources ascending per destination like scipy's csc_matvec):
  f32         fp32 FMA chain, fp32 storage (round 1's default)
  f32_exact   exact (float64) accumulation, fp32 storage
  f32_sr      exact accumulation, stochastic rounding to fp32
  td<m>       "truncated double": float64 arithmetic, iterate stored with an m-bit mantissa (round to nearest)
  f64         float64 everything (the reference)
Usage: python tools/precision_lab.py [n] [starts] [iters]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from cy2path_b200.synth import knn_velocity_tpm  # noqa: E402


def ell_of_transpose(T):
    Tt = T.T.tocsr()
    Tt.sort_indices()
    n = T.shape[0]
    deg = np.diff(Tt.indptr)
    W = int(deg.max())
    src = np.zeros((n, W), dtype=np.int64)
    val = np.zeros((n, W), dtype=np.float32)
    for j in range(n):
        a, b = Tt.indptr[j], Tt.indptr[j + 1]
        src[j, : b - a] = Tt.indices[a:b]
        val[j, : b - a] = Tt.data[a:b]
    return src, val


def round_mantissa(y, m):
    """Round float64 y to an m-bit stored mantissa (m explicit bits), nearest-even."""
    if m >= 52:
        return y
    b = y.view(np.int64)
    drop = 52 - m
    half = np.int64(1) << (drop - 1)
    lsb = (b >> drop) & 1
    b2 = (b + half - 1 + lsb) >> drop << drop
    return b2.view(np.float64)


def step(X, src, val, fmt, rng):
    n, W = src.shape
    if fmt.startswith("split"):
        # X = (hi fp32, lo float64 residual possibly quantised); hi-sum by fp32 FMA chain, lo-sum exact
        hi, lo = X
        acc = np.zeros(hi.shape, dtype=np.float32)
        accl = np.zeros(hi.shape, dtype=np.float64)
        for k in range(W):
            v = val[:, k, None].astype(np.float64)
            acc = (acc.astype(np.float64) + v * hi[src[:, k]].astype(np.float64)).astype(np.float32)
            accl += v * lo[src[:, k]]
        y = acc.astype(np.float64) + accl
        nh = y.astype(np.float32)
        nl = y - nh.astype(np.float64)
        if fmt == "split_bf16":
            nl = round_mantissa(nl, 7)
        return (nh, nl)
    if fmt == "f32":
        acc = np.zeros(X.shape, dtype=np.float32)
        for k in range(W):
            acc = (acc.astype(np.float64) + val[:, k, None].astype(np.float64) * X[src[:, k]].astype(np.float64)).astype(np.float32)
        return acc
    acc = np.zeros(X.shape, dtype=np.float64)
    for k in range(W):
        acc += val[:, k, None].astype(np.float64) * X[src[:, k]].astype(np.float64)
    if fmt == "f64":
        return acc
    if fmt == "f32_exact":
        return acc.astype(np.float32)
    if fmt == "f32_sr":
        lo = acc.astype(np.float32)
        lo = np.where(lo.astype(np.float64) > acc, np.nextafter(lo, np.float32(-np.inf)), lo)
        hi = np.nextafter(lo, np.float32(np.inf))
        gap = hi.astype(np.float64) - lo.astype(np.float64)
        p = np.where(gap > 0, (acc - lo.astype(np.float64)) / np.where(gap > 0, gap, 1), 0.0)
        return np.where(rng.random(acc.shape) < p, hi, lo).astype(np.float32)
    if fmt.startswith("td"):
        return round_mantissa(acc, int(fmt[2:]))
    raise ValueError(fmt)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3696
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    fmts = sys.argv[4].split(",") if len(sys.argv) > 4 else ["f32", "f32_exact", "f32_sr", "td24", "td28", "td32", "td36"]
    g = knn_velocity_tpm(n, 30, seed=1)
    T = g["T"]
    src, val = ell_of_transpose(T)
    rng = np.random.default_rng(5)
    starts = rng.choice(n, B, replace=False)
    X0 = np.zeros((n, B))
    for b, c in enumerate(starts):
        nb = T.indices[T.indptr[c]: T.indptr[c + 1]]
        X0[c, b] = 1
        X0[nb, b] = 1
        X0[:, b] /= X0[:, b].sum()
    marks = [10, 100, 500, 1000, 2000]
    ref = X0.copy()
    refs = {}
    for t in range(1, iters + 1):
        ref = step(ref, src, val, "f64", rng)
        if t in marks:
            refs[t] = ref.copy()
    for fmt in fmts:
        X = X0.astype(np.float32) if fmt.startswith("f32") else X0.copy()
        if fmt.startswith("split"):
            X = (X0.astype(np.float32), X0 - X0.astype(np.float32).astype(np.float64))
        out = []
        for t in range(1, iters + 1):
            X = step(X, src, val, fmt, rng)
            if t in refs:
                r = refs[t]
                Xs = X
                X = X[0].astype(np.float64) + X[1] if isinstance(X, tuple) else X
                rel = (np.abs(X - r).sum(0) / np.abs(r).sum(0)).max()
                out.append(f"{t}: relL1 {rel:.2e} maxabs {np.abs(X - r).max():.1e}")
                X = Xs
        print(fmt, " | ".join(out), flush=True)


if __name__ == "__main__":
    main()
