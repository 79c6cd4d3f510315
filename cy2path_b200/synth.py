"""Synthetic kNN velocity transition-probability matrices.

The reference ships no data; every benchmark/parity input is generated here,
deterministically from a seed, with the shape of a scvelo ``T_forward``:
row-stochastic float32 CSR, int32 indices, k out-neighbours per cell.

This is input generation (host numpy/scipy), not part of the hot path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _trajectory(n, d, rng):
    """Three branches leaving a common trunk, embedded in R^d with isotropic noise."""
    tau = rng.random(n)
    branch = rng.integers(0, 3, n)
    dirs = rng.standard_normal((4, d))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    trunk, split = dirs[0], 0.3
    pos = np.empty((n, d))
    drift = np.empty((n, d))
    on_trunk = tau < split
    pos[on_trunk] = tau[on_trunk, None] * trunk
    drift[on_trunk] = trunk
    for b in range(3):
        m = (~on_trunk) & (branch == b)
        v = dirs[1 + b]
        pos[m] = split * trunk + (tau[m, None] - split) * v
        drift[m] = v
    pos += 0.05 * rng.standard_normal((n, d))
    return tau, pos, drift


def _knn(pos, k):
    """Exact k nearest neighbours excluding self -> int64 [n, k]."""
    n = pos.shape[0]
    try:
        import torch

        if torch.cuda.is_available() and n > 100_000:
            return _knn_torch(pos, k)
    except Exception:  # pragma: no cover - torch always present in this image
        pass
    from scipy.spatial import cKDTree

    tree = cKDTree(pos)
    _, idx = tree.query(pos, k=k + 1, workers=-1)
    mask = idx != np.arange(n)[:, None]
    noself = mask.all(1)  # self pushed out by exact-distance ties: drop the farthest instead
    mask[noself, -1] = False
    dup = mask.sum(1) != k  # self listed more than once cannot happen; guard anyway
    if dup.any():
        raise RuntimeError("kNN self-removal failed")
    return idx[mask].reshape(n, k).astype(np.int64)


def _knn_torch(pos, k, block=2048):  # pragma: no cover - GPU only
    import torch

    x = torch.as_tensor(pos, dtype=torch.float32, device="cuda")
    sq = (x * x).sum(1)
    n = x.shape[0]
    out = torch.empty((n, k), dtype=torch.int64, device="cuda")
    for s in range(0, n, block):
        e = min(n, s + block)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (x[s:e] @ x.T)
        d2[torch.arange(e - s, device="cuda"), torch.arange(s, e, device="cuda")] = float("inf")
        out[s:e] = d2.topk(k, dim=1, largest=False).indices
    return out.cpu().numpy()


def knn_velocity_tpm(n, k=30, d=8, seed=0, variable_degree=False):
    """Build a synthetic velocity TPM.

    Returns a dict with
      T            scipy CSR float32 [n, n], int32 indices/indptr, sorted columns,
                   every row has exactly k stored entries (1..k if variable_degree)
      tau          latent time per cell (float64)
      root_cells   float64 0/1 indicator of the 1% lowest-tau cells
      end_points   float64 0/1 indicator of the 1% highest-tau cells
    """
    rng = np.random.default_rng(seed)
    tau, pos, drift = _trajectory(n, d, rng)
    nbr = _knn(pos, k)
    rows = np.repeat(np.arange(n), k)
    cols = nbr.reshape(-1)
    disp = pos[cols] - pos[rows]
    disp /= np.maximum(np.linalg.norm(disp, axis=1, keepdims=True), 1e-12)
    cosv = np.einsum("ij,ij->i", disp, drift[rows])
    if variable_degree:
        keep = cosv > 0
        # make sure every row keeps at least its best neighbour
        best = cosv.reshape(n, k).argmax(1) + np.arange(n) * k
        keep[best] = True
        w = np.expm1(10.0 * np.clip(cosv, 0.0, 1.0)) + 1e-3
        rows, cols, w = rows[keep], cols[keep], w[keep]
    else:
        # small floor keeps all k entries strictly positive (scvelo's TPMs have no stored zeros)
        w = np.expm1(10.0 * np.clip(cosv, 0.0, 1.0)) + 1e-3
    T = sp.csr_matrix((w, (rows, cols)), shape=(n, n), dtype=np.float64)
    T.sum_duplicates()
    T.sort_indices()
    rs = np.asarray(T.sum(1)).ravel()
    T = sp.diags(1.0 / rs) @ T
    T = sp.csr_matrix(T, dtype=np.float32)
    T.sort_indices()
    T.indices = T.indices.astype(np.int32)
    T.indptr = T.indptr.astype(np.int32)

    m = max(1, n // 100)
    order = np.argsort(tau, kind="stable")
    root = np.zeros(n)
    root[order[:m]] = 1.0
    end = np.zeros(n)
    end[order[-m:]] = 1.0
    return {"T": T, "tau": tau, "root_cells": root, "end_points": end, "pos": pos}


def stationary_by_power(T, iters=300):
    """Normalised stationary estimate by repeated x <- x @ T from uniform (host, float64).

    Used for large synthetic graphs where ARPACK is slow; the vector is only an
    *input* (the pi the cosine distance is measured against)."""
    n = T.shape[0]
    x = np.full(n, 1.0 / n)
    Tt = sp.csr_matrix(T.T, dtype=np.float64)
    for _ in range(iters):
        x = Tt @ x
    x = np.abs(x)
    return x / x.sum()


def batched_starts(T, B, seed):
    """B start distributions, each uniform over a random cell and its out-neighbours.

    Returns float32 [n, B] in the cell-major layout the kernels use."""
    n = T.shape[0]
    rng = np.random.default_rng(seed + 1)
    cells = rng.choice(n, B)
    X0 = np.zeros((n, B), dtype=np.float32)
    for b, c in enumerate(cells):
        nb = T.indices[T.indptr[c]:T.indptr[c + 1]]
        sup = np.unique(np.concatenate([[c], nb]))
        X0[sup, b] = np.float32(1.0 / len(sup))
    return X0


def chain_inputs(T, root_prior, C, max_iter, seed):
    """Initial states and the float64 uniforms each chain consumes (bit-exact sampler parity)."""
    n = T.shape[0]
    init_state = np.random.default_rng(seed + 2).choice(n, C, p=root_prior / root_prior.sum()).astype(np.int32)
    uniforms = np.random.default_rng(seed + 3).random((C, max_iter - 1))
    return init_state, uniforms
