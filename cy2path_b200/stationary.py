"""Dominant left eigenpairs of the transition matrix by block subspace iteration.

The reference asks scvelo's ``eigs`` (ARPACK on the host) for the eigenvalues of T^T within ``eps`` of 1 and uses
the single leading eigenvector as the stationary distribution (utils.py:104-118). The same question is answered
here with the K1 kernel: X <- X . T for a block of k start vectors IS one batched propagation update
(cy2p_propagate_rows_f64, B = k), so the only other work is a k-column QR and a k x k Rayleigh-Ritz problem
every few updates (torch.linalg on the device, tiny).

``apply_T(X) -> X . T`` is passed in: the product path hands the CUDA update, the CPU tests a scipy product, so the
algorithm is tested without a GPU while the kernels stay the only device arithmetic.
"""
from __future__ import annotations

import numpy as np


def dominant_left_eigenpairs(apply_T, n, k=10, eps=1e-3, tol=1e-7, max_updates=20000, block=8, device=None, seed=0):
    """Ritz pairs (theta, V) of the left eigenproblem v T = theta v for the k dominant eigenvalues, those with
    real part >= 1 - eps kept, ordered by decreasing real part, eigenvectors in absolute value (as ``eigs``).

    apply_T: float64 tensor [n, k] -> [n, k], column-wise x . T.  Stops when every kept pair's residual
    ||v T - theta v|| / ||v|| is below ``tol`` (and the number kept did not change) or after ``max_updates``.
    Returns (values [m] float64 numpy, vectors [n, m] float64 numpy, info dict)."""
    import torch

    k = int(max(1, min(k, n - 2 if n > 2 else 1)))
    gen = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.rand((n, k), dtype=torch.float64, generator=gen).to(device)
    X[:, 0] = 1.0 / n  # the uniform vector: a good start for the stationary direction
    updates, kept_prev, info = 0, -1, {"converged": False}
    theta = vecs = None
    while updates < max_updates:
        for _ in range(block):
            X = apply_T(X)
        updates += block
        Q, _ = torch.linalg.qr(X)
        Z = apply_T(Q)
        updates += 1
        H = Q.T @ Z
        w, Y = torch.linalg.eig(H)  # k x k, complex
        order = torch.argsort(w.real, descending=True)
        w, Y = w[order], Y[:, order]
        keep = w.real >= 1 - eps
        m = int(keep.sum())
        R = Z.to(Y.dtype) @ Y - (Q.to(Y.dtype) @ Y) * w[None, :]
        res = torch.linalg.norm(R, dim=0)  # ||Q y|| = 1
        theta, vecs = w, Q.to(Y.dtype) @ Y
        worst = float(res[:max(m, 1)].max())
        # also require the first rejected Ritz value to be settled, or the count could still change
        settled = m == k or float(res[m]) < max(tol, 0.1 * abs(float(w.real[m]) - (1 - eps)))
        if m == kept_prev and worst < tol and settled:
            info["converged"] = True
            break
        kept_prev = m
        X = Z  # continue from the orthonormal basis (one update ahead)
    m = int((theta.real >= 1 - eps).sum())
    info.update(updates=updates, residual=float(res[:max(m, 1)].max()), ritz=theta.real.cpu().numpy())
    vals = theta.real[:m].cpu().numpy()
    V = np.absolute(vecs[:, :m].real.cpu().numpy())
    return vals, V, info
