"""TEST INFRASTRUCTURE — CPU restatement of the reference's MSM state-probability simulation.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the
checker or the timed CPU baseline — never as the thing shipped.

Pinned against: outputs of the reference itself (``tests/golden/msm_*.npz`` produced by
``tests/golden/make_golden.py`` running ``/root/reference/cy2path/sampling.py`` unmodified).
The reference has no tests or golden vectors of its own.

Third-party arithmetic on this path that is NOT under /root/reference: scipy
``_sparsetools.csc_matvec`` (scipy 1.18.1 here, unpinned by the reference). Its
published algorithm is restated in ``spmv_left`` below and in ``oracle/csrc/oracle.c``.
"""
import math

import numpy as np


def check_convergence_criteria(eps_history, tol=1e-5):
    """reference: cy2path/sampling.py:13-25.

    d[i] = |eps[i+1]-eps[i]|; returns 1 + len(d) - (index from the end of the last d > tol),
    i.e. (last i with d[i] > tol) + 2, or None when no difference exceeds tol."""
    d = [abs(eps_history[i + 1] - eps_history[i]) for i in range(len(eps_history) - 1)]
    for back, v in enumerate(d[::-1]):
        if v > tol:
            return 1 + len(d) - back
    return None


def spmv_left(x, T):
    """y = x @ T exactly as scipy evaluates it for ndarray @ csr_matrix.

    reference call site: cy2path/sampling.py:40-42 (``csr_matrix.dot(x, T)``).
    scipy turns this into ``T.T._matmul_vector(x)``; T.T of a CSR is a CSC view and
    ``csc_matvec`` runs  for i in rows(T): for p in indptr[i]..indptr[i+1]: y[indices[p]] += data[p]*x[i]
    with y, and the products, in the upcast dtype (float64 for float32 data and float64 x).
    Pure-Python loops: small cases only. ``spmv_left_fast`` is the same arithmetic in C."""
    n = T.shape[0]
    out_dtype = np.result_type(T.dtype, np.asarray(x).dtype)
    y = np.zeros(n, dtype=out_dtype)
    x = np.asarray(x, dtype=out_dtype)
    indptr, indices, data = T.indptr, T.indices, T.data.astype(out_dtype)
    for i in range(n):
        xi = x[i]
        for p in range(indptr[i], indptr[i + 1]):
            y[indices[p]] += data[p] * xi
    return y


def cosine_distance(u, v):
    """scipy.spatial.distance.cosine (scipy 1.18.1 ``correlation(centered=False)``):
    1 - u.v / sqrt(u.u * v.v), clipped to [0, 2]. reference call site: sampling.py:44."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    uv = float(np.dot(u, v))
    uu = float(np.dot(u, u))
    vv = float(np.dot(v, v))
    d = 1.0 - uv / math.sqrt(uu * vv)
    return min(max(d, 0.0), 2.0)


def iterate_state_probability(T, init, stationary, max_iter=1000, tol=1e-5):
    """reference: cy2path/sampling.py:29-60.

    Row i of the history holds the iterate *before* update i; eps[i] is the cosine
    distance of the *updated* iterate to the stationary vector (sampling.py:39-45).
    A ``None`` / ``>= max_iter`` convergence index is mapped to max_iter (sampling.py:49-57).
    Returns (state_history, state_history_max_iter, convergence_check, eps_history)."""
    n = T.shape[0]
    hist = np.empty((max_iter, n))
    eps = np.empty(max_iter)
    x = init
    for i in range(max_iter):
        hist[i] = x
        x = x @ T  # scipy csc_matvec; same arithmetic as spmv_left
        eps[i] = cosine_distance(x, stationary)
    conv = check_convergence_criteria(eps, tol=tol)
    if not (isinstance(conv, int) and conv < max_iter):
        conv = max_iter
    return hist[:conv], hist, conv, eps


def propagate_batch(T, X0, iters):
    """Batched form of the same loop: X0 is [n, B] (cell-major); returns X after ``iters`` steps.

    scipy's batched ``X.T @ T`` is bitwise equal to per-start ``x @ T``, so one
    batched call is a valid oracle at any B. float64 accumulation as in the reference."""
    X = np.asarray(X0, dtype=np.float64).T.copy()  # [B, n]
    for _ in range(iters):
        X = np.asarray(X @ T)
    return X.T.copy()


def rel_l1(a, b):
    """Relative L1 error used by the parity bar (north_star: <= 1e-6)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300))
