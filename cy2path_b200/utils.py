"""Host-side boundary glue mirroring cy2path/utils.py (reference utils.py:47-118).

These are input checks, not arithmetic: they run on the host exactly as in the reference.
``anndata`` is never imported: any object with ``.obsp / .obs / .uns / .shape / .copy()`` works.
"""
from __future__ import annotations

import collections.abc
import logging

import numpy as np
from scipy.sparse import csr_matrix, issparse

logger = logging.getLogger("cy2path_b200")


def check_TPM(adata, matrix_key="T_forward", recalc_matrix=False, self_transitions=False):
    """reference utils.py:47-77. The recompute branches call scvelo (``get_transition_matrix`` /
    ``terminal_states``), which builds the *input* of the hot path and is out of scope here: a
    missing matrix or missing root/end annotations raise instead of being recomputed."""
    if recalc_matrix:
        raise NotImplementedError(
            "recalc_matrix=True needs scvelo.utils.get_transition_matrix (input construction, outside the "
            "accelerated path); compute adata.obsp[matrix_key] with scvelo first")
    try:
        T = adata.obsp[matrix_key]
        assert T.shape
    except Exception:
        raise ValueError(f"adata.obsp['{matrix_key}'] is missing; compute the transition matrix with scvelo first")
    if not issparse(T):
        adata.obsp[matrix_key] = csr_matrix(T)  # reference utils.py:68-69
    obs_cols = getattr(adata.obs, "columns", adata.obs)
    if "root_cells" not in obs_cols or "end_points" not in obs_cols:
        raise ValueError("adata.obs needs 'root_cells' and 'end_points' (scvelo.tl.terminal_states)")


def check_root_init(adata, init="root_cells"):
    """reference utils.py:81-100 (with collections.abc so custom inits work on Python >= 3.10)."""
    if isinstance(init, str):
        if init == "root_cells":
            rc = np.asarray(adata.obs["root_cells"], dtype=np.float64)
            return rc / rc.sum(), "root_cells"
        if init == "uniform":
            n = adata.shape[0]
            return [1 / n] * n, "uniform"  # the reference returns a Python list here
        raise ValueError("Incorrect initialisation of state probabilities.")
    if isinstance(init, (collections.abc.Sequence, np.ndarray)):
        return init, "custom"
    raise ValueError("Incorrect initialisation of state probabilities.")


def eigs(T, k=10, eps=1e-3):
    """Restatement of scvelo.tools.terminal_states.eigs as recalled (source not available offline —
    unverified): the k eigenpairs of T^T with largest real part, those with eigenvalue
    >= 1 - eps kept, eigenvectors in absolute value."""
    from scipy.sparse.linalg import eigs as sp_eigs

    try:
        k = min(k, T.shape[0] - 2)
        vals, vecs = sp_eigs(T.T.astype(np.float64), k=k, which="LR")
        order = np.argsort(vals)[::-1]
        vals, vecs = vals.real[order], vecs.real[:, order]
        keep = vals >= 1 - eps
        return vals[keep], np.absolute(vecs[:, keep])
    except Exception:
        return np.empty(0), np.zeros((T.shape[0], 0))


def eigs_device(T, k=10, eps=1e-3, tol=1e-7, max_updates=20000):
    """``eigs`` answered on the device: block subspace iteration whose update is the K1 kernel with B = k columns
    in float64 (cy2path_b200/stationary.py). Returns (values, |vectors|) like ``eigs``."""
    import torch

    from .dist import cuda_row_update
    from .sampling import get_plan
    from .stationary import dominant_left_eigenpairs

    plan = get_plan(T)
    update = cuda_row_update(plan)

    def apply_T(X):
        X = X.contiguous()
        Y = torch.empty_like(X)
        update(X, Y, 0, plan.n)
        return Y

    vals, vecs, info = dominant_left_eigenpairs(apply_T, plan.n, k=k, eps=eps, tol=tol, max_updates=max_updates,
                                                device=plan.device)
    if not info["converged"]:
        logger.warning("subspace iteration stopped at %d updates with residual %.2e", info["updates"], info["residual"])
    return vals, vecs


def estimate_stationary_state(adata, matrix_key="T_forward", method=None):
    """reference utils.py:104-118: the single leading left eigenvector, else the end_points prior.

    ``method``: "arpack" (host scipy, the default: closest to the scvelo call the reference makes) or "device"
    (``eigs_device``); the environment variable CY2P_STATIONARY sets the default."""
    import os

    method = method or os.environ.get("CY2P_STATIONARY", "arpack")
    if method not in ("arpack", "device"):
        raise ValueError("method must be 'arpack' or 'device'")
    vecs = (eigs_device if method == "device" else eigs)(adata.obsp[matrix_key])[1]
    if vecs.shape[1] == 1:
        stat = vecs / vecs.sum()
    else:
        ep = np.asarray(adata.obs["end_points"], dtype=np.float64)
        stat = ep / ep.sum()
        logger.warning("Multiple eigenvalues > 1, falling back to end_points to infer stationary distribution.")
    return np.asarray(stat).flatten()
