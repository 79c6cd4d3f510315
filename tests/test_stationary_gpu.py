"""GPU test of the device stationary-state estimate: block subspace iteration on the K1 kernel
against the host ARPACK answer (cy2path_b200.utils.eigs) and an analytic stationary distribution."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _symmetric_walk(n, k, seed):
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k)
    cols = rng.integers(0, n, n * k)
    A = sp.coo_matrix((np.ones(n * k), (rows, cols)), shape=(n, n)).tocsr()
    ring = sp.diags([np.ones(n - 1), np.ones(n - 1)], [1, -1], shape=(n, n))
    A = ((A + A.T + ring + sp.identity(n)) > 0).astype(np.float64)
    deg = np.asarray(A.sum(1)).ravel()
    T = sp.csr_matrix(sp.diags(1.0 / deg) @ A)
    T.data = T.data.astype(np.float32)
    T.sort_indices()
    return T, deg


class _Adata:
    def __init__(self, T, end=None):
        self.obsp = {"T_forward": T}
        self.shape = (T.shape[0], 0)
        n = T.shape[0]
        self.obs = pd.DataFrame({"root_cells": np.ones(n), "end_points": np.ones(n) if end is None else end})
        self.uns = {}


def test_device_eigs_unique_stationary_distribution():
    from cy2path_b200.utils import eigs, eigs_device, estimate_stationary_state

    T, deg = _symmetric_walk(20000, 8, seed=5)
    vals, V = eigs_device(T)
    assert len(vals) == 1 and abs(vals[0] - 1) < 1e-6
    pi = V[:, 0] / V[:, 0].sum()
    # float32 T: rows sum to 1 +- 6e-8, so pi ~ degree to ~1e-6 in L1
    assert np.abs(pi - deg / deg.sum()).sum() < 1e-5
    rv, rV = eigs(T)
    assert len(rv) == 1 and np.abs(rV[:, 0] / rV[:, 0].sum() - pi).sum() < 1e-6
    ad = _Adata(T)
    s_dev = estimate_stationary_state(ad, method="device")
    s_host = estimate_stationary_state(ad, method="arpack")
    assert s_dev.shape == (20000,) and np.abs(s_dev - s_host).sum() < 1e-6
    with pytest.raises(ValueError):
        estimate_stationary_state(ad, method="nope")


def test_device_eigs_degenerate_falls_back_to_end_points(monkeypatch):
    from cy2path_b200.synth import knn_velocity_tpm
    from cy2path_b200.utils import eigs, eigs_device, estimate_stationary_state

    g = knn_velocity_tpm(5000, k=15, seed=3)
    vals, V = eigs_device(g["T"])
    assert len(vals) == len(eigs(g["T"])[0]) and len(vals) > 1  # one eigenvalue 1 per absorbing class
    ad = _Adata(g["T"], end=g["end_points"])
    monkeypatch.setenv("CY2P_STATIONARY", "device")
    s = estimate_stationary_state(ad)
    assert np.allclose(s, g["end_points"] / g["end_points"].sum())
