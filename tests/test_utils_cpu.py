"""CPU tests of the host-side boundary glue (cy2path_b200/utils.py) against the behaviour of the reference's
utils.py:47-118: input checks and the stationary-state rule. No CUDA involved."""
import logging

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from cy2path_b200 import utils
from cy2path_b200.synth import knn_velocity_tpm, stationary_by_power


class Duck:
    def __init__(self, T, root=None, end=None):
        self.obsp = {} if T is None else {"T_forward": T}
        self.shape = (0 if T is None else T.shape[0], 0)
        self.obs = pd.DataFrame({} if root is None else {"root_cells": root, "end_points": end})
        self.uns = {}


def _graph(n=300, k=8, seed=4):
    g = knn_velocity_tpm(n, k=k, seed=seed)
    return g["T"], g["root_cells"], g["end_points"]


def test_check_root_init_types_follow_the_reference():
    T, root, end = _graph()
    ad = Duck(T, root, end)
    p, kind = utils.check_root_init(ad, "root_cells")  # reference utils.py:84-89
    assert kind == "root_cells" and isinstance(p, np.ndarray) and p.dtype == np.float64
    assert np.array_equal(p, root / root.sum())
    p, kind = utils.check_root_init(ad, "uniform")  # reference utils.py:90-94: a Python list
    assert kind == "uniform" and isinstance(p, list) and len(p) == T.shape[0] and p[0] == 1 / T.shape[0]
    custom = np.full(T.shape[0], 1.0 / T.shape[0])
    p, kind = utils.check_root_init(ad, custom)  # reference utils.py:95-97
    assert kind == "custom" and p is custom
    assert utils.check_root_init(ad, list(custom))[1] == "custom"
    for bad in ("end_points", 3, None):
        with pytest.raises(ValueError):
            utils.check_root_init(ad, bad)


def test_check_tpm_coerces_dense_and_names_what_is_missing():
    T, root, end = _graph()
    ad = Duck(T.toarray(), root, end)
    utils.check_TPM(ad)  # reference utils.py:68-69: a dense matrix becomes CSR in place
    assert sp.issparse(ad.obsp["T_forward"]) and ad.obsp["T_forward"].format == "csr"
    assert (ad.obsp["T_forward"] != T).nnz == 0
    with pytest.raises(NotImplementedError):  # scvelo's get_transition_matrix is input construction
        utils.check_TPM(Duck(T, root, end), recalc_matrix=True)
    with pytest.raises(ValueError, match="T_forward"):
        utils.check_TPM(Duck(None, root, end))
    with pytest.raises(ValueError, match="root_cells"):
        utils.check_TPM(Duck(T))


def test_stationary_state_single_eigenvalue_and_end_point_fallback(caplog):
    # an irreducible walk: one eigenvalue at 1 -> the normalised leading left eigenvector (reference utils.py:107-108)
    n = 200
    A = sp.random(n, n, density=0.05, random_state=1, format="csr") + sp.identity(n) + sp.diags([np.ones(n - 1)], [1])
    A = sp.csr_matrix(A + sp.csr_matrix(([1.0], ([n - 1], [0])), shape=(n, n)))
    T = sp.csr_matrix(sp.diags(1.0 / np.asarray(A.sum(1)).ravel()) @ A, dtype=np.float32)
    end = np.zeros(n)
    end[-3:] = 1
    stat = utils.estimate_stationary_state(Duck(T, np.ones(n), end), method="arpack")
    assert stat.shape == (n,) and abs(stat.sum() - 1) < 1e-12 and (stat >= 0).all()
    assert np.abs(stat - stationary_by_power(T, 2000)).sum() < 1e-6
    # several absorbing classes: eigenvalue 1 more than once -> the end_points prior, with the reference's warning
    Tk, root, endk = _graph(1500, 12, 3)
    with caplog.at_level(logging.WARNING, logger="cy2path_b200"):
        stat = utils.estimate_stationary_state(Duck(Tk, root, endk), method="arpack")
    assert np.array_equal(stat, endk / endk.sum())
    assert any("falling back to end_points" in r.message for r in caplog.records)
    with pytest.raises(ValueError):
        utils.estimate_stationary_state(Duck(Tk, root, endk), method="lanczos")
