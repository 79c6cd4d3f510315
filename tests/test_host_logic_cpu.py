"""CPU tests of host-side logic that has a device twin (same comparisons on torch tensors)."""
import numpy as np
import torch

from cy2path_b200.sampling import batch_convergence, batch_convergence_device, check_convergence_criteria


def test_batch_convergence_variants_agree():
    rng = np.random.default_rng(4)
    iters, B, tol = 60, 37, 1e-3
    eps = np.abs(rng.standard_normal((iters, B))) * np.exp(-np.arange(iters)[:, None] / rng.uniform(2, 30, B)[None, :])
    eps[:, 0] = 0.5          # never moves: no difference above tol -> max_iter
    eps[:, 1] = np.arange(iters)  # always moving -> max_iter
    eps[-1, 2] += 1.0        # last difference above tol -> iters + ... clipped to max_iter
    host = batch_convergence(eps, tol, iters)
    dev = batch_convergence_device(torch.from_numpy(eps), tol, iters).numpy()
    assert np.array_equal(host, dev)
    for b in range(B):  # the reference's per-start rule (sampling.py:13-25, 48-57)
        c = check_convergence_criteria(eps[:, b], tol)
        want = c if (c is not None and c < iters) else iters
        assert host[b] == want
    one = batch_convergence_device(torch.from_numpy(eps[:1]), tol, 1).numpy()
    assert np.array_equal(one, np.full(B, 1))


def _ring_of_random_neighbours(n, k, seed):
    """Connected, aperiodic, symmetric graph -> reversible walk with stationary distribution ~ degree."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k)
    cols = rng.integers(0, n, n * k)
    A = sp.coo_matrix((np.ones(n * k), (rows, cols)), shape=(n, n)).tocsr()
    ring = sp.diags([np.ones(n - 1), np.ones(n - 1)], [1, -1], shape=(n, n))
    A = ((A + A.T + ring + sp.identity(n)) > 0).astype(np.float64)
    deg = np.asarray(A.sum(1)).ravel()
    return sp.csr_matrix(sp.diags(1.0 / deg) @ A), deg


def test_subspace_iteration_single_and_degenerate_eigenvalue():
    """The algorithm of cy2path_b200/stationary.py with a scipy product standing in for the CUDA update."""
    import scipy.sparse as sp

    from cy2path_b200.stationary import dominant_left_eigenpairs
    from cy2path_b200.synth import knn_velocity_tpm
    from cy2path_b200.utils import eigs

    T, deg = _ring_of_random_neighbours(600, 6, seed=2)
    Tt = sp.csr_matrix(T.T)
    vals, V, info = dominant_left_eigenpairs(lambda X: torch.from_numpy(Tt @ X.numpy()), 600)
    assert info["converged"] and len(vals) == 1 and abs(vals[0] - 1) < 1e-8
    assert np.abs(V[:, 0] / V[:, 0].sum() - deg / deg.sum()).sum() < 1e-6
    rv, rV = eigs(T)
    assert len(rv) == 1 and np.abs(rV[:, 0] / rV[:, 0].sum() - V[:, 0] / V[:, 0].sum()).sum() < 1e-6

    g = knn_velocity_tpm(1500, k=12, seed=3)  # three absorbing classes: eigenvalue 1 three times
    Tt = sp.csr_matrix(g["T"].T.astype(np.float64))
    vals, V, info = dominant_left_eigenpairs(lambda X: torch.from_numpy(Tt @ X.numpy()), 1500)
    assert info["converged"] and len(vals) == len(eigs(g["T"])[0]) and len(vals) > 1
    assert np.allclose(vals, 1.0, atol=1e-6) and V.shape == (1500, len(vals))


def test_plan_cache_fingerprint_sees_in_place_edits():
    """the plan cache key must change when T is edited in place (values, column indices or row pointers),
    also far from a strided sample of the values."""
    import scipy.sparse as sp

    from cy2path_b200.sampling import _fingerprint

    rng = np.random.default_rng(0)
    T = sp.random(3000, 3000, density=0.03, format="csr", random_state=1, dtype=np.float32)
    assert T.nnz > 200_000
    base = _fingerprint(T)
    assert _fingerprint(T) == base
    for k in rng.integers(0, T.nnz, 20):
        old = T.data[k]
        T.data[k] = old * 0.5 + 0.125  # e.g. making a terminal row absorbing
        assert _fingerprint(T) != base
        T.data[k] = old
    assert _fingerprint(T) == base
    k = int(rng.integers(1, T.nnz - 1))
    old = T.indices[k]
    T.indices[k] = (old + 1) % 3000
    assert _fingerprint(T) != base
    T.indices[k] = old
    a, b = T.data[10], T.data[11]
    if a != b:  # swapping two values keeps the plain sum; the position-weighted sample or the CRC must catch it
        T.data[10], T.data[11] = b, a
        assert _fingerprint(T) != base
