"""GPU parity tests for a user-supplied float64 transition matrix.

scvelo writes `T_forward` as float32, but nothing in the reference forces that: with a float64 matrix
`sampling.py:40-42` multiplies by the float64 values and `simulation.py:24-25` accumulates the CDF in float64 before
the float32 store. The plan then keeps the values as hi + lo float32 pairs (`cy2p_plan_create_host_f64`) and the
float64-arithmetic paths use both halves. The oracle here is scipy/numpy on the float64 matrix itself, which is what
the reference executes.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import chains as ochains
from oracle import msm as omsm

pytestmark = pytest.mark.gpu


class Duck:
    def __init__(self, T):
        self.obsp = {"T_forward": T}
        self.shape = (T.shape[0], 0)
        self.uns = {}


def float64_tpm(n=3000, k=14, seed=0):
    """Row-stochastic float64 matrix of a kNN-like graph whose values are NOT float32-representable."""
    from cy2path_b200.synth import knn_velocity_tpm

    g = knn_velocity_tpm(n, k=k, seed=seed, variable_degree=True)
    T = g["T"].astype(np.float64)
    rng = np.random.default_rng(seed + 1)
    T.data = T.data * (1.0 + 0.25 * rng.random(T.nnz))
    T = sp.diags(1.0 / np.asarray(T.sum(1)).ravel()) @ T
    T = sp.csr_matrix(T)
    T.sort_indices()
    assert T.dtype == np.float64 and not np.array_equal(T.data.astype(np.float32).astype(np.float64), T.data)
    return T, g


def test_plan_detects_float64_values():
    from cy2path_b200._lib import Plan

    T, _ = float64_tpm(400)
    assert Plan(T).float64_matrix
    assert not Plan(T.astype(np.float32)).float64_matrix
    assert not Plan(T.astype(np.float32).astype(np.float64)).float64_matrix  # float64 dtype, float32 values


def test_single_start_history_follows_the_float64_product():
    import cy2path_b200 as c2p

    T, g = float64_tpm()
    n = T.shape[0]
    init = g["root_cells"] / g["root_cells"].sum()
    stat = g["end_points"] / g["end_points"].sum()
    iters = 600
    _, ref_hist, ref_conv, ref_eps = omsm.iterate_state_probability(T, init, stat, max_iter=iters)
    _, hist, conv = c2p.iterate_state_probability(Duck(T), init=init, stationary=stat, max_iter=iters)
    assert hist.shape == (iters, n) and hist.dtype == np.float64
    # hi + lo carries 48 of the 53 mantissa bits: 2^-48 relative per product
    assert np.abs(hist - ref_hist).max() <= 1e-12
    assert omsm.rel_l1(hist[-1], ref_hist[-1]) <= 1e-11
    assert conv == ref_conv
    # the float32-rounded matrix would be 4 orders of magnitude away: the test separates the two
    T32 = T.astype(np.float32)
    _, h32, _, _ = omsm.iterate_state_probability(T32, init, stat, max_iter=iters)
    assert omsm.rel_l1(h32[-1], ref_hist[-1]) > 5e-9


def test_single_start_eps_is_bit_reproducible_and_matches_scipy_cosine():
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan

    T, g = float64_tpm(2500, seed=4)
    init = g["root_cells"] / g["root_cells"].sum()
    stat = g["end_points"] / g["end_points"].sum()
    iters = 200
    ref_eps = omsm.iterate_state_probability(T, init, stat, max_iter=iters)[3]
    plan = get_plan(T)
    x0 = torch.from_numpy(init).cuda()
    a = _lib.propagate(plan, x0, iters, stationary=stat, want_final=True, history_stride=0, want_eps=True)
    b = _lib.propagate(plan, x0, iters, stationary=stat, want_final=True, history_stride=0, want_eps=True)
    ea, eb = a["eps"].cpu().numpy().ravel(), b["eps"].cpu().numpy().ravel()
    assert np.array_equal(ea.view(np.uint64), eb.view(np.uint64))
    assert np.abs(ea - ref_eps).max() <= 1e-12
    assert np.array_equal(a["final"].cpu().numpy(), b["final"].cpu().numpy())


@pytest.mark.parametrize("precision,bar", [("f48", 2e-9), ("f40", 1e-6)])
def test_batched_long_run_on_a_float64_matrix(precision, bar):
    import cy2path_b200 as c2p
    from cy2path_b200.synth import batched_starts

    T, g = float64_tpm(2000, seed=2)
    B, iters = 64, 2000
    X0 = batched_starts(T.astype(np.float32), B, seed=9)
    stat = g["end_points"] / g["end_points"].sum()
    ref = omsm.propagate_batch(T, X0, iters)
    out = c2p.propagate_batch(T, X0, iters, stationary=stat, precision=precision)
    got = out["final"].astype(np.float64)
    rel = (np.abs(got - ref).sum(0) / np.abs(ref).sum(0)).max()
    # float32 output rounds once at the end (6e-8 relative per entry, ~3e-8 relative L1)
    assert np.abs(got - ref).max() <= 1e-5 and rel <= max(bar, 6e-8)
    import torch

    out64 = c2p.propagate_batch(T, X0, iters, precision=precision, out_dtype=torch.float64)
    rel64 = (np.abs(out64["final"] - ref).sum(0) / np.abs(ref).sum(0)).max()
    assert rel64 <= bar
    # and the rounded matrix is visibly elsewhere (2e-9 per update): the lo halves are doing the work
    ref32 = omsm.propagate_batch(T.astype(np.float32), X0, iters)
    assert (np.abs(ref32 - ref).sum(0) / np.abs(ref).sum(0)).max() > 1e-7


def test_cdf_tables_and_chains_follow_the_float64_cumsum():
    import cy2path_b200 as c2p

    T, _ = float64_tpm(1500, seed=7)
    ad = Duck(T)
    idx_ref, cum_ref = ochains.extract_nonzero_entries(T)
    idx, cum = c2p.extract_nonzero_entries(ad)
    assert idx.dtype == np.int32 and cum.dtype == np.float32
    assert np.array_equal(idx, idx_ref)
    assert np.array_equal(cum.view(np.uint32), cum_ref.view(np.uint32))
    # (the 3-decimal rounding hides almost every difference between the float64 and the float32 cumsum: on this graph
    # the tables of the rounded matrix are the same words; the check above pins the arithmetic order, not a rare flip)
    rng = np.random.default_rng(11)
    steps, C = 300, 24
    uniforms = rng.random((C, steps)) * 0.998
    init_states = rng.integers(0, T.shape[0], C).astype(np.int32)
    for c in range(0, C, 5):
        s_ref, p_ref = ochains.iterate_markov_chain(T, idx_ref, cum_ref, init_states[c], uniforms[c])
        s, p = c2p.iterate_markov_chain(ad, max_iter=steps + 1, init_state=init_states[c], nonzero_entries=(idx, cum),
                                        uniforms=uniforms[c])
        assert np.array_equal(s, s_ref)
        assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32))
