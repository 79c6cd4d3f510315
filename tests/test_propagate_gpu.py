"""GPU parity tests for K1 (MSM propagation) — CUDA path through the C ABI vs oracle and golden vectors.

Parity bar (BASELINE north_star): max-abs <= 1e-5 and relative L1 <= 1e-6 on every stored iterate,
identical convergence index.
"""
import numpy as np
import pandas as pd
import pytest

from oracle import fast as ofast
from oracle import msm as omsm

pytestmark = pytest.mark.gpu

MAX_ABS, REL_L1 = 1e-5, 1e-6


class Duck:
    def __init__(self, T, root=None, end=None):
        self.obsp = {"T_forward": T}
        self.shape = (T.shape[0], 0)
        self.obs = pd.DataFrame({"root_cells": root, "end_points": end}) if root is not None else pd.DataFrame()
        self.uns = {}

    def copy(self):
        import copy

        return copy.deepcopy(self)


def _assert_close(a, b):
    assert np.abs(np.asarray(a, dtype=np.float64) - b).max() <= MAX_ABS
    assert omsm.rel_l1(a, b) <= REL_L1


@pytest.mark.parametrize("name", ["msm_small.npz", "msm_vardeg.npz", "msm_f64matrix.npz"])
def test_history_matches_reference_golden(golden, name):
    import cy2path_b200 as c2p

    g = golden(name)
    ad = Duck(g["T"])
    sh, shm, conv = c2p.iterate_state_probability(ad, init=g["init"], stationary=g["stationary"],
                                                  max_iter=g["history"].shape[0], tol=float(g["tol"]))
    assert shm.dtype == np.float64 and shm.shape == g["history"].shape
    assert conv == int(g["convergence"])
    assert sh.shape[0] == conv
    for i in range(shm.shape[0]):
        _assert_close(shm[i], g["history"][i])
    assert np.array_equal(shm[0], g["init"])
    # the single-start path iterates in float64 on the device: far inside the bar
    assert np.abs(shm - g["history"]).max() <= 1e-12


def test_config1_shape_full_history_vs_oracle():
    """BASELINE config 1: 3,696 cells, k=30, 1,000 iterations from the root prior."""
    import cy2path_b200 as c2p
    from cy2path_b200.synth import knn_velocity_tpm

    g = knn_velocity_tpm(3696, k=30, seed=1)
    init = g["root_cells"] / g["root_cells"].sum()
    stat = g["end_points"] / g["end_points"].sum()
    _, ref_hist, ref_conv, ref_eps = omsm.iterate_state_probability(g["T"], init, stat, max_iter=1000)
    ad = Duck(g["T"])
    _, hist, conv = c2p.iterate_state_probability(ad, init=init, stationary=stat, max_iter=1000)
    assert conv == ref_conv
    assert np.abs(hist - ref_hist).max() <= MAX_ABS
    worst = max(omsm.rel_l1(hist[i], ref_hist[i]) for i in range(0, 1000, 7))
    assert worst <= REL_L1
    # the float64 single-start path tracks the reference far inside the bar, including the slow growth of the
    # total mass (float32-rounded rows of T do not sum to exactly 1; the reference's float64 state keeps that)
    assert np.abs(hist - ref_hist).max() <= 1e-12
    assert np.abs(hist.sum(1) - ref_hist.sum(1)).max() <= 1e-12 and hist.min() >= 0


@pytest.mark.parametrize("B,iters", [(1, 5), (3, 4), (4, 6), (32, 9), (128, 7), (132, 5), (384, 3)])
def test_batched_matches_oracle_and_equals_looped(B, iters):
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(700, k=12, seed=5, variable_degree=(B % 2 == 1))
    T = g["T"]
    X0 = batched_starts(T, B, seed=3)
    stat = g["end_points"] / g["end_points"].sum()
    ref = ofast.propagate_batch(T, X0, iters)
    plan = get_plan(T)
    out = _lib.propagate(plan, torch.from_numpy(X0).cuda(), iters, stationary=stat, want_eps=True)
    got = out["final"].cpu().numpy()
    _assert_close(got, ref)
    # eps against the oracle's definition, float64 dots
    Xd = np.asarray(X0, dtype=np.float64)
    eps_ref = np.empty((iters, B))
    for i in range(iters):
        Xd = ofast.propagate_batch(T, Xd, 1)
        eps_ref[i] = [omsm.cosine_distance(Xd[:, b], stat) for b in range(B)]
    assert np.abs(out["eps"].cpu().numpy() - eps_ref).max() <= 1e-6
    # batched == looped single starts, bit for bit (same kernel arithmetic per column family)
    if B <= 4:
        for b in range(B):
            one = _lib.propagate(plan, torch.from_numpy(np.ascontiguousarray(X0[:, b])).cuda(), iters)
            assert np.abs(one["final"].cpu().numpy() - got[:, b]).max() <= 1e-7


def test_long_run_precision_fp32_vs_fp64_iterate():
    """1,000 iterations, 8 starts on the config-1 graph. A float32 iterate cannot follow the reference's float64
    iterate to 1e-6 relative L1 over this many steps (the float32-rounded rows of T sum to 1 +- 1e-8 and the
    reference's float64 state accumulates that growth, which float32 storage rounds away): it is held to the
    max-abs bar and to a stated 1e-4 relative L1; the float64 iterate is held to the full north_star bar."""
    import torch

    from cy2path_b200.sampling import propagate_batch
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(3696, k=30, seed=1)
    T = g["T"]
    X0 = batched_starts(T, 8, seed=4)
    ref = ofast.propagate_batch(T, X0, 1000)
    r32 = propagate_batch(T, X0, 1000)["final"]
    r64 = propagate_batch(T, X0, 1000, precision="fp64")["final"]
    assert r32.dtype == np.float32 and r64.dtype == np.float64
    assert np.abs(r64 - ref).max() <= 1e-12 and omsm.rel_l1(r64, ref) <= 1e-12
    assert np.abs(r32 - ref).max() <= MAX_ABS
    assert omsm.rel_l1(r32, ref) <= 1e-4
    # short horizons: the float32 iterate meets the full bar
    ref60 = ofast.propagate_batch(T, X0, 60)
    r32s = propagate_batch(T, X0, 60)["final"]
    _assert_close(r32s, ref60)


def test_locality_ordering_path_matches_oracle():
    """Batched runs long enough to amortise it switch to the plan's Cuthill-McKee cell order (reorder.cu): the
    first update reads the caller's order, the middle ones run permuted -> permuted, history slots and the final
    iterate are written back in the caller's order."""
    import ctypes

    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(2000, k=12, seed=21)
    T = g["T"]
    plan = _lib.Plan(T)
    info = (ctypes.c_int64 * 8)()
    X0 = batched_starts(T, 64, seed=9)
    stat = g["end_points"] / g["end_points"].sum()
    iters = 70
    out = _lib.propagate(plan, torch.from_numpy(X0).cuda(), iters, stationary=stat, want_eps=True, history_stride=16)
    _lib.check(_lib.load().cy2p_plan_info(plan.handle, info))
    assert info[6] == 1, "the synthetic cell order is random: the ordering must have been accepted"
    ref = np.asarray(X0, dtype=np.float64)
    eps_ref = np.empty((iters, 64))
    hist_ref = []
    for it in range(iters):
        if it % 16 == 0:
            hist_ref.append(ref.copy())
        ref = ofast.propagate_batch(T, ref, 1)
        eps_ref[it] = [omsm.cosine_distance(ref[:, b], stat) for b in range(64)]
    h = out["history"].cpu().numpy()
    assert h.shape == (len(hist_ref), 2000, 64)
    for a, b in zip(h, hist_ref):
        assert np.abs(a - b).max() <= MAX_ABS and omsm.rel_l1(a, b) <= 3e-6  # 64 iterations of fp32 rounding
    _assert_close(h[0], hist_ref[0])
    got = out["final"].cpu().numpy()
    assert np.abs(got - ref).max() <= MAX_ABS and omsm.rel_l1(got, ref) <= 3e-6
    assert np.abs(out["eps"].cpu().numpy() - eps_ref).max() <= 1e-6
    # same answer (to fp32 summation order) as the caller-order path of a fresh plan with the ordering disabled
    import os

    os.environ["CY2P_REORDER"] = "0"
    try:
        plan2 = _lib.Plan(T)
        out2 = _lib.propagate(plan2, torch.from_numpy(X0).cuda(), iters)
        _lib.check(_lib.load().cy2p_plan_info(plan2.handle, info))
        assert info[6] == -1
    finally:
        del os.environ["CY2P_REORDER"]
    assert np.abs(out2["final"].cpu().numpy() - got).max() <= 1e-6


def test_history_stride_and_linearity():
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(500, k=8, seed=6)
    plan = get_plan(g["T"])
    X0 = torch.from_numpy(batched_starts(g["T"], 8, seed=1)).cuda()
    out = _lib.propagate(plan, X0, 10, history_stride=3)
    h = out["history"].cpu().numpy()  # iterates before updates 0, 3, 6, 9
    assert h.shape == (4, 500, 8)
    ref = np.asarray(X0.cpu().numpy(), dtype=np.float64)
    for it in range(10):
        if it % 3 == 0:
            _assert_close(h[it // 3], ref)
        ref = ofast.propagate_batch(g["T"], ref, 1)
    _assert_close(out["final"].cpu().numpy(), ref)
    # linearity: propagate(a*x + b*y) == a*propagate(x) + b*propagate(y)
    a, b = 0.3, 0.7
    mix = (a * X0[:, :4] + b * X0[:, 4:]).contiguous()
    pm = _lib.propagate(plan, mix, 6)["final"].cpu().numpy()
    ps = _lib.propagate(plan, X0, 6)["final"].cpu().numpy()
    assert np.abs(pm - (a * ps[:, :4] + b * ps[:, 4:])).max() <= 1e-6


def test_row_partition_equals_full_update():
    """Virtual ranks on one GPU: the row-partitioned update (config 5) tiles to the full update."""
    import ctypes

    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(900, k=10, seed=7)
    plan = get_plan(g["T"])
    for B in (1, 8):
        X = torch.from_numpy(batched_starts(g["T"], B, seed=2)).cuda()
        full = _lib.propagate(plan, X, 1)["final"]
        Y = torch.zeros_like(X)
        bounds = [0, 200, 455, 900]
        for r0, r1 in zip(bounds[:-1], bounds[1:]):
            _lib.check(_lib.load().cy2p_propagate_rows(plan.handle, _lib.ptr(X), _lib.ptr(Y), B, r0, r1,
                                                       _lib.stream_ptr()))
        if B == 1:  # B == 1 full update uses the lane-parallel row sum: same values up to fp32 rounding
            assert torch.allclose(Y, full, rtol=0, atol=1e-7)
        else:
            assert torch.equal(Y, full)


def test_row_partition_in_locality_order():
    """Config-5 driver with the iterate kept in the plan's locality order: virtual ranks == unsharded run."""
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200 import dist as c2dist
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(1500, k=10, seed=23)
    plan = _lib.Plan(g["T"])
    order = c2dist.plan_order(plan)
    assert sorted(order.cpu().tolist()) == list(range(1500))
    upd = c2dist.cuda_row_update_ordered(plan)
    for B in (1, 32):
        X = torch.from_numpy(batched_starts(g["T"], B, seed=3)).cuda()
        ref = _lib.propagate(plan, X, 6)["final"]
        got = c2dist.propagate_row_partitioned_cuda(plan, X, 6, ordered=True)  # world size 1
        assert torch.allclose(got, ref, rtol=0, atol=2e-7)
        # three virtual ranks over the permuted rows tile to the same update
        Xp = X.index_select(0, order)
        Yp = torch.zeros_like(Xp)
        for r0, r1 in ((0, 400), (400, 1100), (1100, 1500)):
            upd(Xp, Yp, r0, r1)
        one = torch.empty_like(Yp)
        one.index_copy_(0, order, Yp)
        assert torch.allclose(one, _lib.propagate(plan, X, 1)["final"], rtol=0, atol=2e-7)


def test_bad_csr_and_workspace_errors():
    import scipy.sparse as sp

    from cy2path_b200 import _lib

    T = sp.csr_matrix(np.eye(4, dtype=np.float32))
    T.indices = T.indices.copy()
    T.indices[2] = 9
    with pytest.raises(ValueError):
        _lib.Plan(T)


def test_drop_in_wrapper_keys():
    import cy2path_b200 as c2p
    from cy2path_b200.synth import knn_velocity_tpm

    g = knn_velocity_tpm(600, k=10, seed=8)
    ad = Duck(g["T"], g["root_cells"], g["end_points"])
    c2p.sample_state_probability(ad, max_iter=50)
    u = ad.uns["state_probability_sampling"]
    assert set(u) == {"state_history", "state_history_max_iter", "init_state_probability",
                      "stationary_state_probability", "sampling_params"}
    assert set(u["sampling_params"]) == {"recalc_matrix", "self_transitions", "init_type", "max_iter", "convergence",
                                         "tol", "copy"}
    assert u["state_history_max_iter"].shape == (50, 600)


def test_edge_cases_tiny_graphs_empty_rows_and_short_runs():
    """Cells without out-edges or in-edges, a 1-cell graph, max_iter of 1 and 0 updates."""
    import scipy.sparse as sp
    import torch

    import cy2path_b200 as c2p
    from cy2path_b200 import _lib

    # 6 cells: cell 2 has no out-edges (empty CSR row), cell 4 has no in-edges
    rows = np.array([0, 0, 1, 1, 3, 3, 4, 5, 5])
    cols = np.array([1, 2, 0, 3, 2, 5, 0, 3, 5])
    vals = np.array([0.5, 0.5, 0.25, 0.75, 0.4, 0.6, 1.0, 0.3, 0.7], dtype=np.float32)
    T = sp.csr_matrix((vals, (rows, cols)), shape=(6, 6))
    T.indices = T.indices.astype(np.int32)
    T.indptr = T.indptr.astype(np.int32)
    init = np.array([0.5, 0.0, 0.0, 0.0, 0.5, 0.0])
    stat = np.full(6, 1 / 6)
    ad = Duck(T)
    _, hist, conv = c2p.iterate_state_probability(ad, init=init, stationary=stat, max_iter=12)
    _, ref, rconv, _ = omsm.iterate_state_probability(T, init, stat, max_iter=12)
    assert conv == rconv and np.abs(hist - ref).max() <= 1e-14
    # batched, every small B, fp32 and fp64
    plan = _lib.Plan(T)
    for B in (1, 2, 5, 8):
        X0 = np.random.default_rng(B).random((6, B)).astype(np.float32)
        X0 /= X0.sum(0)
        refB = ofast.propagate_batch(T, X0, 7)
        for dt in (torch.float32, torch.float64):
            got = _lib.propagate(plan, torch.from_numpy(X0).to(dt).cuda(), 7)["final"].cpu().numpy()
            assert np.abs(got - refB).max() <= (1e-6 if dt == torch.float32 else 1e-14)
    # zero updates returns the start; max_iter == 1 keeps only the start row
    x = torch.from_numpy(X0).cuda()
    assert torch.equal(_lib.propagate(plan, x, 0)["final"], x)
    _, h1, c1 = c2p.iterate_state_probability(ad, init=init, stationary=stat, max_iter=1)
    assert h1.shape == (1, 6) and c1 == 1 and np.array_equal(h1[0], init)
    # a single absorbing cell
    T1 = sp.csr_matrix(np.ones((1, 1), dtype=np.float32))
    _, h, c = c2p.iterate_state_probability(Duck(T1), init=np.ones(1), stationary=np.ones(1), max_iter=5)
    assert np.array_equal(h, np.ones((5, 1)))
    # float64 user matrix and int64 indices are accepted (cast to the device layout)
    T64 = sp.csr_matrix(T, dtype=np.float64)
    T64.indices = T64.indices.astype(np.int64)
    T64.indptr = T64.indptr.astype(np.int64)
    _, h64, _ = c2p.iterate_state_probability(Duck(T64), init=init, stationary=stat, max_iter=12)
    assert np.abs(h64 - ref).max() <= 1e-14


def test_batched_convergence_cut_matches_reference_rule(golden):
    """propagate_batch's per-start convergence (computed on the device) against the reference's rule applied
    to each column of the returned eps history (sampling.py:13-25, 48-57), and against the oracle's eps."""
    from cy2path_b200.sampling import check_convergence_criteria, propagate_batch

    g = golden("msm_small.npz")
    T, n = g["T"], g["T"].shape[0]
    rng = np.random.default_rng(8)
    B, iters, tol = 12, 120, float(g["tol"])
    X0 = np.zeros((n, B), dtype=np.float32)
    X0[:, 0] = (g["init"] / g["init"].sum()).astype(np.float32)
    for b in range(1, B):
        X0[rng.choice(n, 5, replace=False), b] = 0.2
    r = propagate_batch(T, X0, iters, stationary=g["stationary"], tol=tol, precision="fp64")
    assert r["eps"].shape == (iters, B) and r["convergence"].shape == (B,) and r["convergence"].dtype == np.int64
    for b in range(B):
        c = check_convergence_criteria(r["eps"][:, b], tol)
        assert r["convergence"][b] == (c if (c is not None and c < iters) else iters)
    assert r["convergence"][0] == int(g["convergence"])  # column 0 is the golden run's start
    assert (r["convergence"] < iters).any()


def test_staged_gather_experiment_matches_plain(monkeypatch):
    """The opt-in shared-memory staged gather (CY2P_STAGED=1; slower than the default, kept as a measured
    experiment) must produce the same iterates and eps as the default gather."""
    from cy2path_b200.sampling import clear_plan_cache, propagate_batch
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(6000, k=20, seed=17)
    T = g["T"]
    X0 = batched_starts(T, 128, seed=2)
    stat = stationary_by_power(T, 50)
    clear_plan_cache()
    ref = propagate_batch(T, X0, 40, stationary=stat)
    for prec in ("fp32", "fp64"):
        monkeypatch.setenv("CY2P_STAGED", "1")
        clear_plan_cache()  # the block lists are built with the plan's ordering
        got = propagate_batch(T, X0, 40, stationary=stat, precision=prec)
        monkeypatch.setenv("CY2P_STAGED", "0")
        assert np.abs(got["final"] - ref["final"]).max() <= 2e-7
        assert np.abs(got["eps"] - ref["eps"]).max() <= 1e-6
    clear_plan_cache()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_halo_push_with_virtual_ranks_on_one_gpu(dtype):
    """cy2p_propagate_rows_push / cy2p_propagate_rows_allgather with the 'peers' being separate buffers on one
    device: every virtual rank keeps a full-size iterate, computes its own rows and stores a row only where the
    per-row mask says it is needed (buffers are poisoned with NaN first: a missing push would surface). The assembled
    result equals the unsharded propagation; world size 1 of the drivers equals it too."""
    import ctypes

    import torch

    from cy2path_b200 import _lib
    from cy2path_b200 import dist as c2dist
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    lib = _lib.load()
    n, iters, world = 4000, 5, 3
    g = knn_velocity_tpm(n, k=12, seed=29)
    plan = _lib.Plan(g["T"])
    tdt = torch.float32 if dtype == "float32" else torch.float64
    for B in (1, 64):
        X = torch.from_numpy(batched_starts(g["T"], B, seed=4)).cuda().to(tdt)
        ref = _lib.propagate(plan, X, iters)["final"]
        tol = 2e-7 if dtype == "float32" else 1e-14
        assert torch.allclose(c2dist.propagate_row_partitioned_halo(plan, g["T"], X, iters), ref, rtol=0, atol=tol)
        assert torch.allclose(c2dist.propagate_row_partitioned_fused(plan, X, iters), ref, rtol=0, atol=tol)
        order = c2dist.plan_order(plan)
        masks_np = c2dist.row_peer_masks(g["T"], order.cpu().numpy(), world)
        masks = torch.from_numpy(masks_np.view("int32")).cuda()
        m, blocks = c2dist.row_blocks(n, world)
        bufs = [torch.full((2, n, B), float("nan"), dtype=tdt, device="cuda") for _ in range(world)]
        for b in bufs:
            b[0].copy_(X.index_select(0, order))
        fn = lib.cy2p_propagate_rows_push_f64 if dtype == "float64" else lib.cy2p_propagate_rows_push
        for it in range(iters):
            cur, nxt = it & 1, (it + 1) & 1
            ptrs = torch.tensor([b[nxt].data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
            for b in bufs:
                b[nxt].fill_(float("nan"))
            for r, (r0, r1) in enumerate(blocks):
                _lib.check(fn(plan.handle, _lib.ptr(bufs[r][cur]), ctypes.c_void_p(ptrs.data_ptr()), world,
                              _lib.ptr(masks), B, r0, r1, _lib.stream_ptr(X.device)))
            torch.cuda.synchronize()
        fin = iters & 1
        out_p = torch.cat([bufs[r][fin][r0:r1] for r, (r0, r1) in enumerate(blocks)])
        assert torch.isfinite(out_p).all()
        out = torch.empty_like(out_p)
        out.index_copy_(0, order, out_p)
        assert torch.allclose(out, ref, rtol=0, atol=tol)
        # the masks push fewer rows than a full exchange, and un-needed rows really stay untouched (NaN)
        pushed = sum(int(((masks_np >> np.uint32(p)) & np.uint32(1)).sum()) for p in range(world))
        assert pushed < world * n
        assert torch.isnan(bufs[0][fin]).any()
