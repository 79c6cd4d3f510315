"""GPU parity tests for the extended-precision batched propagation (csrc/propagate_x.cu, `cy2p_propagate_x`) —
the default of `propagate_batch` and the path bench.py's headline runs.

Parity bar (BASELINE north_star): max-abs <= 1e-5 and relative L1 <= 1e-6 against the reference's float64 iterate
(`sampling.py:33,40-42`, restated in oracle/), at the iteration counts the configs name (1,000-2,000); per-column
convergence cut identical to `check_convergence_criteria` on the oracle's float64 eps (`sampling.py:13-25`).
"""
import numpy as np
import pytest

from oracle import fast as ofast
from oracle import msm as omsm

pytestmark = pytest.mark.gpu

MAX_ABS, REL_L1 = 1e-5, 1e-6


def _rel_l1_cols(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b).sum(0) / np.abs(b).sum(0)).max())


def _oracle_run(T, X0, iters, stat=None):
    """float64 oracle, step by step; returns (final, eps[iters, B] or None)."""
    X = np.asarray(X0, dtype=np.float64)
    eps = np.empty((iters, X.shape[1])) if stat is not None else None
    for i in range(iters):
        X = ofast.propagate_batch(T, X, 1)
        if stat is not None:
            uv = X.T @ stat
            eps[i] = np.clip(1.0 - uv / np.sqrt((X * X).sum(0) * (stat @ stat)), 0.0, 2.0)
    return X, eps


@pytest.mark.parametrize("B,iters,tail", [(1, 5, 16), (3, 4, 16), (4, 1, 16), (32, 9, 8), (128, 7, 16), (132, 5, 8),
                                           (384, 3, 16), (130, 2, 16)])
def test_small_cases_match_oracle(B, iters, tail):
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(700, k=12, seed=5, variable_degree=(B % 2 == 1))
    T = g["T"]
    X0 = batched_starts(T, B, seed=3)
    stat = g["end_points"] / g["end_points"].sum()
    ref, eps_ref = _oracle_run(T, X0, iters, stat)
    plan = get_plan(T)
    x0 = torch.from_numpy(X0).cuda()
    out = _lib.propagate_x(plan, x0, iters, stationary=stat, tail_bits=tail, want_eps=True)
    got = out["final"].cpu().numpy()
    assert got.dtype == np.float32
    assert np.abs(got - ref).max() <= MAX_ABS and _rel_l1_cols(got, ref) <= REL_L1
    assert np.abs(out["eps"].cpu().numpy() - eps_ref).max() <= (1e-10 if tail == 16 else 1e-7)
    # float64 output: only the format's own rounding is left
    out64 = _lib.propagate_x(plan, x0, iters, tail_bits=tail, out_dtype=torch.float64)
    g64 = out64["final"].cpu().numpy()
    assert g64.dtype == np.float64
    assert _rel_l1_cols(g64, ref) <= (iters * 2e-11 if tail == 16 else iters * 5e-9)


@pytest.mark.parametrize("cpl,warps,minb,u,persist", [(4, 8, 2, 2, 0), (4, 8, 3, 4, 1), (4, 16, 1, 4, 0), (8, 8, 3, 2, 0),
                                                      (8, 16, 1, 2, 1), (8, 8, 2, 4, 1)])
def test_kernel_shape_variants_agree(monkeypatch, cpl, warps, minb, u, persist):
    """The tuning variants (columns per lane, warps per CTA) walk the same rows in the same order: identical bits."""
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(3000, k=15, seed=11)
    T = g["T"]
    plan = _lib.Plan(T)
    X0 = torch.from_numpy(batched_starts(T, 200, seed=1)).cuda()
    stat = stationary_by_power(T, 30)
    base = _lib.propagate_x(plan, X0, 30, stationary=stat, want_eps=True)
    monkeypatch.setenv("CY2PX_CPL", str(cpl))
    monkeypatch.setenv("CY2PX_WARPS", str(warps))
    monkeypatch.setenv("CY2PX_MINB", str(minb))
    monkeypatch.setenv("CY2PX_U", str(u))
    monkeypatch.setenv("CY2PX_PERSIST", str(persist))
    for tail in (16, 8):
        ref = base if tail == 16 else None
        got = _lib.propagate_x(plan, X0, 30, stationary=stat, want_eps=True, tail_bits=tail)
        if ref is not None:
            assert torch.equal(got["final"], ref["final"])
            # eps: the per-thread partial sums cover different row sets per shape, the integer sums differ in the
            # last bits of the fixed-point grid only
            assert float((got["eps"] - ref["eps"]).abs().max()) <= 1e-13
        else:
            ref8, _ = _oracle_run(T, X0.cpu().numpy(), 30)
            assert _rel_l1_cols(got["final"].cpu().numpy(), ref8) <= REL_L1


def test_multi_iteration_persistent_variant_agrees(monkeypatch):
    """CY2PX_MULTI=1: all middle updates of a column tile in one cooperative launch (grid barrier per update)."""
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(3000, k=15, seed=11)
    T = g["T"]
    plan = _lib.Plan(T)
    X0 = torch.from_numpy(batched_starts(T, 300, seed=1)).cuda()
    stat = stationary_by_power(T, 30)
    for iters in (30, 31, 4):
        base = _lib.propagate_x(plan, X0, iters, stationary=stat, want_eps=True)
        monkeypatch.setenv("CY2PX_MULTI", "1")
        monkeypatch.setenv("CY2PX_CPL", "4")
        monkeypatch.setenv("CY2P_TILE", "128")
        got = _lib.propagate_x(plan, X0, iters, stationary=stat, want_eps=True)
        monkeypatch.delenv("CY2PX_MULTI")
        monkeypatch.delenv("CY2PX_CPL")
        monkeypatch.delenv("CY2P_TILE")
        assert torch.equal(got["final"], base["final"])
        assert float((got["eps"] - base["eps"]).abs().max()) <= 1e-13


def test_long_run_2000_iterations_all_formats():
    """2,000 updates of 8 starts on the config-1 graph (the horizon of BASELINE configs 2 / 4)."""
    import torch

    from cy2path_b200.sampling import propagate_batch
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(3696, k=30, seed=1)
    T = g["T"]
    X0 = batched_starts(T, 8, seed=4)
    ref, _ = _oracle_run(T, X0, 2000)
    r48 = propagate_batch(T, X0, 2000)["final"]  # the default
    assert r48.dtype == np.float32
    assert np.abs(r48 - ref).max() <= MAX_ABS and _rel_l1_cols(r48, ref) <= REL_L1
    r48d = propagate_batch(T, X0, 2000, out_dtype=torch.float64)["final"]
    assert _rel_l1_cols(r48d, ref) <= 1e-8  # 36 mantissa bits: ~5e-10 measured by emulation (tools/precision_lab.py)
    r40 = propagate_batch(T, X0, 2000, precision="f40", out_dtype=torch.float64)["final"]
    assert np.abs(r40 - ref).max() <= MAX_ABS and _rel_l1_cols(r40, ref) <= REL_L1
    r64 = propagate_batch(T, X0, 2000, precision="fp64")["final"]
    assert np.abs(r64 - ref).max() <= 1e-12 and _rel_l1_cols(r64, ref) <= 1e-12
    # the float32 iterate is OUTSIDE the bar at this horizon (kept as a measured statement, DESIGN.md "precision")
    r32 = propagate_batch(T, X0, 2000, precision="fp32")["final"]
    assert np.abs(r32 - ref).max() <= 5e-5
    assert 1e-6 < _rel_l1_cols(r32, ref) <= 2e-4


def test_eps_bit_reproducible_and_cut_matches_oracle():
    """256 starts x 300 updates: eps is the same bits on every run (fixed-point integer atomics) and the per-column
    convergence cut is the one the reference's rule gives on the ORACLE's float64 eps (sampling.py:13-25,48-57)."""
    from cy2path_b200.sampling import check_convergence_criteria, clear_plan_cache, propagate_batch
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(3696, k=30, seed=1)
    T = g["T"]
    B, iters, tol = 256, 300, 1e-5
    X0 = batched_starts(T, B, seed=6)
    stat = stationary_by_power(T, 200)
    _, eps_ref = _oracle_run(T, X0, iters, stat)
    runs = []
    for _ in range(3):
        clear_plan_cache()
        runs.append(propagate_batch(T, X0, iters, stationary=stat, tol=tol))
    for r in runs[1:]:
        assert np.array_equal(r["eps"].view(np.uint64), runs[0]["eps"].view(np.uint64))
        assert np.array_equal(r["final"].view(np.uint32), runs[0]["final"].view(np.uint32))
    eps = runs[0]["eps"]
    assert np.abs(eps - eps_ref).max() <= 1e-10
    want = np.empty(B, dtype=np.int64)
    margin = np.empty(B)
    for b in range(B):
        c = check_convergence_criteria(eps_ref[:, b], tol)
        want[b] = c if (c is not None and c < iters) else iters
        margin[b] = np.abs(np.abs(np.diff(eps_ref[:, b])) - tol).min()
    # a column whose |eps[i+1]-eps[i]| comes within 1e-10 of tol could legitimately flip; none does here
    assert margin.min() > 1e-9
    assert np.array_equal(runs[0]["convergence"], want)
    assert (want < iters).any() and (want > 2).any()


def test_odd_shapes_and_unaligned_columns():
    """B not a multiple of the slab or of 4, a tile boundary inside the batch (CY2P_TILE), tiny graphs with empty
    rows, zero updates."""
    import os

    import scipy.sparse as sp
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(900, k=9, seed=12, variable_degree=True)
    T = g["T"]
    plan = _lib.Plan(T)
    os.environ["CY2P_TILE"] = "128"
    try:
        for B in (5, 127, 129, 300):
            X0 = batched_starts(T, B, seed=B)
            ref, _ = _oracle_run(T, X0, 6)
            for tail in (16, 8):
                got = _lib.propagate_x(plan, torch.from_numpy(X0).cuda(), 6, tail_bits=tail)["final"].cpu().numpy()
                assert np.abs(got - ref).max() <= MAX_ABS and _rel_l1_cols(got, ref) <= REL_L1
    finally:
        del os.environ["CY2P_TILE"]
    rows = np.array([0, 0, 1, 1, 3, 3, 4, 5, 5])
    cols = np.array([1, 2, 0, 3, 2, 5, 0, 3, 5])
    vals = np.array([0.5, 0.5, 0.25, 0.75, 0.4, 0.6, 1.0, 0.3, 0.7], dtype=np.float32)
    T6 = sp.csr_matrix((vals, (rows, cols)), shape=(6, 6))
    T6.indices, T6.indptr = T6.indices.astype(np.int32), T6.indptr.astype(np.int32)
    p6 = _lib.Plan(T6)
    for B in (1, 2, 7):
        X0 = np.random.default_rng(B).random((6, B)).astype(np.float32)
        ref, _ = _oracle_run(T6, X0, 9)
        got = _lib.propagate_x(p6, torch.from_numpy(X0).cuda(), 9, out_dtype=torch.float64)["final"].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-9
    x = torch.from_numpy(X0).cuda()
    assert torch.equal(_lib.propagate_x(p6, x, 0)["final"], x)


def test_config2_full_size_4096_starts_2000_iterations():
    """BASELINE config 2 at FULL size through the public API: 50,000 cells, 4,096 starts, 2,000 updates, the shape
    and iteration count bench.py's headline runs. Eight columns (both column tiles, tile edges) against the float64
    oracle at the north_star bar, their eps histories and convergence cuts against the oracle's."""
    from cy2path_b200.sampling import check_convergence_criteria, propagate_batch
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(50_000, k=30, seed=2)
    T = g["T"]
    B, iters, tol = 4096, 2000, 1e-5
    X0 = batched_starts(T, B, seed=2)
    stat = stationary_by_power(T, 50)
    r = propagate_batch(T, X0, iters, stationary=stat, tol=tol)
    X = r["final"]
    assert X.shape == (50_000, B) and X.dtype == np.float32 and r["eps"].shape == (iters, B)
    cols = [0, 3, 1000, 2047, 2048, 3001, 4000, 4095]
    ref, eps_ref = _oracle_run(T, X0[:, cols], iters, stat)
    got = X[:, cols]
    assert np.abs(got - ref).max() <= MAX_ABS
    assert _rel_l1_cols(got, ref) <= REL_L1
    assert np.abs(r["eps"][:, cols] - eps_ref).max() <= 1e-9
    for j, c in enumerate(cols):
        cc = check_convergence_criteria(eps_ref[:, j], tol)
        assert r["convergence"][c] == (cc if (cc is not None and cc < iters) else iters)
    # every start distribution stays a distribution (the reference's float64 mass grows by ~1e-8 per step)
    mass = X.sum(0, dtype=np.float64)
    assert np.abs(mass - 1).max() <= 1e-4 and X.min() >= 0
    assert np.isfinite(r["eps"]).all() and r["eps"].min() >= 0 and r["eps"].max() <= 2


def test_config4_size_250k_cells_1000_iterations():
    """BASELINE config 4's graph at full size (250,000 cells, 7.5e6 transitions), one GPU's share of the strong-scaled
    run at 8 GPUs (512 of the 4,096 starts), all 1,000 updates: four columns against the float64 oracle at the
    north_star bar, with their eps histories and convergence cuts."""
    from cy2path_b200.sampling import check_convergence_criteria, propagate_batch
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    n = 250_000
    g = knn_velocity_tpm(n, k=30, seed=4)
    T = g["T"]
    B, iters, tol = 512, 1000, 1e-5
    X0 = batched_starts(T, B, seed=4)
    stat = stationary_by_power(T, 30)
    r = propagate_batch(T, X0, iters, stationary=stat, tol=tol)
    X = r["final"]
    assert X.shape == (n, B) and X.dtype == np.float32 and r["eps"].shape == (iters, B)
    cols = [0, 255, 256, 511]
    ref, eps_ref = _oracle_run(T, X0[:, cols], iters, stat)
    got = X[:, cols]
    assert np.abs(got - ref).max() <= MAX_ABS
    assert _rel_l1_cols(got, ref) <= REL_L1
    assert np.abs(r["eps"][:, cols] - eps_ref).max() <= 1e-9
    for j, c in enumerate(cols):
        cc = check_convergence_criteria(eps_ref[:, j], tol)
        assert r["convergence"][c] == (cc if (cc is not None and cc < iters) else iters)
    mass = X.sum(0, dtype=np.float64)
    assert np.abs(mass - 1).max() <= 1e-4 and X.min() >= 0
