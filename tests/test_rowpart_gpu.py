"""GPU tests of the row-partitioned propagation (cy2path_b200/rowpart.py, BASELINE config 5) on ONE device: the ranks
are virtual — each has its own sliced plan (its rows of T^T only), its own full-size iterate buffers (NaN-poisoned so a
missing push would surface) and the same `cy2p_propagate_rows_push` kernel; the "peers" are buffers on the same GPU.
Parity is against the float64 oracle (scipy `x @ T`, sampling.py:40-42), including the config-5 size n = 10^6."""
import ctypes

import numpy as np
import pytest

from oracle import fast as ofast

pytestmark = pytest.mark.gpu


def run_virtual(T, world, X0, iters, dtype, sweeps=16, order=None):
    """All `world` virtual ranks on the current device; returns (final iterate in caller order, partitions)."""
    import torch

    from cy2path_b200 import _lib, rowpart

    lib = _lib.load()
    n, B = T.shape[0], X0.shape[1]
    base = rowpart.host_order(T) if order is None else order
    parts = [rowpart.RowPartition(T, world, r, order=base, sweeps=sweeps) for r in range(world)]
    rp0 = parts[0]
    dev = torch.device("cuda")
    order_t = torch.from_numpy(rp0.order.astype(np.int64)).to(dev)
    masks = torch.from_numpy(rp0.masks.view(np.int32)).to(dev)
    X = torch.from_numpy(np.ascontiguousarray(X0)).to(dev).to(dtype)
    bufs = [torch.full((2, n, B), float("nan"), dtype=dtype, device=dev) for _ in range(world)]
    for b in bufs:
        b[0].copy_(X.index_select(0, order_t))
    fn = lib.cy2p_propagate_rows_push_f64 if dtype == torch.float64 else lib.cy2p_propagate_rows_push
    for it in range(iters):
        cur, nxt = it & 1, (it + 1) & 1
        ptrs = torch.tensor([b[nxt].data_ptr() for b in bufs], dtype=torch.int64, device=dev)
        for b in bufs:
            b[nxt].fill_(float("nan"))
        for r, rp in enumerate(parts):
            _lib.check(fn(rp.plan.handle, _lib.ptr(bufs[r][cur]), ctypes.c_void_p(ptrs.data_ptr()), world,
                          _lib.ptr(masks), B, rp.q0, rp.q1, _lib.stream_ptr(dev)))
        torch.cuda.synchronize()
    fin = iters & 1
    out_p = torch.cat([bufs[r][fin][rp.q0:rp.q1] for r, rp in enumerate(parts)])
    out = torch.empty_like(out_p)
    out.index_copy_(0, order_t, out_p)
    return out, parts, bufs[0][fin]


@pytest.mark.parametrize("world,B", [(1, 1), (3, 1), (3, 8), (8, 64)])
def test_virtual_ranks_match_oracle_float64(world, B):
    import torch

    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(20_000, k=15, seed=41)
    T = g["T"]
    X0 = batched_starts(T, B, seed=6)
    out, parts, buf0 = run_virtual(T, world, X0, 30, torch.float64)
    ref = ofast.propagate_batch(T, X0, 30)
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 1e-12
    assert (np.abs(got - ref).sum(0) / np.abs(ref).sum(0)).max() <= 1e-12
    # genuinely partitioned: the ranks' plans hold disjoint slices that add up to T
    assert sum(rp.nnz_local for rp in parts) == T.nnz
    assert all(rp.plan.nnz == rp.nnz_local for rp in parts)
    if world > 1:
        assert max(rp.plan.nnz for rp in parts) < 0.7 * T.nnz
        assert bool(torch.isnan(buf0).any())  # rows rank 0 does not gather were never pushed to it


def test_float32_rows_and_public_driver_world1():
    import torch

    from cy2path_b200 import _lib, rowpart
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    g = knn_velocity_tpm(6001, k=12, seed=43)  # a row count that is not a multiple of the rows a warp covers
    T = g["T"]
    X0 = batched_starts(T, 16, seed=2)
    ref = ofast.propagate_batch(T, X0, 12)
    out, _, _ = run_virtual(T, 4, X0, 12, torch.float32)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-6
    rp = rowpart.RowPartition(T, 1, 0)
    for dt, tol in ((torch.float32, 1e-6), (torch.float64, 1e-13)):
        got = rowpart.propagate(rp, torch.from_numpy(X0).cuda().to(dt), 12)
        assert np.abs(got.cpu().numpy() - ref).max() <= tol
    # B = 1: the opt-in persistent kernel (all updates in one cooperative launch), then the default CUDA-graph path (two updates per
    # replay) for the even part with the eager path for the odd rest
    import os

    x1 = np.ascontiguousarray(X0[:, :1])
    for persist in ("1", "0"):
        os.environ["CY2P_ROWPART_PERSIST"] = persist
        try:
            for iters in (13, 12, 3, 1):
                ref1 = ofast.propagate_batch(T, x1, iters)
                got = rowpart.propagate(rp, torch.from_numpy(x1).cuda().double(), iters)
                assert np.abs(got.cpu().numpy() - ref1).max() <= 1e-13
                got32 = rowpart.propagate(rp, torch.from_numpy(x1).cuda(), iters)
                assert np.abs(got32.cpu().numpy() - ref1).max() <= 1e-6
        finally:
            del os.environ["CY2P_ROWPART_PERSIST"]


def test_config5_size_one_million_cells_against_scipy():
    """n = 10^6, k = 30 (3e7 entries), 4 virtual ranks, B = 1 (the reference's own case) and 20 updates, float64:
    parity with scipy's `x @ T` in float64 — the reference's arithmetic (sampling.py:40-42)."""
    import torch

    from cy2path_b200.synth import knn_velocity_tpm

    g = knn_velocity_tpm(1_000_000, k=30, seed=5)
    T = g["T"]
    x0 = (g["root_cells"] / g["root_cells"].sum()).reshape(-1, 1)
    out, parts, _ = run_virtual(T, 4, x0.astype(np.float32), 20, torch.float64, sweeps=4)
    x = x0.astype(np.float32).astype(np.float64).ravel()
    for _ in range(20):
        x = x @ T
    got = out.cpu().numpy().ravel()
    assert np.abs(got - x).max() <= 1e-12
    assert np.abs(got - x).sum() / np.abs(x).sum() <= 1e-12
    assert max(rp.plan.nnz for rp in parts) < 0.4 * T.nnz
