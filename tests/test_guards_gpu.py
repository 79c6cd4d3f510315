"""Out-of-bounds write hunt without compute-sanitizer (closed on the GPU pool): with CY2P_GUARD=1 the ctypes layer
fences every output and work-space buffer it hands to the library with 64 KB canaries (cy2path_b200/_lib.py
`dev_empty`), and `check_guards()` fails if one canary byte changed. Odd, unaligned and ragged shapes of every entry
point, including the fallback variants of the kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def guarded(monkeypatch):
    from cy2path_b200 import _lib

    monkeypatch.setenv("CY2P_GUARD", "1")
    _lib._GUARDS.clear()
    yield
    _lib._GUARDS.clear()


def _graph(n, k=11, seed=0, variable=True):
    from cy2path_b200.synth import knn_velocity_tpm

    return knn_velocity_tpm(n, k=k, seed=seed, variable_degree=variable)


def test_canary_detects_a_planted_overrun():
    import torch

    from cy2path_b200 import _lib

    t = _lib.dev_empty((10,), torch.float32, torch.device("cuda", 0), "planted")
    assert t.numel() == 10
    flat = _lib._GUARDS[-1][3]
    flat[_lib._GUARD_BYTES + 40] = 7  # first byte after the 40-byte payload
    with pytest.raises(RuntimeError, match="AFTER"):
        _lib.check_guards()


@pytest.mark.parametrize("n,B,iters", [(701, 1, 37), (1033, 3, 5), (1033, 130, 9), (517, 257, 4)])
def test_propagation_entry_points(n, B, iters, monkeypatch):
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts

    g = _graph(n, seed=n)
    T = g["T"]
    stat = g["end_points"] / g["end_points"].sum()
    plan = _lib.Plan(T)
    X0 = batched_starts(T, B, seed=1)
    x32 = torch.from_numpy(X0).cuda()
    for tail in (16, 8):
        _lib.propagate_x(plan, x32, iters, stationary=stat, tail_bits=tail, want_eps=True)
        _lib.propagate_x(plan, x32, iters, tail_bits=tail, out_dtype=torch.float64)
    for shape_env in ({"CY2PX_CPL": "4"}, {"CY2PX_ROWS": "16"}, {"CY2PX_PERSIST": "1"}, {"CY2PX_EPS_PARK": "0"}):
        for k, v in shape_env.items():
            monkeypatch.setenv(k, v)
        _lib.propagate_x(plan, x32, iters, stationary=stat, want_eps=True)
        for k in shape_env:
            monkeypatch.delenv(k)
    for x in (x32, x32.double()):
        _lib.propagate(plan, x, iters, stationary=stat, want_final=True, history_stride=0, want_eps=True)
        _lib.propagate(plan, x, iters, stationary=stat, want_final=True, history_stride=2, want_eps=True)
        _lib.propagate(plan, x, iters, stationary=stat, want_final=False, history_stride=1, want_eps=True)
    assert _lib.check_guards() > 0


def test_float64_matrix_and_chains():
    import pandas as pd
    import scipy.sparse as sp
    import torch

    import cy2path_b200 as c2p
    from cy2path_b200 import _lib

    g = _graph(911, seed=3)
    T = g["T"].astype(np.float64)
    T.data *= 1.0 + 0.1 * np.random.default_rng(0).random(T.nnz)
    T = sp.csr_matrix(sp.diags(1.0 / np.asarray(T.sum(1)).ravel()) @ T)
    stat = g["end_points"] / g["end_points"].sum()
    plan = _lib.Plan(T)
    assert plan.float64_matrix
    x0 = torch.from_numpy(g["root_cells"] / g["root_cells"].sum()).cuda()
    _lib.propagate(plan, x0, 33, stationary=stat, want_final=True, history_stride=1, want_eps=True)
    _lib.propagate_x(plan, x0.float().reshape(-1, 1).repeat(1, 5).contiguous(), 7, stationary=stat, want_eps=True)

    class Duck:
        pass

    ad = Duck()
    ad.obsp, ad.shape, ad.uns = {"T_forward": T}, (T.shape[0], 0), {}
    ad.obs = pd.DataFrame({"root_cells": g["root_cells"], "end_points": g["end_points"]})
    idx, cum = c2p.extract_nonzero_entries(ad)
    rng = np.random.default_rng(5)
    u = rng.random((37, 61)) * 0.998
    c2p.sample_markov_chains(ad, num_chains=37, max_iter=62, convergence=10, uniforms=u,
                             init_states=rng.integers(0, T.shape[0], 37).astype(np.int32))
    assert _lib.check_guards() > 0


@pytest.mark.parametrize("kind,S,L,N,I,restricted", [("fhmm", 10, 4, 1237, 61, True), ("fhmm", 5, 3, 777, 45, False),
                                                     ("lssm", 6, 3, 1237, 130, True), ("nssm", 5, 2, 901, 23, True),
                                                     ("fhmm", 3, 1, 130, 7, True)])
def test_model_entry_points(kind, S, L, N, I, restricted, monkeypatch):
    import torch

    from cy2path_b200 import _lib, models

    g = _graph(N, seed=N)
    plan = _lib.Plan(g["T"])
    x0 = torch.tensor((g["root_cells"] / g["root_cells"].sum()).astype(np.float32)).cuda()
    D = _lib.propagate(plan, x0, I, history_stride=1, want_final=False)["history"].contiguous()
    torch.manual_seed(1)
    cls = {"fhmm": models.FHMM, "lssm": models.LSSM, "nssm": models.NSSM}[kind]
    m = cls(S, L, N, I, restricted=restricted, use_gpu=True)
    params = [p.data for p in m.parameters()]
    wts = (1.0, 0.1, 0.1, 0.05)
    for env in ({}, {"CY2P_FHMM_SPLIT_SMALL": "0"}, {"CY2P_FHMM_SPLIT_SMALL": "0", "CY2P_FHMM_FUSED_BWD": "0",
                                                    "CY2P_FHMM_HF_SMEM": "0"}, {"CY2P_FHMM_MMA": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _lib.fhmm_loss_grad(plan, *params, D, restricted, wts, model=kind)
        _lib.fhmm_forward(*params, I, restricted, want_pred=True, want_joint=(N < 1000), model=kind)
        for k in env:
            monkeypatch.delenv(k)
    _lib.fhmm_decode(*params, D, restricted, model=kind)
    assert _lib.check_guards() > 0


def test_trajectory_distances():
    import torch

    from cy2path_b200 import _lib

    s = torch.rand((23, 41, 13), dtype=torch.float64, device="cuda")
    _lib.trajectory_distance_matrix(s, "dtw")
    _lib.trajectory_distance_matrix(s, "hausdorff")
    assert _lib.check_guards() > 0
