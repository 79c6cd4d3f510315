"""GPU tests of the error behaviour of the later entry points (models, trajectory distances, halo push): bad
arguments surface as the exception types the Python layer documents, never as a crash or a silent fallback."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_model_size_limits_and_workspace():
    from cy2path_b200 import _lib, models
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import knn_velocity_tpm

    lib = _lib.load()
    assert lib.cy2p_fhmm_ws_bytes(33, 2, 100, 10) == 0 and lib.cy2p_lssm_ws_bytes(4, 17, 100, 10) == 0
    assert lib.cy2p_nssm_ws_bytes(20, 4, 100, 10) == 0  # S*L = 80 > 64 emission rows
    m = models.NSSM(20, 4, 50, 6)
    with pytest.raises(ValueError):
        m.forward_model()
    m = models.FHMM(33, 2, 50, 6)
    with pytest.raises(ValueError):
        m.forward_model()
    # unrestricted FHMM / LSSM with S*L > 64 are rejected by the C layer itself
    m = models.LSSM(20, 4, 50, 6, restricted=False)
    with pytest.raises((ValueError, _lib.Cy2pError)):
        m.forward_model()
    # a work space that is too small or a plan of the wrong size is an error, not a crash
    g = knn_velocity_tpm(300, k=8, seed=1)
    plan = get_plan(g["T"])
    m = models.LSSM(3, 2, 300, 5)
    ps = [p.data for p in m.parameters()]
    D = torch.rand(5, 300, device="cuda")
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    terms = torch.empty(8, device="cuda")
    grads = [torch.empty_like(p) for p in ps]
    hw = (ctypes.c_float * 4)(1.0, 0.1, 0.1, 0.0)
    rc = lib.cy2p_lssm_loss_grad(plan.handle, 3, 2, 300, 5, 1, *[_lib.ptr(p) for p in ps], _lib.ptr(D), hw,
                                 _lib.ptr(terms), *[_lib.ptr(x) for x in grads], _lib.ptr(ws), 256, None)
    assert rc == _lib.ERR_WORKSPACE and b"work space" in lib.cy2p_last_error()
    m2 = models.LSSM(3, 2, 299, 5)
    with pytest.raises(ValueError):
        _lib.fhmm_loss_grad(plan, *[p.data for p in m2.parameters()], torch.rand(5, 299, device="cuda"), True,
                            (1.0, 0.1, 0.1, 0.0), model="lssm")
    with pytest.raises(ValueError):
        _lib.fhmm_forward(*ps, 5, True, model="hmm")


def test_trajectory_distance_argument_checks():
    from cy2path_b200 import _lib
    from cy2path_b200.cytopath import compute_dtw_distance_matrix, compute_Hausdorff_distance

    one = compute_dtw_distance_matrix(np.random.default_rng(0).standard_normal((1, 9, 3)))
    assert one.shape == (1, 1) and one[0, 0] == 0.0
    assert compute_Hausdorff_distance(np.zeros((3, 1, 2))).tolist() == [[0.0] * 3] * 3  # single-point series
    with pytest.raises(ValueError):
        compute_dtw_distance_matrix(np.zeros((3, 5, 65)))  # more than 64 dimensions
    with pytest.raises(ValueError):
        _lib.trajectory_distance_matrix(torch.zeros(3, 5, device="cuda"), "dtw")
    with pytest.raises(ValueError):
        _lib.trajectory_distance_matrix(torch.zeros(3, 5, 2, device="cuda"), "euclid")
    with pytest.raises(NotImplementedError):
        compute_Hausdorff_distance(np.zeros((2, 3, 2)), distance="manhattan")
    lib = _lib.load()
    s = torch.zeros(2, 4, 8, dtype=torch.float64, device="cuda")
    out = torch.empty(2, 2, dtype=torch.float64, device="cuda")
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")
    assert lib.cy2p_dtw_matrix(_lib.ptr(s), 2, 4, 8, _lib.ptr(out), _lib.ptr(ws), 64, None) == _lib.ERR_WORKSPACE
    assert lib.cy2p_dtw_matrix(_lib.ptr(s), 2, 4, 7, _lib.ptr(out), _lib.ptr(ws), 64, None) == _lib.ERR_INVALID_ARG


def test_halo_push_argument_checks():
    from cy2path_b200 import _lib
    from cy2path_b200.synth import knn_velocity_tpm

    lib = _lib.load()
    g = knn_velocity_tpm(200, k=6, seed=2)
    plan = _lib.Plan(g["T"])
    X = torch.rand(200, 4, device="cuda")
    Y = torch.zeros_like(X)
    ptrs = torch.tensor([Y.data_ptr()], dtype=torch.int64, device="cuda")
    masks = torch.ones(200, dtype=torch.int32, device="cuda")
    ok = lib.cy2p_propagate_rows_push(plan.handle, _lib.ptr(X), ctypes.c_void_p(ptrs.data_ptr()), 1, _lib.ptr(masks), 4,
                                      0, 200, None)
    assert ok == _lib.OK
    torch.cuda.synchronize()
    ref = torch.zeros_like(X)
    assert lib.cy2p_propagate_rows_ordered(plan.handle, _lib.ptr(X), _lib.ptr(ref), 4, 0, 200, None) == _lib.OK
    torch.cuda.synchronize()
    assert torch.equal(Y, ref)
    bad = lib.cy2p_propagate_rows_push(plan.handle, _lib.ptr(X), ctypes.c_void_p(ptrs.data_ptr()), 33, _lib.ptr(masks),
                                       4, 0, 200, None)
    assert bad == _lib.ERR_INVALID_ARG
    bad = lib.cy2p_propagate_rows_push(plan.handle, _lib.ptr(X), None, 1, _lib.ptr(masks), 4, 0, 200, None)
    assert bad == _lib.ERR_INVALID_ARG
    bad = lib.cy2p_propagate_rows_push(plan.handle, _lib.ptr(X), ctypes.c_void_p(ptrs.data_ptr()), 1, _lib.ptr(masks), 4,
                                       10, 300, None)
    assert bad == _lib.ERR_INVALID_ARG
