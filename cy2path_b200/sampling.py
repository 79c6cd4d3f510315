"""Drop-in mirror of cy2path/sampling.py — MSM state-probability simulation on the B200.

Same function names, arguments, return values and ``adata.uns`` keys as the reference
(sampling.py:13-124). The product ``x @ T`` and the cosine distance to the stationary vector run in
libcy2path_b200.so (csrc/propagate.cu); the host keeps only the argument checks, the convergence-cut
bookkeeping (a list scan over ``max_iter`` numbers) and the float64 export the reference promises.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _lib
from .utils import check_root_init, check_TPM, estimate_stationary_state

logger = logging.getLogger("cy2path_b200")

_PLAN_CACHE: dict = {}


def _fingerprint(T):
    """Content fingerprint of a scipy CSR matrix: sums over ALL stored values, column indices and row pointers (an
    in-place edit of any entry changes at least one of them up to cancellation) plus a CRC of the head and tail of
    each array. ~1 ms per million entries."""
    import zlib

    data, idx, ptr = np.asarray(T.data), np.asarray(T.indices), np.asarray(T.indptr)

    def crc(a):
        b = a.view(np.uint8).reshape(-1) if a.flags.c_contiguous else np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        return zlib.crc32(b[:65536].tobytes()) ^ zlib.crc32(b[-65536:].tobytes())

    w = np.arange(1, 1 + min(data.size, 1 << 16), dtype=np.float64)  # position-weighted sample: catches swaps
    stride = max(1, data.size // w.size)
    samp = data[::stride][: w.size].astype(np.float64)
    return (float(data.sum(dtype=np.float64)), float((samp * w[: samp.size]).sum()), int(idx.sum(dtype=np.int64)),
            int(ptr.sum(dtype=np.int64)), crc(data), crc(idx), crc(ptr))


def get_plan(T, device=None):
    """Plan (device copy of T, T^T, tables) for a scipy matrix, cached per matrix object.

    The cache holds a weak reference to the matrix (an entry dies with it, so a new matrix that reuses the id or the
    buffers of a freed one cannot hit a stale plan) and keys on a fingerprint of the whole contents (`_fingerprint`),
    so an in-place edit of T — e.g. making terminal rows absorbing through ``T.data[k] = ...`` — rebuilds the plan."""
    import weakref

    import torch

    dev = torch.cuda.current_device() if device is None else device
    key = (dev, T.shape, int(T.nnz), np.asarray(T.data).dtype.str) + _fingerprint(T)
    hit = _PLAN_CACHE.get((id(T), dev))
    if hit is not None and hit[0] == key and hit[2]() is T:
        return hit[1]
    plan = _lib.Plan(T, device=dev)
    if len(_PLAN_CACHE) >= 4:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    try:
        ref = weakref.ref(T, lambda _r, k=(id(T), dev): _PLAN_CACHE.pop(k, None))
    except TypeError:  # not weak-referenceable: keep a strong reference (the cache is 4 entries deep)
        ref = (lambda t: (lambda: t))(T)
    _PLAN_CACHE[(id(T), dev)] = (key, plan, ref)
    return plan


def torch_dtype(precision):
    import torch

    return torch.float32 if precision == "fp32" else torch.float64


def clear_plan_cache():
    _PLAN_CACHE.clear()


def check_convergence_criteria(eps_history, tol=1e-5):
    """reference sampling.py:13-25: (last i with |eps[i+1]-eps[i]| > tol) + 2, ``None`` if there is none."""
    eps = np.asarray(eps_history, dtype=np.float64)
    if eps.size < 2:
        return None
    above = np.nonzero(np.abs(np.diff(eps)) > tol)[0]
    if above.size == 0:
        return None
    return int(above[-1]) + 2


def batch_convergence(eps, tol, max_iter):
    """check_convergence_criteria + the None / >= max_iter -> max_iter mapping (sampling.py:48-57), per column."""
    eps = np.asarray(eps, dtype=np.float64)
    if eps.shape[0] < 2:
        return np.full(eps.shape[1], max_iter, dtype=np.int64)
    above = np.abs(np.diff(eps, axis=0)) > tol  # [iters-1, B]
    last = above.shape[0] - 1 - np.argmax(above[::-1], axis=0)
    conv = np.where(above.any(axis=0), last + 2, max_iter)
    return np.where(conv < max_iter, conv, max_iter).astype(np.int64)


def batch_convergence_device(eps, tol, max_iter):
    """batch_convergence on the device tensor eps [iters, B] (float64): same comparisons, so the same integers;
    saves a host pass over iters x B values per call."""
    import torch

    if eps.shape[0] < 2:
        return torch.full((eps.shape[1],), max_iter, dtype=torch.int64, device=eps.device)
    above = (eps[1:] - eps[:-1]).abs() > tol  # [iters-1, B]
    idx = torch.arange(above.shape[0], device=eps.device, dtype=torch.int64)[:, None]
    last = torch.where(above, idx, torch.full_like(idx, -1)).amax(0)
    conv = torch.where(last >= 0, last + 2, torch.full_like(last, max_iter))
    return torch.clamp(conv, max=max_iter)


def iterate_state_probability(adata, matrix_key="T_forward", init=None, stationary=None, max_iter=1000, tol=1e-5):
    """reference sampling.py:29-60. Returns (state_history, state_history_max_iter, convergence_check);
    histories are float64 host arrays as in the reference. The single-start case runs with a float64
    iterate on the device (cy2p_propagate_f64): it is latency-bound, so float64 costs nothing, and the
    history then tracks the reference's float64 arithmetic to ~1e-15 instead of drifting."""
    import torch

    T = adata.obsp[matrix_key]
    plan = get_plan(T)
    n = plan.n
    init64 = np.asarray(init, dtype=np.float64).reshape(n)
    dev = plan.device
    x0 = torch.from_numpy(np.ascontiguousarray(init64)).to(dev, non_blocking=True)
    out = _lib.propagate(plan, x0, int(max_iter), stationary=stationary, want_final=False, history_stride=1,
                         want_eps=True)
    if max_iter > 0:
        state_history_max_iter = _lib.to_host(out["history"]).reshape(max_iter, n)
    else:
        state_history_max_iter = np.empty((0, n), dtype=np.float64)
    eps_history = out["eps"].cpu().numpy() if max_iter > 0 else np.empty(0)

    convergence_check = check_convergence_criteria(eps_history, tol=tol)
    if isinstance(convergence_check, int) and convergence_check < max_iter:
        logger.info("Tolerance reached after {} iterations of {}.".format(convergence_check, max_iter))
    else:
        logger.warning("Max number ({}) of iterations reached.".format(max_iter))
        convergence_check = max_iter
    state_history = state_history_max_iter[:convergence_check]
    return state_history, state_history_max_iter, convergence_check


PRECISIONS = ("f48", "f40", "fp64", "fp32")


def propagate_batch(adata_or_T, X0, iters, stationary=None, matrix_key="T_forward", tol=1e-5, return_device=False,
                    precision="f48", out_dtype=None):
    """Batched form of the same simulation (BASELINE configs 2/4): B start distributions at once.

    X0: [n, B] (numpy or cuda tensor), cell-major. Returns dict with ``final`` [n, B] and, when
    ``stationary`` is given, ``eps`` [iters, B] float64 and ``convergence`` [B] (the per-start cut of
    reference sampling.py:48-57). The full [iters, n, B] history is not materialised (1.6 PB at config 2).

    ``precision`` names the iterate the device carries between updates (the reference's is float64,
    sampling.py:33; products and row sums are float64 in every mode but "fp32"):
      "f48"  (default) truncated float64, 36 mantissa bits, 6 bytes: within 1e-9 relative L1 of the reference
             after 2,000 updates; X0 is taken as float32, ``final`` is float32 (``out_dtype=torch.float64``: float64)
      "f40"  truncated float64, 28 mantissa bits, 5 bytes: ~1e-10 relative L1 per update (2e-7 after 2,000)
      "fp64" float64 iterate, 8 bytes (X0 / final float64)
      "fp32" float32 iterate and arithmetic: the fastest, but it leaves the north_star's 1e-6 relative-L1 bar after
             ~60-100 updates (2e-8 per update, DESIGN.md "precision"); kept for short runs and as a measurement"""
    import torch

    if precision not in PRECISIONS:
        raise ValueError("precision must be one of %s" % (PRECISIONS,))
    T = adata_or_T.obsp[matrix_key] if hasattr(adata_or_T, "obsp") else adata_or_T
    plan = get_plan(T)
    if precision in ("f48", "f40"):
        x0 = _lib.to_device(X0, torch.float32, plan.device)
        out = _lib.propagate_x(plan, x0, int(iters), stationary=stationary, tail_bits=16 if precision == "f48" else 8,
                               want_eps=stationary is not None, out_dtype=out_dtype)
    else:
        x0 = _lib.to_device(X0, torch_dtype(precision), plan.device)
        out = _lib.propagate(plan, x0, int(iters), stationary=stationary, want_final=True, history_stride=0,
                             want_eps=stationary is not None)
    conv_d = batch_convergence_device(out["eps"], tol, int(iters)) if out["eps"] is not None else None
    fin, eps, conv = _lib.to_host_many([None if return_device else out["final"], out["eps"], conv_d])
    res = {"final": out["final"] if return_device else fin}
    if eps is not None:
        res["eps"] = eps
        res["convergence"] = conv
    return res


def sample_state_probability(data, matrix_key="T_forward", recalc_matrix=False, self_transitions=False,
                             init="root_cells", max_iter=1000, tol=1e-5, copy=False):
    """reference sampling.py:64-124 — same checks, same ``adata.uns['state_probability_sampling']`` keys."""
    adata = data.copy() if copy else data
    check_TPM(adata, matrix_key=matrix_key, recalc_matrix=recalc_matrix, self_transitions=self_transitions)
    init_state_probability, init_type = check_root_init(adata, init=init)
    stationary_state_probability = estimate_stationary_state(adata, matrix_key=matrix_key)
    state_history, state_history_max_iter, convergence_check = iterate_state_probability(
        adata, matrix_key=matrix_key, init=init_state_probability, stationary=stationary_state_probability,
        max_iter=max_iter, tol=tol)
    adata.uns["state_probability_sampling"] = {
        "state_history": state_history,
        "state_history_max_iter": state_history_max_iter,
        "init_state_probability": init_state_probability,
        "stationary_state_probability": stationary_state_probability,
        "sampling_params": {
            "recalc_matrix": recalc_matrix,
            "self_transitions": self_transitions,
            "init_type": init_type,
            "max_iter": max_iter,
            "convergence": convergence_check,
            "tol": tol,
            "copy": copy,
        },
    }
    if copy:
        return adata
