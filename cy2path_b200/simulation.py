"""Drop-in mirror of cy2path/simulation.py — Markov-chain sampling on the B200.

Same function names, arguments, return values and ``adata.uns['markov_chain_sampling']`` keys as the
reference (simulation.py:15-265). Table construction, the inverse-CDF walks and the per-step histogram
run in libcy2path_b200.so (csrc/plan.cu, csrc/sample.cu). The reference fans chains out over joblib
processes whose RNGs are unseeded; here the uniforms are drawn once (``np.random.rand`` by default, or
supplied by the caller) and all chains walk in one kernel, so ``n_jobs`` is accepted and ignored.
"""
from __future__ import annotations

import ctypes
import logging

import numpy as np

from . import _lib
from .sampling import check_convergence_criteria, get_plan, iterate_state_probability
from .utils import check_root_init, check_TPM, estimate_stationary_state

logger = logging.getLogger("cy2path_b200")


def _tables_device(plan):
    """Device CDF tables of a plan (built once per plan, kept on it)."""
    torch = _lib._torch()
    if getattr(plan, "_tables", None) is None:
        W = max(plan.W, 1)
        with torch.cuda.device(plan.device):
            idx = _lib.dev_empty((plan.n, W), torch.int32, plan.device, "CDF idx")
            cum = _lib.dev_empty((plan.n, W), torch.float32, plan.device, "CDF cum")
            _lib.check(_lib.load().cy2p_cdf_build(plan.handle, _lib.ptr(idx), _lib.ptr(cum), W,
                                                  _lib.stream_ptr(plan.device)))
        plan._tables = (idx, cum)
    return plan._tables


def extract_nonzero_entries(adata, matrix_key="T_forward"):
    """reference simulation.py:15-53 -> (int32 [n, W] neighbour indices, float32 [n, W] rounded CDF)."""
    plan = get_plan(adata.obsp[matrix_key])
    idx, cum = _tables_device(plan)
    return idx.cpu().numpy(), cum.cpu().numpy()


def _walk(plan, idx_d, cum_d, init_states, uniforms, want_history_steps=None, want_counts=False):
    """Run cy2p_sample_chains (+ cy2p_chain_history). init_states int32 [C], uniforms float64 [C, steps]
    (numpy or cuda tensors). Returns cuda tensors (states [C, steps+1], probs [C, steps], hist or None);
    with ``want_counts`` the int32 visit counts behind ``hist`` are returned as a fourth element (the
    chain-sharded mode sums them over ranks before normalising, dist.sample_chains_sharded)."""
    torch = _lib._torch()
    lib = _lib.load()
    dev = plan.device

    def to_dev(a, dtype):
        return _lib.to_device(a, dtype, dev)

    with torch.cuda.device(dev):
        init_d = to_dev(init_states, torch.int32).reshape(-1)
        u_d = to_dev(uniforms, torch.float64)
        C, steps = int(u_d.shape[0]), int(u_d.shape[1])
        if init_d.numel() != C:
            raise ValueError("one init state per chain required")
        if C and (int(init_d.min()) < 0 or int(init_d.max()) >= plan.n):
            raise IndexError("init_state out of range")
        states = _lib.dev_empty((C, steps + 1), torch.int32, dev, "chain states")
        probs = _lib.dev_empty((C, steps), torch.float32, dev, "chain probs")
        err = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.check(lib.cy2p_sample_chains(plan.handle, _lib.ptr(idx_d), _lib.ptr(cum_d), int(idx_d.shape[1]),
                                          _lib.ptr(init_d), _lib.ptr(u_d), C, steps, _lib.ptr(states),
                                          _lib.ptr(probs), _lib.ptr(err), _lib.stream_ptr(dev)))
        hist = counts = None
        if want_history_steps:
            counts = _lib.dev_empty((want_history_steps, plan.n), torch.int32, dev, "visit counts")
            hist = _lib.dev_empty((want_history_steps, plan.n), torch.float32, dev, "visit histogram")
            _lib.check(lib.cy2p_chain_history(_lib.ptr(states), C, steps + 1, want_history_steps, plan.n,
                                              _lib.ptr(counts), _lib.ptr(hist), _lib.stream_ptr(dev)))
    if want_counts:
        return states, probs, hist, counts
    return states, probs, hist


def iterate_markov_chain(adata, matrix_key="T_forward", max_iter=1000, init_state=None, nonzero_entries=None,
                         uniforms=None):
    """reference simulation.py:57-87: one trajectory. ``uniforms`` (float64 [max_iter-1]) replaces the
    ``np.random.rand(max_iter - 1)`` draw (simulation.py:69) when given; same global RNG otherwise."""
    torch = _lib._torch()
    plan = get_plan(adata.obsp[matrix_key])
    if nonzero_entries is None:
        idx_d, cum_d = _tables_device(plan)
    else:
        idx_d = torch.as_tensor(np.ascontiguousarray(nonzero_entries[0], dtype=np.int32), device=plan.device)
        cum_d = torch.as_tensor(np.ascontiguousarray(nonzero_entries[1], dtype=np.float32), device=plan.device)
    random_step = np.random.rand(max_iter - 1) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
    init = np.asarray(init_state, dtype=np.int32).reshape(1)
    states, probs, _ = _walk(plan, idx_d, cum_d, init, random_step.reshape(1, -1))
    return states[0].cpu().numpy(), probs[0].cpu().numpy()


def sample_markov_chains(data, matrix_key="T_forward", recalc_matrix=False, self_transitions=False,
                         init="root_cells", repeat_root=True, num_chains=1000, max_iter=1000, convergence="auto",
                         tol=1e-5, n_jobs=-1, copy=False, uniforms=None, init_states=None):
    """reference simulation.py:90-265. Extra keyword-only-in-spirit arguments ``uniforms``
    ([num_chains, max_iter-1] float64) and ``init_states`` ([num_chains] int) pin the random draws
    (bit-exact paths for identical draws); left ``None`` they are drawn from numpy's global RNG in the
    order the reference's parent process consumes it (initial states first)."""
    adata = data.copy() if copy else data
    check_TPM(adata, matrix_key=matrix_key, recalc_matrix=recalc_matrix, self_transitions=self_transitions)
    init_state_probability, init_type = check_root_init(adata, init=init)
    stationary_state_probability = estimate_stationary_state(adata, matrix_key=matrix_key)

    n = adata.shape[0]
    plan = get_plan(adata.obsp[matrix_key])
    idx_d, cum_d = _tables_device(plan)
    if init_states is None:
        p = np.asarray(init_state_probability, dtype=np.float64)
        if repeat_root:
            init_states = np.random.choice(n, num_chains, replace=True, p=p)
        else:  # the reference draws one state per call (size 1), so replace=False never bites
            init_states = np.concatenate([np.random.choice(n, 1, replace=False, p=p) for _ in range(num_chains)])
    if uniforms is None:
        uniforms = np.random.rand(num_chains, max_iter - 1)
    states_d, probs_d, hist_d = _walk(plan, idx_d, cum_d, np.asarray(init_states, dtype=np.int32), uniforms,
                                      want_history_steps=max_iter)
    state_indices = _lib.to_host(states_d)
    state_transition_probabilities = _lib.to_host(probs_d)
    state_history_max_iter = _lib.to_host(hist_d)  # unvisited entries are 0 (np.empty in the reference)

    cluster_out = None
    if isinstance(convergence, str) and convergence == "auto":
        convergence_check = iterate_state_probability(adata, matrix_key=matrix_key, init=init_state_probability,
                                                      stationary=stationary_state_probability, max_iter=max_iter,
                                                      tol=tol)[-1]
    elif isinstance(convergence, (int, np.integer)) and not isinstance(convergence, bool):
        convergence_check = int(convergence)
    elif convergence in adata.obs.columns:
        convergence_check, cluster_out = _cluster_convergence(adata, convergence, stationary_state_probability,
                                                              state_indices, max_iter, tol)
    else:
        raise ValueError("convergence must be 'auto', an int or a column of adata.obs")

    adata.uns["markov_chain_sampling"] = {
        "state_indices": state_indices[:, :convergence_check],
        "state_history": state_history_max_iter[:convergence_check],
        "state_indices_max_iter": state_indices,
        "state_history_max_iter": state_history_max_iter,
        "state_transition_probabilities_max_iter": state_transition_probabilities,
        "init_state_probability": init_state_probability,
        "stationary_state_probability": stationary_state_probability,
        "sampling_params": {
            "recalc_matrix": recalc_matrix,
            "self_transitions": self_transitions,
            "init_type": init_type,
            "max_iter": max_iter,
            "num_chains": num_chains,
            "convergence": convergence_check,
            "convergence_criterion": convergence,
            "tol": tol,
            "copy": copy,
        },
    }
    if cluster_out is not None:
        by_cluster, sequences, proportions = cluster_out
        u = adata.uns["markov_chain_sampling"]
        u["stationary_state_by_cluster"] = by_cluster
        u["cluster_sequences"] = sequences[:, :convergence_check]
        u["cluster_proportions"] = proportions[:, :convergence_check].T
        u["cluster_sequences_max_iter"] = sequences
        u["cluster_proportions_max_iter"] = proportions.T
    if copy:
        return adata


def _cluster_convergence(adata, column, stationary, state_indices, max_iter, tol):
    """reference simulation.py:178-216 (cluster-proportion cosine criterion). Host pandas bookkeeping over
    [num_clusters, max_iter] numbers; the reference's ``np.float`` (removed in numpy 1.24) is float64 here."""
    import pandas as pd
    from scipy.spatial.distance import cosine

    cats = adata.obs[column].astype("category")
    by_cluster = pd.DataFrame({"cluster": cats, "probability": stationary}).groupby("cluster", observed=False).sum()
    sequences = np.asarray(adata.obs[column].astype(str), dtype=object)[state_indices]
    proportions = np.zeros((len(by_cluster), max_iter), dtype=np.float64)
    for r, cat in enumerate(by_cluster.index):
        proportions[r] = (sequences == str(cat)).mean(axis=0)
    target = by_cluster.values.flatten()
    eps_history = np.array([cosine(proportions[:, k], target) for k in range(max_iter)])
    conv = check_convergence_criteria(eps_history, tol)
    if isinstance(conv, int) and conv < max_iter:
        logger.info("Tolerance reached after {} iterations of {}.".format(conv, max_iter))
    else:
        logger.warning("Max number ({}) of iterations reached.".format(max_iter))
        conv = max_iter
    return conv, (by_cluster.values, sequences, proportions)
