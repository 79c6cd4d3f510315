"""Mirror of cy2path/dynamics/infer_posthoc.py:10-138 — conditionals over a subset of latent states and the
heuristic state selection. Host numpy over [S, L, N]-sized arrays; ``joint_probabilities`` may be the lazy
object of dynamics/lazy.py (its ``mean(0)`` / ``sum(-1)`` are closed forms) or a plain ndarray."""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger("cy2path_b200")


def _latent(adata):
    try:
        return adata.uns["latent_dynamics"]
    except KeyError:
        raise ValueError("Model outputs are missing. Run infer_latent_dynamics() first")


def compute_conditionals(adata, use_selected=True):
    """reference infer_posthoc.py:10-60: renormalise the time-averaged joint over the selected states and derive
    the six ``*_selected`` conditionals (unselected states are zeroed, so their rows come out 0/0 = nan exactly
    as in the reference)."""
    ld = _latent(adata)
    if use_selected:
        if "selected_states" not in ld.get("posthoc_computations", {}):
            raise ValueError("No selected states. Run latent_state_selection() first")
        selected = np.asarray(ld["posthoc_computations"]["selected_states"])
    else:
        selected = np.arange(ld["latent_dynamics_params"]["num_states"])
    mean = np.array(ld["model_outputs"]["joint_probabilities"].mean(0))  # [S, L, N]
    P = mean / mean[selected].sum()
    drop = np.ones(P.shape[0], dtype=bool)
    drop[selected] = False
    P[drop] = 0
    cp = ld["conditional_probabilities"]
    with np.errstate(invalid="ignore", divide="ignore"):
        over_l, over_s, over_n = P.sum(1), P.sum(0), P.sum(-1)
        cp["state_given_nodes_selected"] = over_l / over_l.sum(0)
        cp["chain_given_nodes_selected"] = over_s / over_s.sum(0)
        cp["chain_given_state_selected"] = (over_n.T / over_n.sum(1)).T
        cp["chain_state_given_nodes_selected"] = P / P.sum(0, keepdims=True).sum(1, keepdims=True)
        cp["node_given_chain_state_selected"] = P / P.sum(-1, keepdims=True)
        cp["node_given_state_selected"] = over_l / over_l.sum(-1, keepdims=True)


def latent_state_selection(adata, states=None, criteria="argmax_joint", min_ratio=0.01):
    """reference infer_posthoc.py:64-138. ``criteria``: None | 'argmax_joint' | 'argmax_smoothing' | 'viterbi'.
    (The reference fills ``state_ratio`` from ``np.empty``; states that own no node get ratio 0 here.)"""
    ld = _latent(adata)
    if not ld.get("model_outputs"):
        raise ValueError("Model outputs are missing. Run infer_latent_dynamics() first")
    num_states = ld["latent_dynamics_params"]["num_states"]
    mo = ld["model_outputs"]
    if criteria is None:
        selected = np.arange(num_states)
    elif criteria == "argmax_joint":
        selected = np.unique(np.asarray(mo["joint_probabilities"].sum(-1)).argmax(1))
    elif criteria == "argmax_smoothing":
        selected = np.unique(mo["log_smoothing"].argmax(1).flatten())
    elif criteria == "viterbi":
        selected = np.unique(mo["viterbi_path"].flatten())
    else:
        raise ValueError("unknown criteria %r" % (criteria,))
    if states is not None:
        selected = np.array(states)

    ld["posthoc_computations"] = {"selected_states": selected}
    ld["latent_dynamics_params"]["criteria"] = criteria
    ld["latent_dynamics_params"]["min_ratio"] = min_ratio
    compute_conditionals(adata, use_selected=True)

    if min_ratio is not None:  # drop states that are the most probable state of too few nodes
        owner = np.nan_to_num(ld["conditional_probabilities"]["state_given_nodes_selected"], nan=-1.0).argmax(0)
        ratio = np.bincount(owner, minlength=num_states)[:num_states] / adata.shape[0]
        selected = selected[ratio[selected] >= min_ratio]
    if criteria is not None and min_ratio is None and len(selected) == num_states:
        logger.warning("Number of kinetic states equals num_states. Consider initialising with more latent states")
    ld["posthoc_computations"]["selected_states"] = selected
    compute_conditionals(adata, use_selected=True)
