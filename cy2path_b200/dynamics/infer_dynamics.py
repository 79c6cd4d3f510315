"""Mirror of cy2path/dynamics/infer_dynamics.py (reference :13-285): fit a latent dynamic model to the MSM
simulation stored in ``adata`` and export model outputs, parameters and conditionals to
``adata.uns['latent_dynamics']`` under the reference's keys.

Differences that follow from the CUDA path:
* T is handed to the trainer as sparse CSR, not densified (reference :204);
* one fused decoder call instead of three (filtering / smoothing / viterbi, reference :24-26);
* the conditionals come from the time-averaged joint's factors ([S, L, N]); the [I, S, L, N] joint is never built
  on the device, and ``joint_probabilities`` is stored lazily above ``dense_limit_bytes`` (dynamics/lazy.py).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..utils import check_TPM
from .lazy import JointProbabilities

logger = logging.getLogger("cy2path_b200")

LSE = torch.logsumexp


def _np(t):
    return t.detach().cpu().numpy()


def _joint_factors(model, hidden):
    """static [S, L, N] and hidden [I, S, L] (natural logs) with joint = static[None] + hidden[..., None]."""
    S, L = model.num_states, model.num_chains
    kind = model._kind
    logE, logw = model.log_emission_matrix, model.log_chain_weights
    if kind == "nssm":  # log_weights [S, N, L] = log E[s, n] + log cw[s', n, l]   (_NSSM.py:80-82)
        static = (logE.unsqueeze(-1) + logw).permute(0, 2, 1)
        hid = hidden[:, :, None].expand(-1, -1, L)
    else:
        E = logE.expand(S, L, -1)  # [S, 1|L, N] -> [S, L, N]
        if kind == "fhmm":  # _FHMM.py:126-141
            static = E + logw[None, :, None]
            hid = hidden
        else:  # _LSSM.py:98-111
            static = E + logw[:, :, None]
            hid = hidden[:, :, None].expand(-1, -1, L)
    return static.contiguous(), hid.contiguous()


def extract_model_outputs(adata, model, dense_limit_bytes=256 << 20):
    """reference infer_dynamics.py:13-123 — same keys under ``adata.uns['latent_dynamics']``:
    model_outputs {latent_state_history, predicted_state_history, joint_probabilities, log_filtering,
    log_smoothing, log_viterbi, viterbi_path}, model_params {chain_weights, emission_matrix,
    latent_transition_matrix[, conditional_latent_transition_matrix]}, conditional_probabilities {6 keys}."""
    ld = adata.uns["latent_dynamics"]
    pred, _, hidden = model.forward_model(want_joint=False)
    D = torch.as_tensor(np.asarray(adata.uns["state_probability_sampling"]["state_history"]), dtype=torch.float32)
    dec = model._decode(D)  # sets the log_* attributes the way the reference's last decoder call leaves them
    best = _np(dec["best_path"])  # [L, I]

    static, hid = _joint_factors(model, hidden)
    I, S, L, N = hid.shape[0], static.shape[0], static.shape[1], static.shape[2]
    lazy = JointProbabilities(_np(static), _np(hid))
    mo = ld["model_outputs"] = {}
    mo["latent_state_history"] = _np(hidden.exp())
    mo["predicted_state_history"] = _np(pred.exp())
    mo["joint_probabilities"] = np.asarray(lazy) if 4 * I * S * L * N <= dense_limit_bytes else lazy
    mo["log_filtering"] = _np(dec["log_alpha"])
    mo["log_smoothing"] = _np(dec["log_gamma"])
    mo["log_viterbi"] = _np(dec["log_delta"])
    mo["viterbi_path"] = best.astype(int).T

    mp = ld["model_params"] = {}
    mp["chain_weights"] = _np(model.log_chain_weights.exp())
    mp["emission_matrix"] = _np(model.log_emission_matrix.exp())
    A = _np(model.log_transition_matrix.exp())
    if ld["latent_dynamics_params"]["mode"] == "FHMM":
        # the reference scales the per-chain matrices by the chain weights IN PLACE before summing, so its
        # 'conditional_latent_transition_matrix' is the weighted stack as well (infer_dynamics.py:66-77)
        A = A * mp["chain_weights"][:, None, None]
        mp["conditional_latent_transition_matrix"] = A
        mp["latent_transition_matrix"] = A.sum(0)
    else:
        mp["latent_transition_matrix"] = A

    # conditionals of the time-averaged joint M[s,l,n] = static + log mean_t exp(hidden)   (reference :80-123)
    M = static + (LSE(hid, 0) - float(np.log(I)))[..., None]
    Ms, Ml, Msl = LSE(M, 1), LSE(M, 0), LSE(M, -1)
    cp = ld["conditional_probabilities"] = {}
    cp["state_given_nodes"] = _np((Ms - LSE(Ms, 0, keepdim=True)).exp())
    cp["chain_given_nodes"] = _np((Ml - LSE(Ml, 0, keepdim=True)).exp())
    cp["chain_given_state"] = _np((Msl - LSE(Msl, 1, keepdim=True)).exp())
    cp["chain_state_given_nodes"] = _np((M - LSE(LSE(M, 0, keepdim=True), 1, keepdim=True)).exp())
    cp["node_given_chain_state"] = _np((M - LSE(M, -1, keepdim=True)).exp())
    cp["node_given_state"] = _np((Ms - LSE(Ms, -1, keepdim=True)).exp())


def infer_dynamics(data, model=None, num_states=10, num_chains=1, num_epochs=500, mode="FHMM", restricted=True,
                   use_gpu=False, verbose=False, precomputed_emissions=None, precomputed_transitions=None,
                   emissions_grad=False, transitions_grad=False, save_model="./model", load_model=None, copy=False,
                   **kwargs):
    """reference infer_dynamics.py:126-285; same arguments and return value (``model`` or ``(model, adata)``).
    ``use_gpu`` is accepted for compatibility: the model always runs on the CUDA device."""
    from .. import models

    adata = data.copy() if copy else data
    params_ = dict(locals())
    for k in ("data", "adata", "models"):
        params_.pop(k, None)

    try:
        state_history = adata.uns["state_probability_sampling"]["state_history"]
    except (KeyError, TypeError):
        raise ValueError("State probability history could not be recovered. Run sample_state_probability() first")
    check_TPM(adata)
    TPM = adata.obsp["T_forward"]  # sparse; the plan is built from it (the reference densifies, :204)
    state_history = torch.as_tensor(np.asarray(state_history), dtype=torch.float32)

    if model is None:
        if mode not in ("FHMM", "LSSM", "NSSM"):
            raise ValueError("mode must be 'FHMM', 'LSSM' or 'NSSM'")
        model = getattr(models, mode)(num_states, num_chains, adata.shape[0], state_history.shape[0],
                                      restricted=restricted, use_gpu=use_gpu)
        if load_model is not None:  # the reference dereferences model=None here (:245-246); load the state dict
            sd = torch.load(load_model) if isinstance(load_model, (str, bytes)) else load_model
            model.load_state_dict(sd)
    else:
        num_states, num_chains = model.num_states, model.num_chains
        params_["num_states"], params_["num_chains"] = num_states, num_chains

    dev = model.unnormalized_state_init.device
    if precomputed_emissions is not None:
        model.unnormalized_emission_matrix = torch.nn.Parameter(
            torch.log(torch.as_tensor(precomputed_emissions, dtype=torch.float32)).to(dev))
        model.unnormalized_emission_matrix.requires_grad_(emissions_grad)
    if precomputed_transitions is not None:
        model.unnormalized_transition_matrix = torch.nn.Parameter(
            torch.log(torch.as_tensor(precomputed_transitions, dtype=torch.float32)).to(dev))
        model.unnormalized_transition_matrix.requires_grad_(transitions_grad)

    model.train(state_history, TPM=TPM, num_epochs=num_epochs, verbose=verbose, **kwargs)

    adata.uns["latent_dynamics"] = {"latent_dynamics_params": params_}
    extract_model_outputs(adata, model)
    adata.uns["latent_dynamics"]["log_likelihood"] = model.log_likelihood.item()
    if save_model is not None:
        torch.save(model.state_dict(), save_model)
    return (model, adata) if copy else model
