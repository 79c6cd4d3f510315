"""Mirror of cy2path/models/_FHMM.py — same constructor, parameters and methods (reference _FHMM.py:7-313)."""
from __future__ import annotations

import torch

from .. import _lib
from .methods import log_transform_params
from .trainer import train


class FHMM(torch.nn.Module):
    """Latent dynamic model with a common latent state space and conditional state transitions per
    chain, trained on a Markov state simulation of an observed-state TPM (reference _FHMM.py:7-32).

    P(node | iter) = sum_chain sum_state P(node | state, chain) P(state | chain, iter) P(chain)

    Parameters are the reference's four unnormalised tensors, drawn in the same order from the same
    (CPU) generator so ``torch.manual_seed`` reproduces the reference's initialisation, then kept on
    the CUDA device: there is no CPU execution path (``use_gpu`` is accepted for compatibility).
    """

    _kind = "fhmm"  # selects the cy2p_<kind>_* entry points (LSSM overrides it)

    def __init__(self, num_states, num_chains, num_nodes, num_iters, restricted=True, use_gpu=False):
        super().__init__()
        self.num_nodes = num_nodes
        self.num_chains = num_chains
        self.num_states = num_states
        self.num_iters = num_iters
        self.restricted = restricted
        self.use_gpu = use_gpu
        _lib._torch()  # raises without a CUDA device

        self.unnormalized_state_init = torch.nn.Parameter(torch.randn(self.num_states, self.num_chains))
        self.unnormalized_chain_weights = torch.nn.Parameter(torch.randn(self.num_chains))
        if self.restricted:
            em_init = torch.randn(self.num_states, 1, self.num_nodes)
        else:
            em_init = torch.randn(self.num_states, self.num_chains, self.num_nodes)
        self.unnormalized_emission_matrix = torch.nn.Parameter(em_init)
        self.unnormalized_transition_matrix = torch.nn.Parameter(
            torch.randn(self.num_chains, self.num_states, self.num_states))
        self.is_cuda = True
        self.cuda()

    def forward_model(self, want_joint=True):
        """reference _FHMM.py:84-150 -> (log P(node|iter) [I,N], joint [I,S,L,N], hidden [I,S,L]).

        ``want_joint=False`` skips the I*S*L*N tensor (4 GB at S=10, L=4, N=50k, I=500)."""
        out = _lib.fhmm_forward(self.unnormalized_state_init, self.unnormalized_chain_weights,
                                self.unnormalized_emission_matrix, self.unnormalized_transition_matrix,
                                self.num_iters, self.restricted, want_pred=True, want_joint=want_joint,
                                model=self._kind)
        self.log_state_init = out["log_pi"]
        self.log_chain_weights = out["log_w"]
        self.log_emission_matrix = out["log_E"]
        self.log_transition_matrix_ = out["log_A"]
        self.log_transition_matrix = out["log_A_mix"]  # mixed over chains, as the reference leaves it (:93-96)
        return out["pred"], out["joint"], out["hidden"]

    # ---- decoders (reference _FHMM.py:153-281): one fused CUDA call computes all three
    def _decode(self, D):
        out = _lib.fhmm_decode(self.unnormalized_state_init, self.unnormalized_chain_weights,
                               self.unnormalized_emission_matrix, self.unnormalized_transition_matrix,
                               torch.as_tensor(D), self.restricted, model=self._kind)
        log_transform_params(self)
        self.log_transition_matrix_ = self.log_transition_matrix
        return out

    def filtering(self, D):
        """reference _FHMM.py:153-187 -> (log_alpha [I,S,L], log_probs [I,L])."""
        out = self._decode(D)
        return out["log_alpha"], out["log_probs"]

    def smoothing(self, D):
        """reference _FHMM.py:190-236 -> (log_beta, log_gamma), both [I,S,L]. Like the reference it conditions every
        step on the LAST row of D (_FHMM.py:209,223)."""
        out = self._decode(D)
        return out["log_beta"], out["log_gamma"]

    def viterbi(self, D):
        """reference _FHMM.py:238-281 -> (log_delta [I,S,L], psi [I,S,L] float, log_max [I,L], best_path).
        best_path is a list (per chain) of lists of length I: floats from psi, the last entry an int, as the
        reference builds it (:270-279)."""
        out = self._decode(D)
        bp = out["best_path"].cpu().numpy()
        best_path = []
        for i in range(self.num_chains):
            row = [float(v) for v in bp[i]]
            row[-1] = int(row[-1])
            best_path.append(row)
        return out["log_delta"], out["psi"], out["log_max"], best_path

    def train(self, D=None, TPM=None, num_epochs=500, sparsity_weight=1.0, exclusivity_weight=1e-1,
              orthogonality_weight=1e-1, TPM_weight=0.0, optimizer=None, criterion=None, swa_scheduler=None,
              swa_start=200, verbose=False):
        """reference _FHMM.py:284-313. ``TPM`` may be a scipy sparse matrix (preferred), a dense
        array/tensor (as the reference passes it) or a cy2path_b200 plan."""
        if D is None or isinstance(D, bool):  # nn.Module.train(mode) compatibility
            return super().train(True if D is None else D)
        train(self, D, TPM=TPM, num_epochs=num_epochs, sparsity_weight=sparsity_weight,
              exclusivity_weight=exclusivity_weight, orthogonality_weight=orthogonality_weight, TPM_weight=TPM_weight,
              optimizer=optimizer, criterion=criterion, swa_scheduler=swa_scheduler, swa_start=swa_start,
              verbose=verbose)

    def _log_transform_params(self):
        return log_transform_params(self)
