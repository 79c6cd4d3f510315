"""Mirror of cy2path/models/_NSSM.py — same constructor, parameters and methods (reference _NSSM.py:7-285)."""
from __future__ import annotations

import torch

from .. import _lib
from ._FHMM import FHMM
from .trainer import train


class NSSM(FHMM):
    """Latent dynamic model with a common latent state space and node-level chain weights, trained on a Markov
    state simulation of an observed-state TPM (reference _NSSM.py:7-31).

    P(node | iter) = sum_chain sum_state P(chain | node, state) P(node | state) P(state | iter),
    P(state | iter) a single hidden Markov chain.

    Parameter shapes and initialisation follow the reference (_NSSM.py:50-72): state_init [S] ~ randn,
    chain_weights [1|S, N, L] ~ randn (1 when ``restricted``), emission [S, N] ~ randn, transition [S, S] ~ randn.
    """

    _kind = "nssm"

    def __init__(self, num_states, num_chains, num_nodes, num_iters, restricted=True, use_gpu=False):
        torch.nn.Module.__init__(self)
        self.num_nodes = num_nodes
        self.num_chains = num_chains
        self.num_states = num_states
        self.num_iters = num_iters
        self.restricted = restricted
        self.use_gpu = use_gpu
        _lib._torch()  # raises without a CUDA device

        self.unnormalized_state_init = torch.nn.Parameter(torch.randn(self.num_states))
        if self.restricted:
            cw_init = torch.randn(1, self.num_nodes, self.num_chains)
        else:
            cw_init = torch.randn(self.num_states, self.num_nodes, self.num_chains)
        self.unnormalized_chain_weights = torch.nn.Parameter(cw_init)
        self.unnormalized_emission_matrix = torch.nn.Parameter(torch.randn(self.num_states, self.num_nodes))
        self.unnormalized_transition_matrix = torch.nn.Parameter(torch.randn(self.num_states, self.num_states))
        self.is_cuda = True
        self.cuda()

    def forward_model(self, want_joint=True):
        """reference _NSSM.py:76-118 -> (log P(node|iter) [I,N], joint [I,S,L,N], hidden [I,S]); also sets
        ``log_weights`` [S,N,L] (:80-82)."""
        out = super().forward_model(want_joint=want_joint)
        self.log_weights = self.log_emission_matrix.unsqueeze(-1) + self.log_chain_weights
        return out

    def _decode(self, D):
        out = super()._decode(D)
        self.log_weights = self.log_emission_matrix.unsqueeze(-1) + self.log_chain_weights
        return out

    def train(self, D=None, TPM=None, num_epochs=500, sparsity_weight=1.0, exclusivity_weight=1e-5,
              orthogonality_weight=1e-1, TPM_weight=0.0, optimizer=None, criterion=None, swa_scheduler=None,
              swa_start=200, verbose=False):
        """reference _NSSM.py:253-285 (note the smaller default exclusivity weight, :259)."""
        if D is None or isinstance(D, bool):  # nn.Module.train(mode) compatibility
            return torch.nn.Module.train(self, True if D is None else D)
        train(self, D, TPM=TPM, num_epochs=num_epochs, sparsity_weight=sparsity_weight,
              exclusivity_weight=exclusivity_weight, orthogonality_weight=orthogonality_weight, TPM_weight=TPM_weight,
              optimizer=optimizer, criterion=criterion, swa_scheduler=swa_scheduler, swa_start=swa_start,
              verbose=verbose)
