"""Mirror of cy2path/models/_LSSM.py — same constructor, parameters and methods (reference _LSSM.py:7-287)."""
from __future__ import annotations

import torch

from .. import _lib
from ._FHMM import FHMM


class LSSM(FHMM):
    """Latent dynamic model with a common latent state space and state-level chain weights, trained on a Markov
    state simulation of an observed-state TPM (reference _LSSM.py:7-32).

    P(node | iter) = sum_chain sum_state P(node | chain, state) P(chain | state) P(state | iter),
    P(state | iter) a single hidden Markov chain.

    Parameter shapes and initialisation follow the reference (_LSSM.py:50-76): state_init [S] ~ randn,
    chain_weights [S, L] = ones, emission [S, 1|L, N] ~ randn, transition [S, S] ~ randn, drawn in the same order
    from the CPU generator. The decoders, ``train`` and the log-likelihood helpers are inherited: the C ABI
    has the same call shapes for both model classes (include/cy2path_b200.h, K3b).
    """

    _kind = "lssm"

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
        self.unnormalized_chain_weights = torch.nn.Parameter(torch.ones(self.num_states, self.num_chains))
        if self.restricted:
            em_init = torch.randn(self.num_states, 1, self.num_nodes)
        else:
            em_init = torch.randn(self.num_states, self.num_chains, self.num_nodes)
        self.unnormalized_emission_matrix = torch.nn.Parameter(em_init)
        self.unnormalized_transition_matrix = torch.nn.Parameter(torch.randn(self.num_states, self.num_states))
        self.is_cuda = True
        self.cuda()
