"""Mirror of cy2path/models/methods.py (reference methods.py:6-48) on top of the CUDA forward."""
from __future__ import annotations

from .. import _lib


def log_transform_params(self):
    """reference methods.py:6-25: the four log_softmax normalisations, set as attributes."""
    out = _lib.fhmm_forward(self.unnormalized_state_init, self.unnormalized_chain_weights,
                            self.unnormalized_emission_matrix, self.unnormalized_transition_matrix,
                            self.num_iters, self.restricted, want_pred=False, want_joint=False,
                            model=getattr(self, "_kind", "fhmm"))
    self.log_state_init = out["log_pi"]
    self.log_chain_weights = out["log_w"]
    self.log_emission_matrix = out["log_E"]
    self.log_transition_matrix = out["log_A"]
    return out


def compute_log_likelihood(self):
    """reference methods.py:28-40. Evaluated by the fused loss kernel (closed form in float64, csrc/fhmm.cu)."""
    import torch

    N = self.num_nodes
    zeros = torch.zeros((self.num_iters, N), dtype=torch.float32, device=self.unnormalized_state_init.device)
    plan = getattr(self, "_plan", None)
    if plan is None:
        import scipy.sparse as sp

        from ..sampling import get_plan

        self._eye = sp.identity(N, dtype="float32", format="csr")
        plan = get_plan(self._eye)
    terms, _, _ = _lib.fhmm_loss_grad(plan, self.unnormalized_state_init.data, self.unnormalized_chain_weights.data,
                                      self.unnormalized_emission_matrix.data, self.unnormalized_transition_matrix.data,
                                      zeros, self.restricted, (0.0, 0.0, 0.0, 0.0), model=getattr(self, "_kind", "fhmm"))
    self.log_likelihood = terms[2].clone()


def compute_aic(self):
    """reference methods.py:43-48."""
    self.num_params = sum(p.numel() for p in self.parameters() if p.requires_grad)
    compute_log_likelihood(self)
    self.aic = 2 * (self.num_params - self.log_likelihood).detach().cpu().numpy()
