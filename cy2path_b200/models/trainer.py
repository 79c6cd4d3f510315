"""Mirror of cy2path/models/trainer.py (reference trainer.py:13-255): the full-batch training loop.

Per epoch the reference runs forward_model, a second forward inside compute_log_likelihood, five
loss terms and ``loss.backward()`` through autograd. Here one call of cy2p_fhmm_loss_grad returns
the seven scalars and the four parameter gradients; the optimiser step stays torch's (default
RMSprop(lr=0.2), user-replaceable, trainer.py:117-119).
"""
from __future__ import annotations

import logging
import time

import numpy as np
import torch

from .. import _lib

logger = logging.getLogger("cy2path_b200")


def _plan_from_tpm(TPM, N):
    import scipy.sparse as sp

    from ..sampling import get_plan

    if isinstance(TPM, _lib.Plan):
        return TPM, None
    if TPM is None:
        raise ValueError("TPM is required (the exclusivity and TPM regularisers multiply by it, trainer.py:175,190)")
    if sp.issparse(TPM):
        M = TPM.tocsr()
    elif torch.is_tensor(TPM):
        t = TPM.detach()
        if t.is_sparse:
            t = t.coalesce().cpu()
            M = sp.csr_matrix((t.values().numpy(), t.indices().numpy()), shape=tuple(t.shape))
        else:
            M = sp.csr_matrix(t.cpu().numpy())
    else:
        M = sp.csr_matrix(np.asarray(TPM))
    if M.shape != (N, N):
        raise ValueError("TPM must be num_nodes x num_nodes")
    return get_plan(M), M


def train(self, D, TPM=None, num_epochs=300, sparsity_weight=1.0, exclusivity_weight=0.0, orthogonality_weight=0.0,
          TPM_weight=0.0, optimizer=None, criterion=None, swa_scheduler=None, swa_start=200, verbose=False):
    """reference trainer.py:13-255; same arguments, same history attributes
    (loss_values, divergence_values, likelihood_values, sparsity_values, orthogonality_values,
    exclusivity_values, regularisation_values, avg_corrcoeff, log_likelihood, elapsed_epochs)."""
    if criterion is not None and not (isinstance(criterion, torch.nn.KLDivLoss) and criterion.reduction == "batchmean"
                                      and not criterion.log_target):
        raise NotImplementedError("only the default KLDivLoss(reduction='batchmean') criterion has an analytic "
                                  "backward in the CUDA path")
    dev = self.unnormalized_state_init.device
    D = torch.as_tensor(D, dtype=torch.float32).to(dev).contiguous()
    plan, keep = _plan_from_tpm(TPM, self.num_nodes)
    self._plan, self._plan_matrix = plan, keep

    self.sparsity_weight = sparsity_weight
    self.orthogonality_weight = orthogonality_weight
    self.exclusivity_weight = exclusivity_weight
    self.TPM_weight = TPM_weight
    for name in ("loss_values", "divergence_values", "likelihood_values", "sparsity_values", "orthogonality_values",
                 "exclusivity_values", "regularisation_values"):
        if not getattr(self, name, None):
            setattr(self, name, [])
    if not getattr(self, "elapsed_epochs", None):
        self.elapsed_epochs = 0

    self.optimizer = optimizer
    if self.optimizer is None:
        self.optimizer = torch.optim.RMSprop(self.parameters(), lr=0.2)
    self.swa_scheduler = swa_scheduler
    self.criterion = criterion
    if self.criterion is None:
        self.criterion = torch.nn.KLDivLoss(reduction="batchmean", log_target=False)

    params = [self.unnormalized_state_init, self.unnormalized_chain_weights, self.unnormalized_emission_matrix,
              self.unnormalized_transition_matrix]
    weights = (sparsity_weight, exclusivity_weight, orthogonality_weight, TPM_weight)
    kind = getattr(self, "_kind", "fhmm")  # the reference's trainer is shared by its model classes
    history = torch.empty((max(num_epochs, 1), 8), dtype=torch.float32, device=dev)
    bufs = None
    start_time = time.time()
    for epoch in range(num_epochs):
        # cy2p_fhmm_loss_grad overwrites every gradient entry, so there is nothing to zero (trainer.py:130 zero_grad)
        terms, grads, bufs = _lib.fhmm_loss_grad(plan, *[p.data for p in params], D, self.restricted, weights, out=bufs,
                                                 terms=history[epoch], model=kind)
        report = epoch % 100 == 99 or epoch <= 9
        if report:  # correlation of the (pre-step) prediction with the data, trainer.py:205-210
            pred = _lib.fhmm_forward(*[p.data for p in params], self.num_iters, self.restricted, want_pred=True,
                                     model=kind)["pred"]
            out = torch.exp(pred)
            a = out - out.mean(1, keepdim=True)
            b = D - D.mean(1, keepdim=True)
            cc = (a * b).sum(1) / torch.sqrt((a * a).sum(1) * (b * b).sum(1))
            self.avg_corrcoeff = torch.mean(cc).item()
        for p, g in zip(params, grads):
            if p.requires_grad:
                if p.grad is not g:
                    p.grad = g  # the kernel writes straight into the tensors the optimiser reads
            else:
                p.grad = None
        self.optimizer.step()
        if self.elapsed_epochs > swa_start and self.swa_scheduler is not None:
            swa_scheduler.step()
        self.elapsed_epochs += 1
        if report and verbose:
            t = history[epoch].cpu().numpy()
            msg = "{:.2f}s. It {} Loss {:.2E} KL {:.2E} Likl {:.2E} Sparse {:.2E} Orth {:.2E}".format(
                time.time() - start_time, self.elapsed_epochs, t[0], t[1], t[2], t[3], t[4])
            if self.num_chains > 1:
                msg += " Exl {:.2E}".format(t[5])
            if self.TPM_weight != 0.0 and self.num_chains > 1:
                msg += " Reg {:.2E}".format(t[6])
            logger.info(msg + " Corcoef {:.2f}".format(self.avg_corrcoeff))
    if num_epochs > 0:
        h = history[:num_epochs].cpu().numpy()  # one device->host copy for the whole run
        self.loss_values += h[:, 0].tolist()
        self.divergence_values += h[:, 1].tolist()
        self.likelihood_values += h[:, 2].tolist()
        self.sparsity_values += h[:, 3].tolist()
        self.orthogonality_values += h[:, 4].tolist()
        if self.num_chains > 1:
            self.exclusivity_values += h[:, 5].tolist()
        self.regularisation_values += h[:, 6].tolist()
        self.log_likelihood = history[num_epochs - 1, 2].clone()
