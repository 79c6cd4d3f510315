"""Drop-in mirror of cy2path/models — the factorial latent dynamic model on the B200.

``FHMM`` keeps the reference's constructor, parameter names/shapes (state_dict compatible),
``forward_model`` / ``train`` methods and training-history attributes; the arithmetic runs in
csrc/fhmm.cu. LSSM / NSSM and the decoders' CUDA kernels are the next rows of SURVEY.md §8(f).
"""
from ._FHMM import FHMM  # noqa: F401

__all__ = ["FHMM"]


def _smoke(T):
    """Tiny forward + loss/grad on cuda:0, checked against the oracle (called by __graft_entry__.smoke)."""
    import numpy as np
    import torch

    from oracle import fhmm as ofhmm

    from .. import _lib
    from ..sampling import get_plan

    N = T.shape[0]
    I, S, L = 12, 4, 3
    plan = get_plan(T)
    x = torch.zeros(N, dtype=torch.float32, device="cuda")
    x[:7] = 1.0 / 7
    D = _lib.propagate(plan, x, I, history_stride=1, want_final=False)["history"].contiguous()
    torch.manual_seed(0)
    model = FHMM(S, L, N, I, restricted=True, use_gpu=True)
    th = [p.detach().cpu().clone().requires_grad_(True) for p in model.parameters()]
    import scipy.sparse as sp

    coo = sp.coo_matrix(T)
    Tt = torch.sparse_coo_tensor(np.vstack([coo.row, coo.col]), coo.data.astype(np.float32), (N, N)).coalesce()
    ref = ofhmm.loss_terms(*th, D.cpu(), Tt, weights=(1.0, 0.1, 0.1, 0.05))
    ref["loss"].backward()
    terms, grads, _ = _lib.fhmm_loss_grad(plan, *[p.data for p in model.parameters()], D, True, (1.0, 0.1, 0.1, 0.05))
    t = terms.cpu().numpy()
    assert abs(t[0] - ref["loss"].item()) <= 1e-4 * max(1.0, abs(ref["loss"].item())), (t[0], ref["loss"].item())
    for g, r in zip(grads, th):
        rg = r.grad.numpy()
        assert np.abs(g.cpu().numpy() - rg).max() <= 1e-5 + 1e-3 * np.abs(rg).max()
