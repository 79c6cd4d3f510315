"""TEST INFRASTRUCTURE — CPU (torch fp32, autograd) restatement of the reference's factorial
latent dynamic model forward pass, log-likelihood and training loss.

Not product code (see oracle/msm.py header for who may import this).
Pinned against ``tests/golden/fhmm_*.npz`` produced by the unmodified reference
(``/root/reference/cy2path/models/{_FHMM,methods,trainer}.py``): forward outputs, the per-term
losses of epoch 0 and the parameter gradients after ``loss.backward()``.

The restatement builds the time recurrence as a list + ``stack`` instead of in-place slice
writes (reference ``_FHMM.py:116-128``), which is the same forward arithmetic but keeps
autograd O(I) instead of O(I^2) so mid-size parity cases finish in seconds.
"""
import math

import torch

LSE = torch.logsumexp


def log_params(theta_pi, theta_w, theta_E, theta_A):
    """reference: cy2path/models/methods.py:6-25 (four log_softmax normalisations)."""
    log_pi = torch.log_softmax(theta_pi, dim=0)  # [S, L]   over states
    log_w = torch.log_softmax(theta_w, dim=-1)  # [L]
    log_E = torch.log_softmax(theta_E, dim=-1)  # [S, 1|L, N] over nodes
    log_A = torch.log_softmax(theta_A, dim=-1)  # [L, S, S] over destination state
    return log_pi, log_w, log_E, log_A


def forward_model(theta_pi, theta_w, theta_E, theta_A, num_iters, want_joint=True):
    """reference: cy2path/models/_FHMM.py:84-150.

    hidden[0] = log pi; hidden[t][s', l] = lse_s(hidden[t-1][s, l] + log A[l, s, s'])  (:116-124)
    joint[t, s, l, n] = log E[s, l', n] + hidden[t, s, l] + log w[l]                     (:126-141)
    pred[t, n] = lse_l lse_s joint                                                       (:144)
    Returns pred [I,N], joint [I,S,L,N] (or None), hidden [I,S,L], log A_mix [S,S]."""
    log_pi, log_w, log_E, log_A = log_params(theta_pi, theta_w, theta_E, theta_A)
    log_A_mix = LSE(log_A + log_w[:, None, None], dim=0)  # _FHMM.py:93-96
    h = [log_pi]
    for _ in range(1, num_iters):
        prev = h[-1]  # [S, L]
        # (prev^T)[l, s, None] + log A[l, s, s'] -> lse over s -> [l, s'] -> transpose
        h.append(LSE(prev.t().unsqueeze(-1) + log_A, dim=1).t())
    hidden = torch.stack(h, 0)  # [I, S, L]
    joint = log_E.unsqueeze(0) + hidden.unsqueeze(-1) + log_w[None, None, :, None]
    pred = LSE(LSE(joint, dim=1), dim=1)
    return pred, (joint if want_joint else None), hidden, log_A_mix


def lssm_forward_model(theta_pi, theta_w, theta_E, theta_A, num_iters, want_joint=True):
    """reference: cy2path/models/_LSSM.py:83-117 (LSSM: one latent chain, state-level chain weights).

    theta_pi [S], theta_w [S,L], theta_E [S,1|L,N], theta_A [S,S].
    hidden[0] = log pi; hidden[t][s'] = lse_s(hidden[t-1][s] + log A[s, s'])            (:103-107)
    joint[t, s, l, n] = log E[s, l', n] + log w[s, l] + hidden[t, s]                     (:98-111)
    Returns pred [I,N], joint [I,S,L,N] (or None), hidden [I,S], log A [S,S]."""
    log_pi = torch.log_softmax(theta_pi, dim=0)
    log_w = torch.log_softmax(theta_w, dim=-1)
    log_E = torch.log_softmax(theta_E, dim=-1)
    log_A = torch.log_softmax(theta_A, dim=-1)
    h = [log_pi]
    for _ in range(1, num_iters):
        h.append(LSE(h[-1].unsqueeze(-1) + log_A, dim=0))
    hidden = torch.stack(h, 0)  # [I, S]
    joint = (log_E + log_w.unsqueeze(-1)).unsqueeze(0) + hidden[:, :, None, None]
    pred = LSE(LSE(joint, dim=1), dim=1)
    return pred, (joint if want_joint else None), hidden, log_A


def nssm_forward_model(theta_pi, theta_w, theta_E, theta_A, num_iters, want_joint=True):
    """reference: cy2path/models/_NSSM.py:76-118 (NSSM: one latent chain, node-level chain weights).

    theta_pi [S], theta_w [1|S,N,L], theta_E [S,N], theta_A [S,S].
    log_weights[s,n,l] = log E[s,n] + log cw[s',n,l]                                     (:80-82)
    joint[t,s,n,l] = log_weights[s,n,l] + hidden[t,s], returned permuted to [I,S,L,N]    (:97-112)
    Returns pred [I,N], joint [I,S,L,N] (or None), hidden [I,S], log A [S,S]."""
    log_pi = torch.log_softmax(theta_pi, dim=0)
    log_cw = torch.log_softmax(theta_w, dim=-1)
    log_E = torch.log_softmax(theta_E, dim=-1)
    log_A = torch.log_softmax(theta_A, dim=-1)
    log_weights = log_E.unsqueeze(-1) + log_cw  # [S, N, L]
    h = [log_pi]
    for _ in range(1, num_iters):
        h.append(LSE(h[-1].unsqueeze(-1) + log_A, dim=0))
    hidden = torch.stack(h, 0)
    joint = log_weights.unsqueeze(0) + hidden[:, :, None, None]  # [I, S, N, L]
    pred = LSE(LSE(joint, dim=1), dim=-1)
    return pred, (joint.permute(0, 1, 3, 2) if want_joint else None), hidden, log_A


def log_domain_mean(x, dim=0):
    """reference: cy2path/utils.py:139-145."""
    return LSE(x, dim) - math.log(x.shape[dim])


def log_likelihood(joint):
    """reference: cy2path/models/methods.py:28-40."""
    z = (joint - LSE(joint, -1, keepdim=True)).sum(0)  # [S, L, N]
    return LSE(log_domain_mean(log_domain_mean(z)), 0)


def _kl_batchmean(log_input, target):
    """torch.nn.KLDivLoss(reduction='batchmean', log_target=False) — reference trainer.py:121-123."""
    return torch.nn.functional.kl_div(log_input, target, reduction="batchmean", log_target=False)


def loss_terms(theta_pi, theta_w, theta_E, theta_A, D, TPM, weights=(1.0, 0.1, 0.1, 0.0), model="fhmm"):
    """reference: cy2path/models/trainer.py:134-193 (one epoch's forward + loss).

    ``weights`` = (sparsity, exclusivity, orthogonality, TPM). TPM may be a dense [N,N] tensor or a
    torch sparse tensor. Returns dict of the scalar terms, 'loss', and the forward outputs."""
    I = D.shape[0]
    if model == "lssm":  # the trainer is shared between the model classes (trainer.py:134-196)
        S, L = theta_w.shape
        pred, joint, hidden, log_A_mix = lssm_forward_model(theta_pi, theta_w, theta_E, theta_A, I)
    elif model == "nssm":
        S, L = theta_E.shape[0], theta_w.shape[-1]
        pred, joint, hidden, log_A_mix = nssm_forward_model(theta_pi, theta_w, theta_E, theta_A, I)
    else:
        S, L = theta_pi.shape
        pred, joint, hidden, log_A_mix = forward_model(theta_pi, theta_w, theta_E, theta_A, I)
    sw, ew, ow, tw = weights
    out = {}
    out["divergence"] = _kl_batchmean(pred, D)  # trainer.py:139
    out["log_likelihood"] = log_likelihood(joint.detach())  # trainer.py:144
    out["sparsity"] = _kl_batchmean(torch.diag(log_A_mix), torch.ones(S))  # trainer.py:148
    M = log_domain_mean(joint, 0)  # [S, L, N]  trainer.py:153-155
    Ms = LSE(M, 1)  # [S, N]
    lngs = Ms - LSE(Ms, -1, keepdim=True)  # log P(n|s)  trainer.py:157-159
    centroid = log_domain_mean(lngs, 0)  # utils.py:149-161 (JSD, weight 1/S)
    out["orthogonality"] = (lngs.exp() * (lngs - centroid[None, :])).sum() / S

    def _mm(a, b):
        return torch.sparse.mm(b.t(), a.t()).t() if b.is_sparse else a @ b

    loss = out["divergence"] + sw * out["sparsity"] + ow * out["orthogonality"]
    if L > 1:
        Ml = LSE(M, 0)  # [L, N]
        lngc = Ml - LSE(Ml, -1, keepdim=True)  # log P(n|l)   trainer.py:166-168
        lcgn = Ml - LSE(Ml, 0, keepdim=True)  # log P(l|n)   trainer.py:170-172
        rs = _mm(lngc.exp(), TPM)  # trainer.py:175
        rsl = rs @ lcgn.t().exp()  # trainer.py:176-178
        out["exclusivity"] = _kl_batchmean(torch.diag(rsl.log()), torch.ones(L))  # trainer.py:180-182
        loss = loss + ew * out["exclusivity"]
    target = _mm(lngs.detach().exp(), TPM)  # trainer.py:188-191
    out["regularisation"] = _kl_batchmean(lngs, target)
    loss = loss + tw * out["regularisation"]
    out["loss"] = loss
    out["pred"], out["hidden"], out["joint"] = pred, hidden, joint
    return out
