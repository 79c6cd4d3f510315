"""GPU tests at BASELINE.json's full shapes (config 2: 50,000 cells, 4,096 starts, 100,000 chains; config 3:
S=10, L=4, N=50,000) through size-independent properties, plus spot checks of a few columns / chains against
the oracle. The oracle cannot run these sizes in full (config 2 would take ~2 h on one core)."""
import numpy as np
import pytest

from oracle import chains as ochains
from oracle import fast as ofast
from oracle import msm as omsm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def graph50k():
    from cy2path_b200.synth import knn_velocity_tpm

    return knn_velocity_tpm(50_000, k=30, seed=2)


def test_config2_propagation_properties(graph50k):
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import batched_starts, stationary_by_power

    T = graph50k["T"]
    plan = get_plan(T)
    B, iters = 4096, 24
    X0h = batched_starts(T, B, seed=2)
    X0 = torch.from_numpy(X0h).cuda()
    stat = stationary_by_power(T, 50)
    out = _lib.propagate(plan, X0, iters, stationary=stat, want_eps=True)
    X = out["final"]
    # mass conservation and non-negativity of every start distribution
    mass = X.double().sum(0)
    assert float((mass - 1).abs().max()) <= 1e-5
    assert float(X.min()) >= 0.0
    # eps in [0, 2], finite, and equal to the cosine distance recomputed from the final iterate
    eps = out["eps"]
    assert torch.isfinite(eps).all() and float(eps.min()) >= 0 and float(eps.max()) <= 2
    pi = torch.from_numpy(stat).cuda()
    cos = 1 - (X.double() * pi[:, None]).sum(0) / torch.sqrt((X.double() ** 2).sum(0) * (pi ** 2).sum())
    # (partial sums inside a CTA - 64 rows - are fp32, every sum across CTAs is fp64: agreement to a few 1e-8)
    assert float((eps[-1] - cos.clamp(0, 2)).abs().max()) <= 5e-7
    # batched == the same columns propagated on their own (different tiling / launch shapes)
    cols = [0, 1, 2047, 2048, 4095]
    sub = _lib.propagate(plan, X0[:, cols].contiguous(), iters)["final"]
    assert float((sub - X[:, cols]).abs().max()) <= 1e-7
    # linearity in the start distribution
    mix = (0.25 * X0[:, 10] + 0.75 * X0[:, 20]).reshape(-1, 1).repeat(1, 4).contiguous()
    pm = _lib.propagate(plan, mix, iters)["final"][:, 0]
    assert float((pm - (0.25 * X[:, 10] + 0.75 * X[:, 20])).abs().max()) <= 1e-6
    # three columns against the float64 oracle
    ref = ofast.propagate_batch(T, X0h[:, [3, 1000, 4000]], iters)
    got = X[:, [3, 1000, 4000]].cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-5 and omsm.rel_l1(got, ref) <= 1e-6


def test_config2_chain_sampling_properties(graph50k):
    import torch

    from cy2path_b200.sampling import get_plan
    from cy2path_b200.simulation import _tables_device, _walk
    from cy2path_b200.synth import chain_inputs

    T = graph50k["T"]
    n = T.shape[0]
    plan = get_plan(T)
    idx_d, cum_d = _tables_device(plan)
    C, max_iter = 100_000, 120
    init, uni = chain_inputs(T, graph50k["root_cells"], C, max_iter, seed=2)
    uni *= float(cum_d.max(1).values.min())
    states, probs, hist = _walk(plan, idx_d, cum_d, init, uni, want_history_steps=max_iter)
    assert states.shape == (C, max_iter) and probs.shape == (C, max_iter - 1)
    assert int(states.min()) >= 0 and int(states.max()) < n
    assert torch.equal(states[:, 0].cpu(), torch.from_numpy(init))
    # every step is an edge of T and the recorded probability is that entry of T
    Td = torch.sparse_csr_tensor(torch.from_numpy(T.indptr.astype(np.int64)), torch.from_numpy(T.indices.astype(np.int64)),
                                 torch.from_numpy(T.data), size=(n, n)).cuda()
    prev, nxt = states[:, :-1].reshape(-1).long(), states[:, 1:].reshape(-1).long()
    sel = torch.randperm(prev.numel(), device="cuda")[:200_000]
    rows = Td.crow_indices()
    colv, val = Td.col_indices(), Td.values()
    # position of nxt within row prev (rows are sorted): binary search per sampled transition
    lo, hi = rows[prev[sel]], rows[prev[sel] + 1]
    width = int((hi - lo).max())
    offs = torch.arange(width, device="cuda")[None, :]
    pos = (lo[:, None] + offs).clamp(max=colv.numel() - 1)
    match = (colv[pos] == nxt[sel][:, None]) & (offs < (hi - lo)[:, None])
    assert bool(match.any(1).all()), "a sampled transition is not an edge of T"
    pv = (val[pos] * match).sum(1)
    assert torch.equal(pv, probs.reshape(-1)[sel])
    # the per-step histogram is a distribution over cells
    assert float((hist.double().sum(1) - 1).abs().max()) <= 1e-5
    # a few chains against the C oracle, bit for bit
    idx, cum = idx_d.cpu().numpy(), cum_d.cpu().numpy()
    pick = [0, 1, 777, 99_999]
    rs, rp = ofast.sample_chains(T, idx, cum, init[pick], uni[pick])
    assert np.array_equal(states[pick].cpu().numpy(), rs)
    assert np.array_equal(probs[pick].cpu().numpy().view(np.uint32), rp.view(np.uint32))


def test_config3_fhmm_gradient_properties(graph50k):
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.models import FHMM
    from cy2path_b200.sampling import get_plan

    T = graph50k["T"]
    N, I, S, L = T.shape[0], 120, 10, 4
    plan = get_plan(T)
    x0 = torch.tensor((graph50k["root_cells"] / graph50k["root_cells"].sum()).astype(np.float32)).cuda()
    D = _lib.propagate(plan, x0, I, history_stride=1, want_final=False)["history"].contiguous()
    torch.manual_seed(3)
    m = FHMM(S, L, N, I, restricted=True, use_gpu=True)
    terms, (g_pi, g_w, g_E, g_A), _ = _lib.fhmm_loss_grad(plan, *[p.data for p in m.parameters()], D, True,
                                                         (1.0, 0.1, 0.1, 0.05))
    t = terms.cpu().numpy()
    assert np.isfinite(t).all() and t[1] > 0  # KL divergence is positive
    # gradients w.r.t. the logits of a softmax sum to zero over the normalised axis
    assert float(g_E.double().sum(-1).abs().max()) <= 1e-4
    assert float(g_pi.double().sum(0).abs().max()) <= 1e-5
    assert float(g_A.double().sum(-1).abs().max()) <= 1e-5
    assert abs(float(g_w.double().sum())) <= 1e-5
    # pred rows are log-distributions over the cells; hidden columns are log-distributions over the states
    pred, _, hidden = m.forward_model(want_joint=False)
    assert float((pred.double().exp().sum(1) - 1).abs().max()) <= 1e-4
    assert float((hidden.double().exp().sum(1) - 1).abs().max()) <= 1e-5
    # finite-difference check of dLoss/dtheta_w (4 numbers) with the same kernel
    base = float(t[0])
    eps = 1e-2
    for l in range(L):
        pw = [p.data.clone() for p in m.parameters()]
        pw[1][l] += eps
        up = float(_lib.fhmm_loss_grad(plan, *pw, D, True, (1.0, 0.1, 0.1, 0.05))[0][0])
        pw[1][l] -= 2 * eps
        dn = float(_lib.fhmm_loss_grad(plan, *pw, D, True, (1.0, 0.1, 0.1, 0.05))[0][0])
        fd = (up - dn) / (2 * eps)
        assert abs(fd - float(g_w[l])) <= 5e-3 + 0.05 * abs(float(g_w[l])), (l, fd, float(g_w[l]), base)
    # a few epochs of training reduce the loss
    m.train(D, TPM=T, num_epochs=30)
    assert m.loss_values[-1] < m.loss_values[0]
