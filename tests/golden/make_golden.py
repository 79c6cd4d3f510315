"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes small .npz fixtures next to this file (or into $CY2P_GOLDEN_OUT: tests/test_golden_reproducible.py
regenerates them there and compares with the committed files). The reference has no golden vectors or tests of
its own; these fixtures are outputs of its own functions on seeded synthetic
inputs and are what pins ``oracle/`` (tests/test_oracle.py) and the CUDA path (tests/test_*_gpu.py).
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

SRC = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(SRC))
HERE = os.environ.get("CY2P_GOLDEN_OUT", SRC)  # where the fixtures are written
sys.path.insert(0, ROOT)

from cy2path_b200.synth import knn_velocity_tpm  # noqa: E402
from oracle import ref_shim  # noqa: E402


def adata_for(g, ref):
    return ref.DuckAnnData(g["T"], pd.DataFrame({"root_cells": g["root_cells"], "end_points": g["end_points"]}))


def csr_fields(T, prefix="T_"):
    return {prefix + "indptr": T.indptr, prefix + "indices": T.indices, prefix + "data": T.data,
            prefix + "n": np.int64(T.shape[0])}


def float64_matrix(g, seed):
    """The same graph with float64 transition probabilities that are NOT float32-representable (scvelo writes float32,
    but nothing in the reference forces it: sampling.py:40-42 then multiplies in float64 and simulation.py:24-25
    accumulates the CDF in float64)."""
    import scipy.sparse as sp

    T = g["T"].astype(np.float64)
    T.data = T.data * (1.0 + 0.25 * np.random.default_rng(seed + 7).random(T.nnz))
    T = sp.csr_matrix(sp.diags(1.0 / np.asarray(T.sum(1)).ravel()) @ T)
    T.sort_indices()
    assert T.dtype == np.float64 and not np.array_equal(T.data.astype(np.float32).astype(np.float64), T.data)
    g = dict(g)
    g["T"] = T
    return g


def golden_msm(ref, name, n, k, seed, max_iter, variable_degree=False, init="root_cells", tol=1e-5, float64_T=False):
    g = knn_velocity_tpm(n, k=k, seed=seed, variable_degree=variable_degree)
    if float64_T:
        g = float64_matrix(g, seed)
    ad = adata_for(g, ref)
    init_p, _ = ref.utils.check_root_init(ad, init=init)
    stat = g["end_points"] / g["end_points"].sum()  # the reference's own fall-back (utils.py:110-112)
    sh, shm, conv = ref.sampling.iterate_state_probability(ad, init=init_p, stationary=stat, max_iter=max_iter, tol=tol)
    np.savez_compressed(os.path.join(HERE, name), **csr_fields(g["T"]), init=np.asarray(init_p, dtype=np.float64),
                        stationary=stat, history=shm, convergence=np.int64(conv), tol=np.float64(tol),
                        root_cells=g["root_cells"], end_points=g["end_points"])
    print(name, shm.shape, "conv", conv)


def golden_chains(ref, name, n, k, seed, C, max_iter, variable_degree=False, float64_T=False):
    g = knn_velocity_tpm(n, k=k, seed=seed, variable_degree=variable_degree)
    if float64_T:
        g = float64_matrix(g, seed)
    ad = adata_for(g, ref)
    idx, cum = ref.simulation.extract_nonzero_entries(ad)
    rng = np.random.default_rng(seed + 100)
    init_states = rng.choice(n, C).astype(np.int32)
    uniforms = rng.random((C, max_iter - 1))
    # rounding to 3 decimals can leave the last CDF entry below 1 (reference FIXME simulation.py:48);
    # keep supplied draws under the row minimum of the final CDF value so the reference itself runs
    rows_last = np.take_along_axis(cum, (np.maximum((cum > 0).sum(1), 1) - 1)[:, None], 1).ravel()
    uniforms *= float(rows_last.min())
    states = np.empty((C, max_iter), dtype=np.int32)
    probs = np.empty((C, max_iter - 1), dtype=np.float32)
    real_rand = np.random.rand
    try:
        for c in range(C):
            np.random.rand = lambda m, _c=c: uniforms[_c].copy()
            s, p = ref.simulation.iterate_markov_chain(ad, max_iter=max_iter, init_state=init_states[c],
                                                       nonzero_entries=(idx, cum))
            states[c], probs[c] = s, p
    finally:
        np.random.rand = real_rand
    np.savez_compressed(os.path.join(HERE, name), **csr_fields(g["T"]), idx=idx, cum=cum, init_states=init_states,
                        uniforms=uniforms, states=states, probs=probs)
    print(name, idx.shape, states.shape)


def golden_fhmm(ref, name, S, L, N, I, restricted, seed, k=6, model_cls="FHMM"):
    g = knn_velocity_tpm(N, k=k, seed=seed)
    ad = adata_for(g, ref)
    init_p, _ = ref.utils.check_root_init(ad, init="root_cells")
    stat = g["end_points"] / g["end_points"].sum()
    _, shm, _ = ref.sampling.iterate_state_probability(ad, init=init_p, stationary=stat, max_iter=I)
    D = torch.Tensor(shm)  # float32, as infer_dynamics.py:207-211
    TPM = torch.Tensor(g["T"].toarray())  # dense, as infer_dynamics.py:204
    cls = getattr(ref.models, model_cls)

    def make():
        torch.manual_seed(seed)
        m = cls(S, L, N, I, restricted=restricted, use_gpu=False)
        if model_cls == "LSSM":  # the reference initialises the chain weights to ones (_LSSM.py:59-64);
            with torch.no_grad():  # randomise them so the fixture exercises the weight gradient
                m.unnormalized_chain_weights.copy_(torch.randn(S, L))
        return m

    model = make()
    params0 = {k_: v.detach().clone().numpy() for k_, v in model.state_dict().items()}
    with torch.no_grad():
        pred, joint, hidden = model.forward_model()
        ref.methods.compute_log_likelihood(model)
        ll = model.log_likelihood.item()
    # one epoch with a zero-learning-rate optimiser: parameters stay put, .grad survives
    model.train(D, TPM=TPM, num_epochs=1, sparsity_weight=1.0, exclusivity_weight=0.1, orthogonality_weight=0.1,
                TPM_weight=0.05,  # explicit: NSSM.train defaults to exclusivity_weight=1e-5 (_NSSM.py:259)
                optimizer=torch.optim.SGD(model.parameters(), lr=0.0))
    grads = {"grad_" + k_: p.grad.detach().clone().numpy() for k_, p in model.named_parameters()}
    terms = dict(loss=model.loss_values[0], divergence=model.divergence_values[0],
                 likelihood=model.likelihood_values[0], sparsity=model.sparsity_values[0],
                 orthogonality=model.orthogonality_values[0], regularisation=model.regularisation_values[0],
                 exclusivity=(model.exclusivity_values[0] if L > 1 else np.nan))
    # three default-optimiser epochs from the same start (trainer drop-in trajectory)
    model2 = make()
    model2.train(D, TPM=TPM, num_epochs=3)
    traj = np.array([model2.loss_values, model2.divergence_values, model2.likelihood_values,
                     model2.sparsity_values, model2.orthogonality_values,
                     model2.exclusivity_values if L > 1 else [np.nan] * 3, model2.regularisation_values])
    params3 = {"after3_" + k_: v.detach().clone().numpy() for k_, v in model2.state_dict().items()}
    # decoders (next-row f-1) on the trained model
    with torch.no_grad():
        model2.forward_model()
        la, lp = model2.filtering(D)
        lb, lg = model2.smoothing(D)
        ld, psi, lmax, best = model2.viterbi(D)
    np.savez_compressed(
        os.path.join(HERE, name), **csr_fields(g["T"]), D=D.numpy(), S=S, L=L, N=N, I=I, restricted=restricted,
        **params0, **grads, **params3, pred=pred.numpy(), joint=joint.numpy(), hidden=hidden.numpy(),
        log_likelihood=np.float64(ll), term_names=np.array(list(terms.keys())),
        term_values=np.array(list(terms.values()), dtype=np.float64), weights=np.array([1.0, 0.1, 0.1, 0.05]),
        traj=traj, log_alpha=la.numpy(), log_probs=lp.numpy(), log_beta=lb.numpy(), log_gamma=lg.numpy(),
        log_delta=ld.numpy(), psi=psi.numpy(), log_max=lmax.numpy(), best_path=np.array(best, dtype=np.float64))
    print(name, "loss", terms["loss"], "ll", ll)


def golden_dynamics(ref, name, mode, S, L, N, I, restricted, seed, k=6):
    """infer_dynamics + latent_state_selection of the reference on a seeded model, one zero-learning-rate epoch
    (the parameters stay at their initial values, so the stored outputs are functions of params0 only)."""
    import importlib

    dyn = importlib.import_module("cy2path.dynamics.infer_dynamics")
    post = importlib.import_module("cy2path.dynamics.infer_posthoc")
    g = knn_velocity_tpm(N, k=k, seed=seed)
    ad = adata_for(g, ref)
    init_p, _ = ref.utils.check_root_init(ad, init="root_cells")
    stat = g["end_points"] / g["end_points"].sum()
    _, shm, _ = ref.sampling.iterate_state_probability(ad, init=init_p, stationary=stat, max_iter=I)
    ad.uns["state_probability_sampling"] = {"state_history": shm}
    torch.manual_seed(seed)
    model = getattr(ref.models, mode)(S, L, N, I, restricted=restricted, use_gpu=False)
    params0 = {k_: v.detach().clone().numpy() for k_, v in model.state_dict().items()}
    dyn.infer_dynamics(ad, model=model, num_epochs=1, mode=mode, save_model=None,
                       optimizer=torch.optim.SGD(model.parameters(), lr=0.0))
    ld = ad.uns["latent_dynamics"]
    out = {}
    for grp in ("model_outputs", "model_params", "conditional_probabilities"):
        for key, val in ld[grp].items():
            out[grp + "/" + key] = np.asarray(val)
    post.latent_state_selection(ad, criteria="argmax_joint", min_ratio=None)
    for key, val in ld["conditional_probabilities"].items():
        if key.endswith("_selected"):
            out["conditional_probabilities/" + key] = np.asarray(val)
    out["selected_states"] = np.asarray(ld["posthoc_computations"]["selected_states"])
    np.savez_compressed(os.path.join(HERE, name), **csr_fields(g["T"]), D=shm, S=S, L=L, N=N, I=I,
                        restricted=restricted, mode=mode, root_cells=g["root_cells"], end_points=g["end_points"],
                        log_likelihood=np.float64(ld["log_likelihood"]), **params0, **out)
    print(name, sorted(out)[:3], "...", len(out), "arrays")


def main():
    """``python make_golden.py [prefix ...]`` regenerates every fixture, or those whose file name starts with a prefix."""
    import sys

    ref = ref_shim.load()
    jobs = [
        (golden_msm, "msm_small.npz", dict(n=400, k=10, seed=11, max_iter=120)),
        (golden_msm, "msm_vardeg.npz", dict(n=300, k=12, seed=12, max_iter=80, variable_degree=True, init="uniform")),
        (golden_msm, "msm_f64matrix.npz", dict(n=260, k=9, seed=14, max_iter=140, variable_degree=True, float64_T=True)),
        (golden_chains, "chains_small.npz", dict(n=400, k=10, seed=11, C=48, max_iter=60)),
        (golden_chains, "chains_f64matrix.npz", dict(n=260, k=9, seed=14, C=24, max_iter=50, variable_degree=True,
                                                     float64_T=True)),
        (golden_chains, "chains_wide.npz", dict(n=300, k=40, seed=13, C=32, max_iter=40)),  # W > 32: multi-chunk search
        (golden_fhmm, "fhmm_restricted.npz", dict(S=4, L=3, N=60, I=25, restricted=True, seed=21)),
        (golden_fhmm, "fhmm_unrestricted.npz", dict(S=3, L=2, N=50, I=20, restricted=False, seed=22)),
        (golden_fhmm, "fhmm_single_chain.npz", dict(S=5, L=1, N=40, I=15, restricted=True, seed=23)),
        (golden_fhmm, "lssm_restricted.npz", dict(S=4, L=3, N=60, I=25, restricted=True, seed=31, model_cls="LSSM")),
        (golden_fhmm, "lssm_unrestricted.npz", dict(S=3, L=2, N=50, I=20, restricted=False, seed=32, model_cls="LSSM")),
        (golden_fhmm, "lssm_single_chain.npz", dict(S=5, L=1, N=40, I=15, restricted=True, seed=33, model_cls="LSSM")),
        (golden_fhmm, "nssm_restricted.npz", dict(S=4, L=3, N=60, I=25, restricted=True, seed=41, model_cls="NSSM")),
        (golden_fhmm, "nssm_unrestricted.npz", dict(S=3, L=2, N=50, I=20, restricted=False, seed=42, model_cls="NSSM")),
        (golden_fhmm, "nssm_single_chain.npz", dict(S=5, L=1, N=40, I=15, restricted=True, seed=43, model_cls="NSSM")),
        (golden_dynamics, "dynamics_fhmm.npz", dict(mode="FHMM", S=4, L=3, N=60, I=25, restricted=True, seed=51)),
        (golden_dynamics, "dynamics_lssm.npz", dict(mode="LSSM", S=3, L=2, N=50, I=20, restricted=False, seed=52)),
        (golden_dynamics, "dynamics_nssm.npz", dict(mode="NSSM", S=4, L=2, N=40, I=15, restricted=True, seed=53)),
    ]
    for fn, name, kw in jobs:
        if len(sys.argv) > 1 and not any(name.startswith(p) for p in sys.argv[1:]):
            continue
        fn(ref, name, **kw)


if __name__ == "__main__":
    main()
