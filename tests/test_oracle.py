"""CPU tests: pin oracle/ against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import chains as ochains
from oracle import fast as ofast
from oracle import fhmm as ofhmm
from oracle import msm as omsm


@pytest.mark.parametrize("name", ["msm_small.npz", "msm_vardeg.npz", "msm_f64matrix.npz"])
def test_msm_oracle_matches_reference_bitwise(golden, name):
    g = golden(name)
    _, hist, conv, _ = omsm.iterate_state_probability(g["T"], g["init"], g["stationary"],
                                                      max_iter=g["history"].shape[0], tol=float(g["tol"]))
    assert np.array_equal(hist, g["history"])  # same scipy arithmetic -> bit-exact float64
    assert conv == int(g["convergence"])


def test_python_and_c_spmv_restatements_match_scipy_bitwise(golden):
    g = golden("msm_vardeg.npz")
    x = g["history"][3]
    y_scipy = x @ g["T"]
    assert np.array_equal(omsm.spmv_left(x, g["T"]), y_scipy)
    assert np.array_equal(ofast.spmv_left(x, g["T"]), y_scipy)


def test_batched_oracle_equals_looped(golden):
    g = golden("msm_small.npz")
    X0 = np.stack([g["history"][0], g["history"][5], g["history"][9]], 1)
    Y = omsm.propagate_batch(g["T"], X0, 7)
    Yc = ofast.propagate_batch(g["T"], X0, 7)
    assert np.array_equal(Y[:, 0], g["history"][7])
    assert np.array_equal(Y[:, 1], g["history"][12])
    assert np.array_equal(Y, Yc)


def test_convergence_criteria_quirks():
    # no difference above tol -> None; last difference above tol -> len(eps)
    assert omsm.check_convergence_criteria(np.zeros(10), 1e-5) is None
    eps = np.zeros(10); eps[-1] = 1.0
    assert omsm.check_convergence_criteria(eps, 1e-5) == 10
    eps = np.zeros(10); eps[3] = 1.0  # d[2], d[3] > tol -> last i = 3 -> 5
    assert omsm.check_convergence_criteria(eps, 1e-5) == 5


@pytest.mark.parametrize("name", ["chains_small.npz", "chains_wide.npz", "chains_f64matrix.npz"])
def test_chain_oracle_matches_reference_bitwise(golden, name):
    g = golden(name)
    idx, cum = ochains.extract_nonzero_entries(g["T"])
    assert idx.dtype == np.int32 and cum.dtype == np.float32
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(cum.view(np.uint32), g["cum"].view(np.uint32))
    s, p = ochains.sample_chains(g["T"], idx, cum, g["init_states"], g["uniforms"])
    assert np.array_equal(s, g["states"])
    assert np.array_equal(p.view(np.uint32), g["probs"].view(np.uint32))
    s2, p2 = ofast.sample_chains(g["T"], idx, cum, g["init_states"], g["uniforms"])
    assert np.array_equal(s2, s) and np.array_equal(p2.view(np.uint32), p.view(np.uint32))


def test_chain_overshoot_raises(golden):
    g = golden("chains_small.npz")
    u = g["uniforms"][:1].copy()
    u[0, 3] = 1.5
    with pytest.raises(IndexError):
        ochains.sample_chains(g["T"], g["idx"], g["cum"], g["init_states"][:1], u)
    with pytest.raises(IndexError):
        ofast.sample_chains(g["T"], g["idx"], g["cum"], g["init_states"][:1], u)


def _params(g):
    return [torch.tensor(g[k]) for k in ("unnormalized_state_init", "unnormalized_chain_weights",
                                         "unnormalized_emission_matrix", "unnormalized_transition_matrix")]


@pytest.mark.parametrize("name", ["fhmm_restricted.npz", "fhmm_unrestricted.npz", "fhmm_single_chain.npz",
                                  "lssm_restricted.npz", "lssm_unrestricted.npz", "lssm_single_chain.npz",
                                  "nssm_restricted.npz", "nssm_unrestricted.npz", "nssm_single_chain.npz"])
def test_fhmm_oracle_matches_reference(golden, name):
    """FHMM, LSSM and NSSM (the same trainer drives both, reference trainer.py:134-196)."""
    g = golden(name)
    kind = name.split("_")[0]
    th = [t.requires_grad_(True) for t in _params(g)]
    D = torch.tensor(g["D"])
    TPM = torch.tensor(g["T"].toarray())
    out = ofhmm.loss_terms(*th, D, TPM, weights=tuple(g["weights"]), model=kind)
    assert torch.allclose(out["pred"], torch.tensor(g["pred"]), atol=1e-5, rtol=0)
    assert torch.allclose(out["hidden"], torch.tensor(g["hidden"]), atol=1e-5, rtol=0)
    assert torch.allclose(out["joint"], torch.tensor(g["joint"]), atol=1e-5, rtol=0)
    ll = ofhmm.log_likelihood(out["joint"].detach()).item()
    assert abs(ll - float(g["log_likelihood"])) <= 1e-6 * abs(float(g["log_likelihood"])) + 1e-5
    terms = dict(zip(g["term_names"].tolist(), g["term_values"].tolist()))
    for k in ("loss", "divergence", "sparsity", "orthogonality", "regularisation"):
        assert abs(out[k].item() - terms[k]) <= 1e-5 + 1e-5 * abs(terms[k]), k
    if int(g["L"]) > 1:
        assert abs(out["exclusivity"].item() - terms["exclusivity"]) <= 1e-5
    out["loss"].backward()
    names = ("unnormalized_state_init", "unnormalized_chain_weights", "unnormalized_emission_matrix",
             "unnormalized_transition_matrix")
    for t, nme in zip(th, names):
        ref = g["grad_" + nme]
        assert np.abs(t.grad.numpy() - ref).max() <= 1e-5 + 1e-4 * np.abs(ref).max(), nme


def test_trajectory_oracle_against_independent_implementations():
    """oracle/traj.py is "parity unpinned" (dtaidistance / hausdorff are not installed). What CAN be checked on this
    image: the Hausdorff restatement against scipy's own ``directed_hausdorff`` (third party, same published
    definition), and the DTW recurrence against an independent formulation — the cheapest monotone lattice path
    found by scipy's Dijkstra on the explicit DTW graph."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import dijkstra
    from scipy.spatial.distance import directed_hausdorff

    from oracle import traj as otraj

    rng = np.random.default_rng(11)
    series = rng.standard_normal((5, 9, 3)).cumsum(1)
    H = otraj.hausdorff_matrix(series)
    D = otraj.dtw_matrix(series)
    assert np.array_equal(H, H.T) and np.array_equal(D, D.T) and not H.diagonal().any() and not D.diagonal().any()
    for i in range(5):
        for j in range(i):
            want = max(directed_hausdorff(series[i], series[j])[0], directed_hausdorff(series[j], series[i])[0])
            assert abs(H[i, j] - want) <= 1e-12 * max(1.0, want)
            a, b = series[i], series[j]
            la, lb = len(a), len(b)
            cost = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
            # node (p, q) -> index p * lb + q; an edge INTO a node carries that node's cost; source = a virtual node
            rows, cols, w = [la * lb], [0], [cost[0, 0]]
            for p in range(la):
                for q in range(lb):
                    for dp, dq in ((1, 0), (0, 1), (1, 1)):
                        if p + dp < la and q + dq < lb:
                            rows.append(p * lb + q)
                            cols.append((p + dp) * lb + q + dq)
                            w.append(cost[p + dp, q + dq])
            G = sp.csr_matrix((np.asarray(w) + 1e-300, (rows, cols)), shape=(la * lb + 1, la * lb + 1))
            best = dijkstra(G, directed=True, indices=la * lb)[la * lb - 1]
            assert abs(D[i, j] - np.sqrt(best)) <= 1e-10 * max(1.0, D[i, j])
