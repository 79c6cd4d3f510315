"""CPU tests for the host-side mirror of cy2path/dynamics: the lazy joint_probabilities object
and the post-hoc conditionals / state selection, against fixtures produced by the unmodified reference
(tests/golden/dynamics_*.npz, make_golden.py::golden_dynamics)."""
import numpy as np
import pytest

from cy2path_b200.dynamics.infer_posthoc import compute_conditionals, latent_state_selection
from cy2path_b200.dynamics.lazy import JointProbabilities

FIXTURES = ["dynamics_fhmm.npz", "dynamics_lssm.npz", "dynamics_nssm.npz"]


class _Adata:
    def __init__(self, n):
        self.shape = (n, 0)
        self.uns = {}


def _from_golden(g, lazy):
    """adata.uns as the reference's infer_dynamics leaves it; joint_probabilities dense or refactored lazily."""
    ad = _Adata(int(g["N"]))
    J = g["model_outputs/joint_probabilities"]
    if lazy:  # recover the factors from the dense tensor: joint = exp(static) * exp(hidden)
        hid = np.log(J.sum(-1))                       # rows of exp(static) sum to r[s,l]; absorbed below
        static = np.log(J[0]) - hid[0][..., None]
        J = JointProbabilities(static, hid)
    ad.uns["latent_dynamics"] = {
        "latent_dynamics_params": {"num_states": int(g["S"]), "num_chains": int(g["L"])},
        "model_outputs": {"joint_probabilities": J, "log_smoothing": g["model_outputs/log_smoothing"],
                          "viterbi_path": g["model_outputs/viterbi_path"]},
        "conditional_probabilities": {},
    }
    return ad


def test_lazy_joint_matches_dense():
    rng = np.random.default_rng(0)
    I, S, L, N = 7, 4, 3, 11
    static = rng.standard_normal((S, L, N)).astype(np.float32) - 3
    hidden = rng.standard_normal((I, S, L)).astype(np.float32) - 1
    dense = np.exp(static[None] + hidden[..., None])
    J = JointProbabilities(static, hidden)
    assert J.shape == dense.shape and len(J) == I and J.ndim == 4
    assert np.allclose(np.asarray(J), dense, rtol=1e-6)
    assert np.allclose(J.mean(0), dense.mean(0), rtol=1e-5)
    assert np.allclose(J.sum(-1), dense.sum(-1), rtol=1e-5)
    assert np.allclose(J.sum(0), dense.sum(0), rtol=1e-5)  # generic reductions materialise
    assert np.allclose(J[2], dense[2]) and np.allclose(J[1:4], dense[1:4])
    assert np.allclose(J[3, 1], dense[3, 1]) and np.allclose(J[1:5, :, 2, 4], dense[1:5, :, 2, 4])
    assert np.allclose(J[..., 0], dense[..., 0])
    assert np.array_equal(J.sum(-1).argmax(1), dense.sum(-1).argmax(1))


@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("name", FIXTURES)
def test_state_selection_and_conditionals_match_reference(golden, name, lazy):
    g = golden(name)
    ad = _from_golden(g, lazy)
    latent_state_selection(ad, criteria="argmax_joint", min_ratio=None)
    ld = ad.uns["latent_dynamics"]
    assert np.array_equal(ld["posthoc_computations"]["selected_states"], g["selected_states"])
    for key in ("state_given_nodes", "chain_given_nodes", "chain_given_state", "chain_state_given_nodes",
                "node_given_chain_state", "node_given_state"):
        got, ref = ld["conditional_probabilities"][key + "_selected"], g["conditional_probabilities/%s_selected" % key]
        assert got.shape == ref.shape, key
        assert np.array_equal(np.isnan(got), np.isnan(ref)), key
        assert np.allclose(got, ref, rtol=2e-4 if lazy else 1e-6, atol=1e-9, equal_nan=True), key


def test_selection_criteria_and_errors(golden):
    g = golden("dynamics_fhmm.npz")
    ad = _from_golden(g, lazy=False)
    S = int(g["S"])
    latent_state_selection(ad, criteria=None, min_ratio=None)
    assert np.array_equal(ad.uns["latent_dynamics"]["posthoc_computations"]["selected_states"], np.arange(S))
    latent_state_selection(ad, criteria="viterbi", min_ratio=None)
    assert set(ad.uns["latent_dynamics"]["posthoc_computations"]["selected_states"]) <= set(range(S))
    latent_state_selection(ad, criteria="argmax_smoothing", min_ratio=0.0)
    latent_state_selection(ad, states=[0, 1], min_ratio=None)
    assert np.array_equal(ad.uns["latent_dynamics"]["posthoc_computations"]["selected_states"], [0, 1])
    # min_ratio drops states that own too few nodes
    latent_state_selection(ad, criteria=None, min_ratio=0.9)
    assert len(ad.uns["latent_dynamics"]["posthoc_computations"]["selected_states"]) <= 1
    with pytest.raises(ValueError):
        latent_state_selection(ad, criteria="nope")
    empty = _Adata(3)
    with pytest.raises(ValueError):
        latent_state_selection(empty)
    del ad.uns["latent_dynamics"]["posthoc_computations"]
    with pytest.raises(ValueError):
        compute_conditionals(ad, use_selected=True)
    compute_conditionals(ad, use_selected=False)
