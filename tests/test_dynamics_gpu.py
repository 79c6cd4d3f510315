"""GPU parity test for infer_dynamics / extract_model_outputs against the reference's own
``infer_dynamics`` run on a seeded model with one zero-learning-rate epoch (tests/golden/dynamics_*.npz)."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ("unnormalized_state_init", "unnormalized_chain_weights", "unnormalized_emission_matrix",
         "unnormalized_transition_matrix")


class _Adata:
    def __init__(self, g):
        self.obsp = {"T_forward": g["T"]}
        self.shape = (g["T"].shape[0], 0)
        self.obs = pd.DataFrame({"root_cells": g["root_cells"], "end_points": g["end_points"]})
        self.uns = {"state_probability_sampling": {"state_history": g["D"]}}

    def copy(self):
        import copy

        return copy.deepcopy(self)


def _run(g, **kw):
    import cy2path_b200 as cy2path
    from cy2path_b200 import models

    mode = str(g["mode"])
    model = getattr(models, mode)(int(g["S"]), int(g["L"]), int(g["N"]), int(g["I"]), restricted=bool(g["restricted"]))
    model.load_state_dict({k: torch.tensor(g[k]) for k in NAMES})
    ad = _Adata(g)
    out = cy2path.infer_dynamics(ad, model=model, num_epochs=1, mode=mode, save_model=None,
                                 optimizer=torch.optim.SGD(model.parameters(), lr=0.0), **kw)
    return ad, out, model


@pytest.mark.parametrize("name", ["dynamics_fhmm.npz", "dynamics_lssm.npz", "dynamics_nssm.npz"])
def test_infer_dynamics_matches_reference(golden, name):
    g = golden(name)
    ad, out, model = _run(g)
    assert out is model
    ld = ad.uns["latent_dynamics"]
    assert abs(ld["log_likelihood"] - float(g["log_likelihood"])) <= 1e-6 * abs(float(g["log_likelihood"])) + 1e-5
    assert ld["latent_dynamics_params"]["mode"] == str(g["mode"])
    assert ld["latent_dynamics_params"]["num_states"] == int(g["S"])
    checked = 0
    for key in g:
        if "/" not in key or key.endswith("_selected"):
            continue
        grp, k = key.split("/")
        got, ref = np.asarray(ld[grp][k]), g[key]
        assert got.shape == ref.shape, key
        if k == "viterbi_path":
            assert np.array_equal(got, ref), key
        elif k.startswith("log_"):
            assert np.abs(got - ref).max() <= 1e-5 + 2e-6 * np.abs(ref).max(), key
        else:  # probabilities
            assert np.allclose(got, ref, rtol=2e-4, atol=1e-7), key
        checked += 1
    assert checked >= 16
    # the reference's post-hoc step runs on these outputs
    from cy2path_b200.dynamics import latent_state_selection

    latent_state_selection(ad, criteria="argmax_joint", min_ratio=None)
    assert np.array_equal(ld["posthoc_computations"]["selected_states"], g["selected_states"])
    ref = g["conditional_probabilities/state_given_nodes_selected"]
    assert np.allclose(ld["conditional_probabilities"]["state_given_nodes_selected"], ref, rtol=2e-4, atol=1e-7)


def test_lazy_joint_and_copy_and_errors(golden):
    from cy2path_b200.dynamics import JointProbabilities, extract_model_outputs

    g = golden("dynamics_fhmm.npz")
    _, (model, ad2), _ = _run(g, copy=True)
    assert "latent_dynamics" in ad2.uns
    dense = ad2.uns["latent_dynamics"]["model_outputs"]["joint_probabilities"]
    assert isinstance(dense, np.ndarray)
    extract_model_outputs(ad2, model, dense_limit_bytes=0)  # force the lazy representation
    lazy = ad2.uns["latent_dynamics"]["model_outputs"]["joint_probabilities"]
    assert isinstance(lazy, JointProbabilities) and lazy.shape == dense.shape
    assert np.allclose(np.asarray(lazy), dense, rtol=1e-5, atol=1e-9)
    assert np.allclose(lazy.mean(0), dense.mean(0), rtol=1e-4, atol=1e-9)

    import cy2path_b200 as cy2path

    bad = _Adata(g)
    bad.uns = {}
    with pytest.raises(ValueError):
        cy2path.infer_dynamics(bad, num_epochs=1, save_model=None)
    with pytest.raises(ValueError):
        cy2path.infer_dynamics(_Adata(g), mode="XYZ", num_epochs=1, save_model=None)


def test_infer_dynamics_trains_from_scratch(tmp_path):
    """The public call a user makes: sample_state_probability -> infer_dynamics (all three modes)."""
    import cy2path_b200 as cy2path
    from cy2path_b200.synth import knn_velocity_tpm

    g = knn_velocity_tpm(1500, k=12, seed=9)

    class A(_Adata):
        def __init__(self):
            self.obsp = {"T_forward": g["T"]}
            self.shape = (1500, 0)
            self.obs = pd.DataFrame({"root_cells": g["root_cells"], "end_points": g["end_points"]})
            self.uns = {}

    for mode in ("FHMM", "LSSM", "NSSM"):
        ad = A()
        stat = g["end_points"] / g["end_points"].sum()
        x0 = g["root_cells"] / g["root_cells"].sum()
        _, hist, _ = cy2path.iterate_state_probability(ad, init=x0, stationary=stat, max_iter=40, tol=1e-12)
        ad.uns["state_probability_sampling"] = {"state_history": hist}
        torch.manual_seed(5)
        path = str(tmp_path / ("model_" + mode))
        model = cy2path.infer_dynamics(ad, num_states=6, num_chains=2, num_epochs=25, mode=mode, save_model=path)
        assert type(model).__name__ == mode and len(model.loss_values) == 25
        assert model.loss_values[-1] < model.loss_values[0]
        mo = ad.uns["latent_dynamics"]["model_outputs"]
        assert mo["predicted_state_history"].shape == (40, 1500)
        assert np.allclose(mo["predicted_state_history"].sum(1), 1.0, atol=1e-3)
        assert mo["viterbi_path"].shape == (40, 2)
        sd = torch.load(path)
        assert set(sd) == set(NAMES)
