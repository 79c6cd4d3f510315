"""GPU parity tests for K0/K2 (CDF tables, chain sampling, histogram): bit-exact vs reference golden + oracle."""
import numpy as np
import pandas as pd
import pytest

from oracle import chains as ochains
from oracle import fast as ofast

pytestmark = pytest.mark.gpu


class Duck:
    def __init__(self, T, root=None, end=None, extra=None):
        self.obsp = {"T_forward": T}
        self.shape = (T.shape[0], 0)
        cols = {}
        if root is not None:
            cols.update(root_cells=root, end_points=end)
        if extra:
            cols.update(extra)
        self.obs = pd.DataFrame(cols)
        self.uns = {}

    def copy(self):
        import copy

        return copy.deepcopy(self)


@pytest.mark.parametrize("name", ["chains_small.npz", "chains_wide.npz", "chains_f64matrix.npz"])
def test_tables_and_paths_bit_exact_vs_reference_golden(golden, name):
    import cy2path_b200 as c2p

    g = golden(name)
    ad = Duck(g["T"])
    idx, cum = c2p.extract_nonzero_entries(ad)
    assert idx.dtype == np.int32 and cum.dtype == np.float32
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(cum.view(np.uint32), g["cum"].view(np.uint32))
    C, steps = g["uniforms"].shape
    for c in range(0, C, 7):  # the reference's single-chain entry point
        s, p = c2p.iterate_markov_chain(ad, max_iter=steps + 1, init_state=g["init_states"][c],
                                        nonzero_entries=(idx, cum), uniforms=g["uniforms"][c])
        assert s.dtype == np.int32 and p.dtype == np.float32
        assert np.array_equal(s, g["states"][c])
        assert np.array_equal(p.view(np.uint32), g["probs"][c].view(np.uint32))


def test_sample_markov_chains_drop_in(golden):
    import cy2path_b200 as c2p

    g = golden("chains_small.npz")
    n = g["T"].shape[0]
    root = np.zeros(n); root[:5] = 1
    end = np.zeros(n); end[-5:] = 1
    ad = Duck(g["T"], root, end)
    C, steps = g["uniforms"].shape
    c2p.sample_markov_chains(ad, num_chains=C, max_iter=steps + 1, convergence=20, uniforms=g["uniforms"],
                             init_states=g["init_states"])
    u = ad.uns["markov_chain_sampling"]
    assert np.array_equal(u["state_indices_max_iter"], g["states"])
    assert np.array_equal(u["state_transition_probabilities_max_iter"].view(np.uint32), g["probs"].view(np.uint32))
    assert u["state_indices"].shape == (C, 20) and u["state_history"].shape == (20, n)
    ref_hist = ochains.empirical_history(g["states"], n)
    assert np.array_equal(u["state_history_max_iter"], ref_hist)
    assert set(u) == {"state_indices", "state_history", "state_indices_max_iter", "state_history_max_iter",
                      "state_transition_probabilities_max_iter", "init_state_probability",
                      "stationary_state_probability", "sampling_params"}
    assert u["sampling_params"]["convergence"] == 20 and u["sampling_params"]["num_chains"] == C


def test_larger_run_vs_c_oracle_and_edge_rows():
    """3,696-cell graph, 512 chains x 300 steps; variable degree (rows with 1 nnz), draws equal to CDF entries."""
    import cy2path_b200 as c2p
    from cy2path_b200.synth import chain_inputs, knn_velocity_tpm

    g = knn_velocity_tpm(3696, k=30, seed=1, variable_degree=True)
    T = g["T"]
    ad = Duck(T)
    idx, cum = c2p.extract_nonzero_entries(ad)
    ridx, rcum = ochains.extract_nonzero_entries(T)
    assert np.array_equal(idx, ridx) and np.array_equal(cum.view(np.uint32), rcum.view(np.uint32))
    init, uni = chain_inputs(T, g["root_cells"], 512, 301, seed=1)
    last = np.take_along_axis(cum, (np.maximum((cum > 0).sum(1), 1) - 1)[:, None], 1).ravel()
    uni *= float(last.min())
    # plant draws exactly equal to a table entry (float32 -> float64 exact): `<=` must pick that column
    uni[0, :30] = min(float(cum[init[0], 0]), float(last.min()))
    uni[1, 5] = 0.0
    ref_s, ref_p = ofast.sample_chains(T, idx, cum, init, uni)
    ad2 = Duck(T, g["root_cells"], g["end_points"])
    c2p.sample_markov_chains(ad2, num_chains=512, max_iter=301, convergence=301, uniforms=uni, init_states=init)
    u = ad2.uns["markov_chain_sampling"]
    assert np.array_equal(u["state_indices_max_iter"], ref_s)
    assert np.array_equal(u["state_transition_probabilities_max_iter"].view(np.uint32), ref_p.view(np.uint32))
    assert np.array_equal(u["state_history_max_iter"], ochains.empirical_history(ref_s, T.shape[0]))


def test_duplicate_and_unsorted_columns():
    import scipy.sparse as sp

    import cy2path_b200 as c2p

    # row 0: unsorted columns with a duplicate (col 2 twice); row order must be preserved in the tables
    indptr = np.array([0, 4, 5, 7], dtype=np.int32)
    indices = np.array([2, 0, 2, 1, 0, 1, 2], dtype=np.int32)
    data = np.array([0.25, 0.25, 0.25, 0.25, 1.0, 0.5, 0.5], dtype=np.float32)
    T = sp.csr_matrix((data, indices, indptr), shape=(3, 3))
    ad = Duck(T)
    # dense row 0 has 3 non-zeros but stores 4 entries -> the reference raises IndexError
    with pytest.raises(IndexError):
        c2p.extract_nonzero_entries(ad)
    # with a wider table (a longer row elsewhere) the reference works and sums duplicates for T[prev,next]
    indices = np.array([2, 0, 2, 1, 0, 1, 2, 3, 1, 2], dtype=np.int32)
    data = np.array([0.25, 0.25, 0.25, 0.25, 0.25, 0.25, 0.25, 0.25, 0.5, 0.5], dtype=np.float32)
    T = sp.csr_matrix((data, indices, np.array([0, 4, 8, 10, 10], dtype=np.int32)), shape=(4, 4))
    ad = Duck(T)
    idx, cum = c2p.extract_nonzero_entries(ad)
    ridx, rcum = ochains.extract_nonzero_entries(T)
    assert np.array_equal(idx, ridx) and np.array_equal(cum, rcum)
    uni = np.array([[0.6, 0.1, 0.6, 0.3]])  # from state 0: r=0.6 -> third entry (col 2, duplicate)
    s, p = c2p.iterate_markov_chain(ad, max_iter=5, init_state=0, uniforms=uni[0])
    rs, rp = ochains.sample_chains(T, idx, cum, np.array([0]), uni)
    assert np.array_equal(s, rs[0]) and np.array_equal(p, rp[0])
    assert p[0] == np.float32(0.5)  # 0.25 + 0.25


def test_overshoot_raises_index_error(golden):
    import cy2path_b200 as c2p

    g = golden("chains_small.npz")
    ad = Duck(g["T"])
    u = g["uniforms"][0].copy()
    u[3] = 1.5
    with pytest.raises(IndexError):
        c2p.iterate_markov_chain(ad, max_iter=len(u) + 1, init_state=g["init_states"][0], uniforms=u)


def test_cluster_convergence_branch():
    import cy2path_b200 as c2p
    from cy2path_b200.synth import knn_velocity_tpm

    g = knn_velocity_tpm(500, k=10, seed=9)
    lab = pd.Categorical(np.where(g["tau"] < 0.5, "early", "late"))
    ad = Duck(g["T"], g["root_cells"], g["end_points"], {"stage": lab})
    np.random.seed(0)
    c2p.sample_markov_chains(ad, num_chains=64, max_iter=40, convergence="stage")
    u = ad.uns["markov_chain_sampling"]
    assert u["cluster_proportions_max_iter"].shape == (40, 2)
    assert 2 <= u["sampling_params"]["convergence"] <= 40


def test_chain_edge_cases_single_step_and_absorbing_state():
    import scipy.sparse as sp

    import cy2path_b200 as c2p

    # state 2 is absorbing (self loop), state 3 has an empty row: walking into it overshoots on the next step
    indptr = np.array([0, 2, 4, 5, 5], dtype=np.int32)
    indices = np.array([1, 2, 2, 3, 2], dtype=np.int32)
    data = np.array([0.5, 0.5, 0.5, 0.5, 1.0], dtype=np.float32)
    T = sp.csr_matrix((data, indices, indptr), shape=(4, 4))
    ad = Duck(T)
    idx, cum = c2p.extract_nonzero_entries(ad)
    ridx, rcum = ochains.extract_nonzero_entries(T)
    assert np.array_equal(idx, ridx) and np.array_equal(cum, rcum)
    # max_iter == 1: no steps at all
    s, p = c2p.iterate_markov_chain(ad, max_iter=1, init_state=1, uniforms=np.empty(0))
    assert s.tolist() == [1] and p.shape == (0,)
    # absorbing walk
    u = np.array([0.9, 0.3, 0.99, 0.0])
    s, p = c2p.iterate_markov_chain(ad, max_iter=5, init_state=0, uniforms=u)
    rs, rp = ochains.sample_chains(T, idx, cum, np.array([0]), u[None, :])
    assert np.array_equal(s, rs[0]) and np.array_equal(p, rp[0]) and s.tolist() == [0, 2, 2, 2, 2]
    # 0 -> 1 -> 3, then the empty row raises like the reference
    with pytest.raises(IndexError):
        c2p.iterate_markov_chain(ad, max_iter=4, init_state=0, uniforms=np.array([0.2, 0.8, 0.5]))
    with pytest.raises(IndexError):
        ochains.sample_chains(T, idx, cum, np.array([0]), np.array([[0.2, 0.8, 0.5]]))
