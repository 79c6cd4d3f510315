"""GPU tests for K4: DTW / Hausdorff distance matrices between trajectories against the CPU
restatement of the published dtaidistance / hausdorff algorithms (oracle/traj.py — parity unpinned, see its header),
plus the properties a distance matrix must have at sizes the oracle cannot reach."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import traj as otraj

pytestmark = pytest.mark.gpu


def _walks(C, ln, d, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((C, ln, d)).cumsum(1)


@pytest.mark.parametrize("C,ln,d", [(6, 40, 6), (5, 33, 8), (4, 64, 30), (7, 31, 1), (3, 97, 50), (2, 1, 3), (5, 70, 64)])
def test_dtw_and_hausdorff_match_oracle(C, ln, d):
    from cy2path_b200.cytopath import compute_dtw_distance_matrix, compute_Hausdorff_distance

    s = _walks(C, ln, d, seed=C * 100 + ln)
    got = compute_dtw_distance_matrix(s)
    ref = otraj.dtw_matrix(s)
    assert got.shape == (C, C) and got.dtype == np.float64
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    goth = compute_Hausdorff_distance(s)
    refh = otraj.hausdorff_matrix(s)
    assert np.allclose(goth, refh, rtol=1e-12, atol=1e-12)


def test_distance_matrix_properties_at_size():
    """200 chains x 300 steps x 30 dims (19,900 pairs): symmetry, zero diagonal, identical series -> 0, DTW bounded by
    the lock-step distance, Hausdorff bounded by DTW's largest matched gap, invariance to a common translation."""
    from cy2path_b200 import _lib

    s = _walks(200, 300, 30, seed=3)
    s[7] = s[3]  # a duplicate trajectory
    sd = torch.from_numpy(s).cuda()
    D = _lib.trajectory_distance_matrix(sd, "dtw").cpu().numpy()
    H = _lib.trajectory_distance_matrix(sd, "hausdorff").cpu().numpy()
    for M in (D, H):
        assert np.array_equal(M, M.T) and np.all(np.diag(M) == 0) and np.isfinite(M).all() and (M >= 0).all()
        assert M[3, 7] == 0.0
    lock = np.sqrt(((s[:40, None] - s[None, :40]) ** 2).sum((-1, -2)))  # lock-step distance, 40 x 40 corner
    assert (D[:40, :40] <= lock + 1e-9).all()
    assert (H <= D + 1e-9).all()  # every point is matched to some point by the warping path
    D2 = _lib.trajectory_distance_matrix(sd + 5.0, "dtw").cpu().numpy()
    assert np.allclose(D, D2, rtol=1e-9)
    sub = otraj.dtw_matrix(s[[0, 50, 199]])
    assert np.allclose(D[np.ix_([0, 50, 199], [0, 50, 199])], sub, rtol=1e-12)


def test_cluster_markov_chains_recovers_planted_lineages():
    """Chains drawn along three planted branches are separated by HDBSCAN and by Ward linkage on the DTW matrix."""
    import cy2path_b200 as cy2path

    rng = np.random.default_rng(0)
    n, d, ln, per = 600, 10, 50, 30
    emb = rng.standard_normal((n, d)) * 0.05
    branch = np.repeat(np.arange(3), n // 3)
    dirs = rng.standard_normal((3, d))
    prog = np.tile(np.linspace(0, 1, n // 3), 3)
    emb += prog[:, None] * dirs[branch] * 10
    chains = np.empty((3 * per, ln), dtype=np.int32)
    for b in range(3):
        cells = np.where(branch == b)[0]
        for c in range(per):
            chains[b * per + c] = np.sort(rng.choice(cells, ln, replace=True))

    class A:
        pass

    ad = A()
    ad.obsm = {"X_pca": emb}
    ad.uns = {"markov_chain_sampling": {"state_indices": chains, "sampling_params": {"num_chains": 3 * per}}}
    labels, model = cy2path.cluster_markov_chains(ad, method="HDBSCAN")
    truth = np.repeat(np.arange(3), per)
    for b in range(3):
        assert len(set(labels[truth == b])) == 1
    assert len(set(labels)) == 3
    labels2, model2 = cy2path.cluster_markov_chains(ad, num_lineages=3, method="linkage", distance_func="hausdorff")
    assert len(set(labels2)) == 3 and all(len(set(labels2[truth == b])) == 1 for b in range(3))
    assert model2.linkage.shape == (3 * per - 1, 4)
    labels3, _ = cy2path.cluster_markov_chains(ad, num_lineages="auto", method="linkage", max_lineages=5, differencing=True)
    assert ad.uns["cytopath"]["optimal_num_lineages"] in (2, 3, 4, 5) and len(labels3) == 3 * per
    with pytest.raises(NotImplementedError):
        cy2path.cluster_markov_chains(ad, num_lineages=3, method="kmedoids")
    with pytest.raises(ValueError):
        cy2path.cluster_markov_chains(ad, num_lineages=None, method="linkage")
