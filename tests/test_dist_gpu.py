"""GPU test of the chain-sharded mode's CUDA branch in a single process (world size 1: no process group, the
all-reduce is skipped); the exchange itself is covered by the world_size-2 gloo test in test_dist_cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sample_chains_sharded_single_rank_equals_reference_golden(golden):
    from cy2path_b200 import dist as c2dist

    g = golden("chains_small.npz")
    out = c2dist.sample_chains_sharded(g["T"], g["init_states"], g["uniforms"])
    states = out["state_indices"].cpu().numpy()
    probs = out["state_transition_probabilities"].cpu().numpy()
    assert np.array_equal(states, g["states"])
    assert np.array_equal(probs.view(np.uint32), g["probs"].view(np.uint32))
    C, L = g["states"].shape
    assert out["num_chains"] == C
    ref = np.zeros((L, g["T"].shape[0]), dtype=np.float32)
    for k in range(L):  # reference simulation.py:156-162
        ind, cnt = np.unique(g["states"][:, k], return_counts=True)
        ref[k, ind] = cnt / cnt.sum()
    hist = out["state_history_max_iter"].cpu().numpy()
    assert hist.dtype == np.float32 and np.array_equal(hist.view(np.uint32), ref.view(np.uint32))
