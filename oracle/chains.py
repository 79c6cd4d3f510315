"""TEST INFRASTRUCTURE — CPU restatement of the reference's Markov-chain sampler.

Not product code (see oracle/msm.py header for who may import this).
Pinned against ``tests/golden/chains_*.npz`` produced by the unmodified reference
(``/root/reference/cy2path/simulation.py``) with ``numpy.random.rand`` patched to return
the supplied uniforms. numpy 2.3 promotion rules define the comparison dtype
(float64 scalar <= float32 row -> float64, NEP 50); see SURVEY.md §8a-a5.
"""
import numpy as np


def extract_nonzero_entries(T):
    """reference: cy2path/simulation.py:15-53.

    idx int32 [n, W], cum float32 [n, W]; per-row sequential float32 cumsum (simulation.py:24-25),
    zero padding, W = max per-row count of non-zeros of the *dense* matrix (simulation.py:28-41),
    then np.round(cum, 3) in float32 (simulation.py:49-51)."""
    n = T.shape[0]
    indptr, indices, data = T.indptr, T.indices, T.data
    Tc = T.copy()
    Tc.sum_duplicates()
    Tc.eliminate_zeros()
    W = int(np.diff(Tc.indptr).max())  # == max(np.count_nonzero(T.toarray(), axis=1))
    idx = np.zeros((n, W), dtype=np.int32)
    cum = np.zeros((n, W), dtype=np.float32)
    for i in range(n):
        s, e = indptr[i], indptr[i + 1]
        if e - s > W:
            raise IndexError(f"index {W} is out of bounds for axis 1 with size {W}")
        idx[i, : e - s] = indices[s:e]
        cum[i, : e - s] = np.cumsum(data[s:e])  # float32 in, float32 sequential adds
    cum = np.round(cum, decimals=3)
    return idx, cum


def iterate_markov_chain(T, idx, cum, init_state, uniforms):
    """reference: cy2path/simulation.py:57-87 with ``random_step`` supplied.

    j = first index with r <= cum[state, j] (float64 compare); IndexError when none.
    The recorded probability is T[prev, next] (scipy scalar indexing sums duplicates)."""
    steps = len(uniforms)
    states = np.empty(steps + 1, dtype=np.int32)
    probs = np.empty(steps, dtype=np.float32)
    states[0] = init_state
    indptr, indices, data = T.indptr, T.indices, T.data
    for k in range(1, steps + 1):
        prev = states[k - 1]
        hit = np.where(uniforms[k - 1] <= cum[prev, :])[0]
        j = hit[0]  # IndexError on CDF overshoot, as in the reference
        nxt = idx[prev, j]
        states[k] = nxt
        s, e = indptr[prev], indptr[prev + 1]
        m = indices[s:e] == nxt
        probs[k - 1] = data[s:e][m].sum() if m.any() else 0.0
    return states, probs


def sample_chains(T, idx, cum, init_states, uniforms):
    """Fan-out of iterate_markov_chain (reference: simulation.py:133-153)."""
    C, steps = uniforms.shape
    states = np.empty((C, steps + 1), dtype=np.int32)
    probs = np.empty((C, steps), dtype=np.float32)
    for c in range(C):
        states[c], probs[c] = iterate_markov_chain(T, idx, cum, init_states[c], uniforms[c])
    return states, probs


def empirical_history(states, n):
    """reference: cy2path/simulation.py:156-162 — per-step histogram, counts / counts.sum().

    The reference scatters into ``np.empty`` (unvisited entries are uninitialised); here they
    are zero, and parity is asserted on visited entries."""
    C, steps1 = states.shape
    hist = np.zeros((steps1, n), dtype=np.float32)
    for k in range(steps1):
        ids, counts = np.unique(states[:, k], return_counts=True)
        hist[k, ids] = counts / counts.sum()
    return hist
