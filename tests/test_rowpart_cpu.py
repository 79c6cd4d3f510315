"""Host-side logic of the row-partitioned propagation (cy2path_b200/rowpart.py) — no GPU: the locality order from the
library's host entry point, balanced cuts, peer masks, per-rank matrix slices, and the exchange protocol simulated in
numpy with NaN-poisoned stale rows, against the float64 oracle (oracle/msm.py restates sampling.py:40-42)."""
import numpy as np
import pytest
import scipy.sparse as sp

from cy2path_b200 import rowpart
from cy2path_b200.synth import batched_starts, knn_velocity_tpm
from oracle import fast as ofast


@pytest.fixture(scope="module")
def graph():
    g = knn_velocity_tpm(4000, k=12, seed=31)
    T = g["T"]
    order = rowpart.host_order(T)
    return T, order, rowpart.permuted_transpose(T, order)


def test_host_order_is_a_permutation_and_local(graph):
    T, order, TtP = graph
    n = T.shape[0]
    assert sorted(order.tolist()) == list(range(n))
    # locality: mean |pos(dst) - pos(src)| far below the random order's n/3
    coo = TtP.tocoo()
    assert np.abs(coo.row - coo.col).mean() < 0.15 * n
    # the permuted transpose is T^T in the new index space
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    i, j = 123, T.indices[T.indptr[123]]
    assert TtP[pos[j], pos[i]] == T[i, j]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_balanced_bounds_masks_and_slices(graph, world):
    T, order, TtP = graph
    n = T.shape[0]
    eq = rowpart.equal_bounds(n, world)
    bal = rowpart.balanced_bounds(TtP, world)
    for b in (eq, bal):
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) > 0) and len(b) == world + 1
    cost = lambda b: (np.diff(b) + 0.6 * rowpart.halo_rows(TtP, b)).max()  # noqa: E731
    assert cost(bal) <= cost(eq) * 1.001
    masks = rowpart.peer_masks(TtP, bal)
    # brute force: rank g needs row q iff it owns q or one of its destination rows has source q
    dense = TtP.toarray() != 0
    for g in range(world):
        q0, q1 = bal[g], bal[g + 1]
        need = dense[q0:q1].any(0)
        need[q0:q1] = True
        assert np.array_equal(((masks >> np.uint32(g)) & 1).astype(bool), need)
    # slices: disjoint, complete, square, only the block's destination columns
    total = sp.csr_matrix((n, n), dtype=np.float32)
    for g in range(world):
        Tr = rowpart.sliced_matrix(TtP, bal[g], bal[g + 1])
        assert Tr.shape == (n, n) and Tr.indices.dtype == np.int32
        cols = np.unique(Tr.indices)
        assert cols.min() >= bal[g] and cols.max() < bal[g + 1]
        total = total + Tr
    assert abs(total - TtP.T).max() == 0
    assert sum(rowpart.sliced_matrix(TtP, bal[g], bal[g + 1]).nnz for g in range(world)) == T.nnz


@pytest.mark.parametrize("world,balance", [(1, False), (2, True), (5, True), (8, False)])
def test_halo_push_protocol_matches_oracle(graph, world, balance):
    """Only own rows + halo are ever written (everything else stays NaN and would poison any sum that touched it):
    after 6 updates the assembled iterate equals the oracle's."""
    T, order, TtP = graph
    n = T.shape[0]
    X0 = batched_starts(T, 3, seed=4).astype(np.float64)
    bounds = rowpart.balanced_bounds(TtP, world) if balance else rowpart.equal_bounds(n, world)
    masks = rowpart.peer_masks(TtP, bounds)
    got_p = rowpart.simulate_halo_push(TtP, bounds, masks, X0[order], 6)
    got = np.empty_like(got_p)
    got[order] = got_p
    ref = ofast.propagate_batch(T, X0, 6)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 1e-15


def test_label_propagation_shrinks_the_halo_and_keeps_the_protocol_exact(graph):
    T, order, TtP = graph
    n, world = T.shape[0], 4
    before = rowpart.halo_rows(TtP, rowpart.equal_bounds(n, world))
    new_order, bounds = rowpart.refine_partition(T, order, world)
    assert sorted(new_order.tolist()) == list(range(n)) and bounds[0] == 0 and bounds[-1] == n
    sizes = np.diff(bounds)
    assert sizes.min() >= 0.9 * n / world and sizes.max() <= 1.1 * n / world
    TtP2 = rowpart.permuted_transpose(T, new_order)
    after = rowpart.halo_rows(TtP2, bounds)
    assert after.sum() < 0.8 * before.sum() and after.max() <= before.max()
    X0 = batched_starts(T, 2, seed=9).astype(np.float64)
    got_p = rowpart.simulate_halo_push(TtP2, bounds, rowpart.peer_masks(TtP2, bounds), X0[new_order], 5)
    got = np.empty_like(got_p)
    got[new_order] = got_p
    assert np.abs(got - ofast.propagate_batch(T, X0, 5)).max() <= 1e-15


def test_row_partition_object_without_device(graph):
    T, order, TtP = graph
    rp = rowpart.RowPartition(T, world=4, rank=2, order=order, build_plan=False)
    s = rp.summary()
    assert sum(s["rows_per_rank"]) == T.shape[0] and s["nnz_total"] == T.nnz
    assert 0 < s["nnz_local"] < T.nnz / 2
    assert rp.rows_pushed > 0 and rp.q1 > rp.q0
