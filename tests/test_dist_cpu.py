"""CPU tests of the multi-GPU layer's host logic with world_size-2 gloo process groups.

The per-rank arithmetic of the product is CUDA; here the row update is supplied by the oracle so that
the sharding and the per-iteration all-gather exchange can be checked against the unsharded result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cy2path_b200 import dist as c2dist


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 4096, 100_001):
        for world in (1, 2, 3, 8):
            spans = [c2dist.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_row_blocks():
    m, blocks = c2dist.row_blocks(10, 4)
    assert m == 3 and blocks == [(0, 3), (3, 6), (6, 9), (9, 10)]
    m, blocks = c2dist.row_blocks(5, 8)
    assert m == 1 and blocks[5:] == [(5, 5)] * 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp

        from cy2path_b200.synth import batched_starts, knn_velocity_tpm

        g = knn_velocity_tpm(301, k=6, seed=17)
        T = g["T"]
        Tt = sp.csr_matrix(T.T.astype(np.float64))
        X0 = torch.from_numpy(batched_starts(T, 5, seed=1).astype(np.float64))

        def update(X, Y, r0, r1):  # oracle arithmetic for rows [r0, r1): y[j,:] = sum_i T[i,j] x[i,:]
            Y[r0:r1] = torch.from_numpy(Tt[r0:r1] @ X[:T.shape[0]].numpy())

        out = c2dist.propagate_row_partitioned(T.shape[0], X0, 9, update)
        ref = X0.numpy().copy()
        for _ in range(9):
            ref = Tt @ ref
        assert np.abs(out.numpy() - ref).max() <= 1e-14
        # ragged gather of per-start results (batch sharding's only exchange)
        b, e = c2dist.shard_range(11, world, rank)
        allv = c2dist.gather_ragged_int64(np.arange(b, e), None)
        assert np.array_equal(allv, np.arange(11))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _chain_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cy2path_b200.synth import chain_inputs, knn_velocity_tpm
        from oracle import chains as ochains
        from oracle import fast as ofast

        g = knn_velocity_tpm(257, k=7, seed=23)
        T = g["T"]
        idx, cum = ochains.extract_nonzero_entries(T)
        C, max_iter = 37, 21  # 19 + 18 chains: uneven shards
        init_s, uni = chain_inputs(T, g["root_cells"], C, max_iter, seed=3)
        uni *= float(cum.max(1).min())

        def walk(i, u):  # the oracle's walk stands in for the CUDA sampler
            return ofast.sample_chains(T, idx, cum, i, u)

        b, e = c2dist.shard_range(C, world, rank)
        out = c2dist.sample_chains_sharded(T, init_s[b:e], uni[b:e], walk=walk)
        states, probs = walk(init_s, uni)  # the whole ensemble, unsharded
        assert out["num_chains"] == C
        assert np.array_equal(out["state_indices"].numpy(), states[b:e])
        assert np.array_equal(out["state_transition_probabilities"].numpy().view(np.uint32), probs[b:e].view(np.uint32))
        ref = np.zeros((max_iter, T.shape[0]), dtype=np.float32)
        for k in range(max_iter):  # reference simulation.py:156-162
            ind, cnt = np.unique(states[:, k], return_counts=True)
            ref[k, ind] = cnt / cnt.sum()
        hist = out["state_history_max_iter"].numpy()
        assert hist.dtype == np.float32 and np.array_equal(hist.view(np.uint32), ref.view(np.uint32))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_chain_sharding_histogram_world2_gloo(tmp_path):
    """Chains sharded over two ranks: local paths equal the unsharded ones, the all-reduced per-step histogram is
    bit-equal to the reference's ``counts / counts.sum()`` over the whole ensemble on both ranks."""
    port = _free_port()
    mp.spawn(_chain_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


@pytest.mark.timeout(180)
def test_row_partition_and_gather_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_row_peer_masks_are_sufficient_for_a_halo_exchange():
    """Emulates the halo push of the row-partitioned mode on the host: every rank keeps a full-size iterate, computes
    its own rows, and pushes a row only to the ranks named in its mask. The assembled result must equal the unsharded
    propagation, and the masks must name far fewer rows than an all-gather moves."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    from cy2path_b200.dist import row_blocks, row_peer_masks
    from cy2path_b200.synth import knn_velocity_tpm

    n, B, iters = 3000, 3, 7
    g = knn_velocity_tpm(n, k=12, seed=5)
    T = g["T"].astype(np.float64)
    order = np.asarray(reverse_cuthill_mckee(sp.csr_matrix(((T + T.T) > 0).astype(np.int8)), symmetric_mode=True))
    rng = np.random.default_rng(0)
    X0 = rng.random((n, B))
    ref = X0.copy()
    for _ in range(iters):
        ref = T.T @ ref
    Tp = sp.csr_matrix(T[order][:, order])  # permuted operator
    for world in (1, 2, 4, 5):
        masks = row_peer_masks(g["T"], order, world)
        assert masks.dtype == np.uint32 and masks.shape == (n,)
        m, blocks = row_blocks(n, world)
        bufs = [X0[order].copy() for _ in range(world)]
        for _ in range(iters):
            new = [np.full((n, B), np.nan) for _ in range(world)]  # stale rows poison the result if ever read
            for r, (r0, r1) in enumerate(blocks):
                rows = (Tp.T.tocsr()[r0:r1] @ np.nan_to_num(bufs[r], nan=np.inf))
                assert np.isfinite(rows).all()  # every source the rank reads was pushed to it
                for peer in range(world):
                    sel = np.nonzero((masks[r0:r1] >> np.uint32(peer)) & np.uint32(1))[0] + r0
                    new[peer][sel] = rows[sel - r0]
            bufs = new
        out_p = np.concatenate([bufs[r][r0:r1] for r, (r0, r1) in enumerate(blocks)])
        out = np.empty_like(out_p)
        out[order] = out_p
        assert np.allclose(out, ref, rtol=1e-12, atol=1e-15), world
        if world >= 2:
            pushed = sum(int(((masks >> np.uint32(p)) & np.uint32(1)).sum()) for p in range(world)) - n
            assert pushed < 0.7 * (world - 1) * n  # the halo is well below the all-gather's (world-1)*n rows


def _rowpart_worker(rank, world, port, out_dir):
    """Two processes, each with its OWN RowPartition (order, refined blocks, masks, matrix slice built independently
    from the same T): per update a rank forms its rows from its slice of T and a full-size iterate in which only own
    rows + halo were ever written (the rest is NaN), and sends each peer exactly the rows whose mask names that peer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cy2path_b200 import rowpart
        from cy2path_b200.synth import batched_starts, knn_velocity_tpm
        from oracle import fast as ofast

        g = knn_velocity_tpm(2500, k=10, seed=19)
        T = g["T"]
        n = T.shape[0]
        rp = rowpart.RowPartition(T, world, rank, build_plan=False)
        TtP = rowpart.permuted_transpose(T, rp.order)
        mine = rowpart.sliced_matrix(TtP, rp.q0, rp.q1)  # the rank's share of T: only destinations in [q0, q1)
        assert mine.nnz == rp.nnz_local < T.nnz
        X0 = batched_starts(T, 3, seed=2).astype(np.float64)
        peer = 1 - rank
        bit_me, bit_peer = np.uint32(1 << rank), np.uint32(1 << peer)
        buf = np.full((n, 3), np.nan)
        need = (rp.masks & bit_me) != 0
        buf[need] = X0[rp.order][need]
        send_rows = np.nonzero((rp.masks[rp.q0:rp.q1] & bit_peer) != 0)[0] + rp.q0
        cnt = torch.tensor([len(send_rows)])
        cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        for _ in range(8):
            rows = (mine.T.tocsr()[rp.q0:rp.q1].astype(np.float64) @ np.nan_to_num(buf, nan=np.inf))
            nxt = np.full((n, 3), np.nan)
            nxt[rp.q0:rp.q1] = rows
            out_t = torch.from_numpy(np.ascontiguousarray(rows[send_rows - rp.q0]))
            idx_t = torch.from_numpy(send_rows.astype(np.int64))
            in_t = torch.empty((int(cnts[peer]), 3), dtype=torch.float64)
            in_idx = torch.empty(int(cnts[peer]), dtype=torch.int64)
            ops = [dist.isend(out_t, peer), dist.isend(idx_t, peer)] if rank == 0 else []
            dist.recv(in_t, peer)
            dist.recv(in_idx, peer)
            if rank == 1:
                ops = [dist.isend(out_t, peer), dist.isend(idx_t, peer)]
            for o in ops:
                o.wait()
            nxt[in_idx.numpy()] = in_t.numpy()
            buf = nxt
        assert np.isfinite(buf[rp.q0:rp.q1]).all()
        ref = ofast.propagate_batch(T, X0, 8)[rp.order]
        assert np.abs(buf[rp.q0:rp.q1] - ref[rp.q0:rp.q1]).max() <= 1e-15
        assert np.isnan(buf).any()  # rows this rank never gathers were never sent to it
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_rowpart_halo_exchange_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_rowpart_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
