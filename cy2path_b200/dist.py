"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink; gloo for CPU tests).

Two partitionings:

* independent units — start distributions (columns of X) and chains are independent, so they are
  sharded across ranks with T replicated and NO collective on the data path (BASELINE configs 2/4);
  only the small per-start results (eps / convergence) are gathered at the end.
* row partition — for atlas-scale T (config 5) each rank owns a contiguous block of destination rows
  of T^T, computes its slice of the update from the whole iterate and the slices are all-gathered
  every iteration (``all_gather_into_tensor``, in place into the next-iterate buffer).

The per-rank arithmetic is always the CUDA library; ``update_rows`` is injectable only so the
world_size-2 gloo tests can exercise the sharding / exchange logic on a CPU box.
"""
from __future__ import annotations

import numpy as np


def shard_range(total, world, rank):
    """Contiguous [begin, end) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def row_blocks(n, world):
    """Equal-height row blocks (the all-gather needs equal slices): m = ceil(n / world) rows per rank,
    the last block is clipped to n. Returns (m, [(begin, end)] per rank)."""
    m = (int(n) + world - 1) // world
    return m, [(min(r * m, n), min((r + 1) * m, n)) for r in range(world)]


def _world(group=None):
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_ragged_int64(values, group=None):
    """All-gather a per-rank int64 vector of possibly different length; returns the concatenation."""
    import torch
    import torch.distributed as dist

    world, _ = _world(group)
    values = np.asarray(values, dtype=np.int64)
    if world == 1:
        return values
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(values)], dtype=torch.int64, device=dev), group=group)
    mx = int(max(c.item() for c in counts))
    buf = torch.full((max(mx, 1),), -1, dtype=torch.int64, device=dev)
    buf[: len(values)] = torch.as_tensor(values, device=dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return np.concatenate([o[: int(c.item())].cpu().numpy() for o, c in zip(outs, counts)])


def propagate_batch_sharded(T, X0, iters, stationary=None, tol=1e-5, group=None, precision="fp32", gather=True):
    """Batch-sharded MSM simulation: this rank propagates its own columns X0 [n, B_local].

    Returns the local result dict of ``sampling.propagate_batch`` plus, when ``gather`` and a
    stationary vector are given, ``convergence_all``: the per-start convergence cuts of every rank
    (the only exchange, after the loop)."""
    from .sampling import propagate_batch

    res = propagate_batch(T, X0, iters, stationary=stationary, tol=tol, precision=precision)
    if gather and "convergence" in res:
        res["convergence_all"] = gather_ragged_int64(res["convergence"], group)
    return res


def _cuda_update_rows(plan):
    from . import _lib

    lib = _lib.load()

    def update(X, Y, r0, r1):
        fn = lib.cy2p_propagate_rows_f64 if X.dtype.itemsize == 8 else lib.cy2p_propagate_rows
        _lib.check(fn(plan.handle, _lib.ptr(X), _lib.ptr(Y), int(X.shape[1]), int(r0), int(r1), _lib.stream_ptr(X.device)))

    return update


def propagate_row_partitioned(n, X0, iters, update_rows, group=None, history_every=0):
    """Row-partitioned propagation (BASELINE config 5).

    X0: [n, B] tensor on this rank's device (identical on every rank). ``update_rows(X, Y, r0, r1)``
    writes rows [r0, r1) of the update of X into Y (the CUDA path passes cy2p_propagate_rows through
    ``cuda_row_update(plan)``). Every iteration ends with an in-place all-gather of the row slices
    over NVLink. Returns the final iterate [n, B] (and a list of snapshots if history_every > 0)."""
    import torch
    import torch.distributed as dist

    world, rank = _world(group)
    m, blocks = row_blocks(n, world)
    B = X0.shape[1]
    n_pad = m * world
    cur = torch.zeros((n_pad, B), dtype=X0.dtype, device=X0.device)
    nxt = torch.zeros_like(cur)
    cur[:n].copy_(X0)
    r0, r1 = blocks[rank]
    snaps = []
    for it in range(int(iters)):
        if history_every and it % history_every == 0:
            snaps.append(cur[:n].clone())
        if r1 > r0:
            update_rows(cur, nxt, r0, r1)
        if world > 1:
            mine = nxt[rank * m:(rank + 1) * m]
            if dist.get_backend(group) != "nccl":
                mine = mine.clone()  # only NCCL documents the in-place form
            dist.all_gather_into_tensor(nxt, mine, group=group)  # NCCL: in place, slice `rank` of the output is the input
        cur, nxt = nxt, cur
    out = cur[:n]
    return (out, snaps) if history_every else out


def cuda_row_update(plan):
    """The product's row update: cy2p_propagate_rows on this rank's plan."""
    return _cuda_update_rows(plan)
