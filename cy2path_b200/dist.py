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


def propagate_batch_sharded(T, X0, iters, stationary=None, tol=1e-5, group=None, precision="f48", gather=True):
    """Batch-sharded MSM simulation: this rank propagates its own columns X0 [n, B_local].

    Returns the local result dict of ``sampling.propagate_batch`` plus, when ``gather`` and a
    stationary vector are given, ``convergence_all``: the per-start convergence cuts of every rank
    (the only exchange, after the loop)."""
    from .sampling import propagate_batch

    res = propagate_batch(T, X0, iters, stationary=stationary, tol=tol, precision=precision)
    if gather and "convergence" in res:
        res["convergence_all"] = gather_ragged_int64(res["convergence"], group)
    return res


def sample_chains_sharded(T, init_states, uniforms, group=None, walk=None):
    """Chain-sharded sampling: this rank walks ITS chains — ``init_states`` int32 [C_local],
    ``uniforms`` float64 [C_local, steps] — with the CDF tables replicated, no collective during the walks; the
    per-step visit counts are then summed over the ranks (one all-reduce of int32 [steps + 1, n]) and normalised by the
    total number of chains, so every rank holds the ``state_history_max_iter`` of the whole ensemble
    (reference simulation.py:156-162: ``counts / counts.sum()`` per step, stored float32).

    ``walk(init_states, uniforms) -> (states [C, steps+1], probs [C, steps])`` replaces the CUDA sampler in the
    world_size-2 gloo tests only. Returns a dict: ``state_indices`` / ``state_transition_probabilities`` of the local
    chains, ``state_history_max_iter`` float32 [steps + 1, n] (global), ``num_chains`` (global)."""
    import torch
    import torch.distributed as dist

    world, _ = _world(group)
    n = int(T.shape[0])
    if walk is None:
        from .sampling import get_plan
        from .simulation import _tables_device, _walk

        plan = get_plan(T)
        idx_d, cum_d = _tables_device(plan)
        steps = int(np.shape(uniforms)[1])
        states, probs, _, counts = _walk(plan, idx_d, cum_d, init_states, uniforms, want_history_steps=steps + 1,
                                         want_counts=True)
    else:
        states, probs = walk(init_states, uniforms)
        states, probs = torch.as_tensor(states), torch.as_tensor(probs)
        counts = torch.zeros((states.shape[1], n), dtype=torch.int32, device=states.device)
        counts.scatter_add_(1, states.t().long(), torch.ones_like(states.t(), dtype=torch.int32))
    total = torch.tensor([int(states.shape[0])], dtype=torch.int64, device=counts.device)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    num_chains = int(total.item())
    hist = (counts.double() / float(max(num_chains, 1))).float()  # the arithmetic of hist_normalise_kernel (sample.cu)
    return {"state_indices": states, "state_transition_probabilities": probs, "state_history_max_iter": hist,
            "num_chains": num_chains}


def _cuda_update_rows(plan):
    from . import _lib

    lib = _lib.load()

    def update(X, Y, r0, r1):
        fn = lib.cy2p_propagate_rows_f64 if X.dtype.itemsize == 8 else lib.cy2p_propagate_rows
        _lib.check(fn(plan.handle, _lib.ptr(X), _lib.ptr(Y), int(X.shape[1]), int(r0), int(r1), _lib.stream_ptr(X.device)))

    return update


def plan_order(plan):
    """Device int64 tensor: original cell id at each position of the plan's locality order (cy2p_plan_order)."""
    import torch

    from . import _lib

    order = torch.empty(plan.n, dtype=torch.int32, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.check(_lib.load().cy2p_plan_order(plan.handle, _lib.ptr(order), _lib.stream_ptr(plan.device)))
    return order.long()


def cuda_row_update_ordered(plan):
    """Row update with X and Y in the plan's locality order (cy2p_propagate_rows_ordered)."""
    from . import _lib

    lib = _lib.load()

    def update(X, Y, r0, r1):
        fn = lib.cy2p_propagate_rows_ordered_f64 if X.dtype.itemsize == 8 else lib.cy2p_propagate_rows_ordered
        _lib.check(fn(plan.handle, _lib.ptr(X), _lib.ptr(Y), int(X.shape[1]), int(r0), int(r1), _lib.stream_ptr(X.device)))

    return update


def propagate_row_partitioned_cuda(plan, X0, iters, group=None, ordered=True):
    """Config-5 driver on the CUDA path: with ``ordered`` the iterate is permuted into the plan's locality order
    once, every update runs in that order (gathers hit a narrow window of X), and the result is un-permuted once."""
    import torch

    if not ordered:
        return propagate_row_partitioned(plan.n, X0, iters, cuda_row_update(plan), group=group)
    order = plan_order(plan)
    Xp = X0.index_select(0, order)
    out_p = propagate_row_partitioned(plan.n, Xp, iters, cuda_row_update_ordered(plan), group=group)
    out = torch.empty_like(out_p)
    out.index_copy_(0, order, out_p)
    return out


def _symmetric_buffers(shape, dtype, dev, group=None):
    """Allocate a buffer of `shape` that every rank of `group` can address from its own kernels.

    Plumbing only: torch's symmetric-memory allocator (cuMem + handle exchange + peer mapping) provides the peer
    pointers and a device-side cross-rank barrier; the data path is cy2p_propagate_rows_allgather's stores.
    Returns (local tensor, [base pointer on rank r as seen from this rank], barrier callable)."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    world, rank = _world(group)
    if world == 1:
        t = torch.zeros(shape, dtype=dtype, device=dev)
        return t, [t.data_ptr()], (lambda: None)
    # allocation + rendezvous cost ~100 ms: keep the most recent buffer for repeated calls with the same shape
    key = (tuple(shape), dtype, str(dev), id(group))
    if _SYMM_CACHE.get("key") == key:
        return _SYMM_CACHE["value"]
    _SYMM_CACHE.clear()
    value = _symmetric_buffers_new(shape, dtype, dev, group)  # (stores its handle in the cache)
    _SYMM_CACHE.update(key=key, value=value)
    return value


_SYMM_CACHE = {}


def _symmetric_buffers_new(shape, dtype, dev, group=None):
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    t = symm.empty(shape, dtype=dtype, device=dev)
    hdl = symm.rendezvous(t, group=group if group is not None else dist.group.WORLD)
    t.zero_()
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    import os

    if os.environ.get("CY2P_FUSED_BARRIER", "nccl") == "symm":
        return t, ptrs, (lambda: hdl.barrier(channel=0))
    # stream-ordered cross-rank barrier: a one-element NCCL all-reduce (a rank leaves it only after every rank's
    # preceding kernel, i.e. its peer stores, has completed)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    _SYMM_CACHE["handle"] = hdl  # keeps the peer mappings alive as long as the cached buffer
    return t, ptrs, (lambda: dist.all_reduce(flag, group=group))


def propagate_row_partitioned_fused(plan, X0, iters, group=None):
    """Config-5 driver with the exchange fused into the update: every rank's kernel stores its finished rows
    straight into ALL ranks' next-iterate buffers (own HBM + NVLink peer stores, cy2p_propagate_rows_allgather),
    so no separate all-gather follows the gather; ranks only meet at a device-side barrier that orders the
    ping-pong buffers. The iterate stays in the plan's locality order throughout."""
    import ctypes

    import torch

    from . import _lib

    lib = _lib.load()
    world, rank = _world(group)
    n, B = plan.n, int(X0.shape[1])
    dev = X0.device
    order = plan_order(plan)
    m, blocks = row_blocks(n, world)
    with torch.cuda.device(dev):
        buf, ptrs, barrier = _symmetric_buffers((2, n, B), X0.dtype, dev, group)
        half = n * B * X0.element_size()
        ptr_arrays = [torch.tensor([p + k * half for p in ptrs], dtype=torch.int64, device=dev) for k in range(2)]
        buf[0].copy_(X0.index_select(0, order))
        fn = lib.cy2p_propagate_rows_allgather_f64 if X0.dtype.itemsize == 8 else lib.cy2p_propagate_rows_allgather
        r0, r1 = blocks[rank]
        barrier()  # every rank has mapped its peers and filled buffer 0
        for it in range(int(iters)):
            cur, nxt = it & 1, (it + 1) & 1
            _lib.check(fn(plan.handle, _lib.ptr(buf[cur]), ctypes.c_void_p(ptr_arrays[nxt].data_ptr()), world, B,
                          int(r0), int(r1), _lib.stream_ptr(dev)))
            barrier()  # all ranks' stores into buffer `nxt` have landed; buffer `cur` may be overwritten next
        out_p = buf[int(iters) & 1]
        out = torch.empty_like(out_p)
        out.index_copy_(0, order, out_p)
        barrier()  # nobody releases its buffers while a peer may still be writing
    return out


def row_peer_masks(T, order, world):
    """For the row partition in the plan's order: uint32 mask per permuted row q, bit g set when rank g needs row q
    of the iterate — it owns q, or one of its destination rows has q as a source. ``T``: scipy CSR (T[i, j] moves
    mass from cell i to cell j), ``order``: original cell id at each permuted position. Host-side, once per plan."""
    import numpy as np

    n = T.shape[0]
    order = np.asarray(order, dtype=np.int64)
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    m, _ = row_blocks(n, world)
    coo = T.tocoo()
    src_pos, dst_owner = pos[coo.row], pos[coo.col] // m
    masks = (np.uint32(1) << (np.arange(n) // m).astype(np.uint32)).astype(np.uint32)
    for g in range(world):
        need = np.unique(src_pos[dst_owner == g])
        masks[need] |= np.uint32(1 << g)
    return masks


def propagate_row_partitioned_halo(plan, T, X0, iters, group=None):
    """Config-5 driver with a HALO push instead of an all-gather: the update kernel stores a finished row only into
    the buffers of the ranks that gather it (cy2p_propagate_rows_push with ``row_peer_masks``), over the same
    symmetric-memory mappings as the fused all-gather. Every rank keeps a full-size iterate of which only its own
    rows and its halo are current; the final iterate is assembled with one all-gather at the end."""
    import ctypes

    import torch
    import torch.distributed as dist

    from . import _lib

    lib = _lib.load()
    world, rank = _world(group)
    if world > 32:
        raise ValueError("the halo push supports at most 32 ranks")
    n, B = plan.n, int(X0.shape[1])
    dev = X0.device
    order = plan_order(plan)
    m, blocks = row_blocks(n, world)
    # host pass over the stored entries, once per (plan, world); kept ON the plan object so the masks live and die
    # with it (a cache keyed by the C handle's address could be hit by a later plan allocated at the same address)
    cache = plan.__dict__.setdefault("_halo_masks", {})
    if world not in cache:
        cache[world] = torch.from_numpy(row_peer_masks(T, order.cpu().numpy(), world).view("int32")).to(dev)
    masks = cache[world]
    with torch.cuda.device(dev):
        buf, ptrs, barrier = _symmetric_buffers((2, n, B), X0.dtype, dev, group)
        half = n * B * X0.element_size()
        ptr_arrays = [torch.tensor([p + k * half for p in ptrs], dtype=torch.int64, device=dev) for k in range(2)]
        x0p = X0.index_select(0, order)
        buf[0].copy_(x0p)
        fn = lib.cy2p_propagate_rows_push_f64 if X0.dtype.itemsize == 8 else lib.cy2p_propagate_rows_push
        r0, r1 = blocks[rank]
        barrier()
        for it in range(int(iters)):
            cur, nxt = it & 1, (it + 1) & 1
            _lib.check(fn(plan.handle, _lib.ptr(buf[cur]), ctypes.c_void_p(ptr_arrays[nxt].data_ptr()), world,
                          _lib.ptr(masks), B, int(r0), int(r1), _lib.stream_ptr(dev)))
            barrier()
        out_p = buf[int(iters) & 1]
        if world > 1:  # only own rows and halo are current: assemble the whole iterate once
            full = torch.zeros((m * world, B), dtype=X0.dtype, device=dev)
            full[rank * m:rank * m + (r1 - r0)].copy_(out_p[r0:r1])
            dist.all_gather_into_tensor(full, full[rank * m:(rank + 1) * m], group=group)
            out_p = full[:n]
        out = torch.empty((n, B), dtype=X0.dtype, device=dev)
        out.index_copy_(0, order, out_p)
        barrier()
    return out


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
