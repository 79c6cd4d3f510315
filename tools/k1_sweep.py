"""Tuning sweep for K1 (not part of the product): times cy2p_propagate on a config-2 shaped problem for
one setting of the CY2P_TILE / CY2P_ROWS_PER_CTA knobs (read from the environment by the library)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def blob_order(T, R=64):
    """Greedy BFS blobs of R cells over the symmetrised graph (locality experiment)."""
    from collections import deque

    import scipy.sparse as sp

    n = T.shape[0]
    A = (T + T.T).tocsr()
    visited = np.zeros(n, bool)
    order, frontier, nxt = [], deque(), 0
    while len(order) < n:
        seed = -1
        while frontier:
            v = frontier.popleft()
            if not visited[v]:
                seed = v
                break
        if seed < 0:
            while visited[nxt]:
                nxt += 1
            seed = nxt
        q, cnt = deque([seed]), 0
        while q and cnt < R:
            v = q.popleft()
            if visited[v]:
                continue
            visited[v] = True
            order.append(v)
            cnt += 1
            for u in A.indices[A.indptr[v]:A.indptr[v + 1]]:
                if not visited[u]:
                    q.append(u)
        frontier.extend(q)
    return np.array(order)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--eps", type=int, default=1)
    ap.add_argument("--f64", type=int, default=0)
    ap.add_argument("--order", default="none", choices=["none", "blob", "rcm", "tau"])
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import scipy.sparse as sp
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(a.n, seed=2)
    T = g["T"]
    if a.order != "none":
        if a.order == "blob":
            perm = blob_order(T)
        elif a.order == "tau":
            perm = np.argsort(g["tau"])
        else:
            from scipy.sparse.csgraph import reverse_cuthill_mckee

            perm = reverse_cuthill_mckee((T + T.T).astype(bool).tocsr(), symmetric_mode=True)
        T = sp.csr_matrix(T[perm][:, perm])
        T.sort_indices()
        T.indices = T.indices.astype(np.int32)
        T.indptr = T.indptr.astype(np.int32)
    stat = stationary_by_power(T, 50)
    X0 = torch.from_numpy(batched_starts(T, a.B, seed=3)).cuda()
    if a.f64:
        X0 = X0.double()
    plan = _lib.Plan(T)
    lib = _lib.load()
    tile = lib.cy2p_propagate_column_tile(plan.handle, a.B)
    st = stat if a.eps else None
    for _ in range(2):
        _lib.propagate(plan, X0, a.iters, stationary=st, want_eps=bool(a.eps))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        _lib.propagate(plan, X0, a.iters, stationary=st, want_eps=bool(a.eps))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rate = a.n * a.B * a.iters / (ms * 1e-3)
    print(json.dumps({"tag": a.tag, "n": a.n, "B": a.B, "iters": a.iters, "eps": a.eps, "f64": a.f64, "order": a.order,
                      "tile_env": os.environ.get("CY2P_TILE"), "rpc_env": os.environ.get("CY2P_ROWS_PER_CTA"),
                      "tile": tile, "ms": ms, "us_per_iter_per_128cols": ms * 1e3 / a.iters / (a.B / 128),
                      "cell_it_per_s": rate}))


if __name__ == "__main__":
    main()
