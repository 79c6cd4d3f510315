"""Host-only probe (no GPU): how thin can the halo of the row-partitioned mode (config 5) get on the synthetic
kNN velocity graphs?  For a partition of the cells into `world` contiguous blocks of an order, the halo of a rank is
the set of source rows its destination rows gather that another rank owns (what `dist.row_peer_masks` pushes).

Compares, at 2 / 4 / 8 ranks: the caller's order, a Cuthill-McKee order (scipy's, a stand-in for the plan's own),
latent time, and the generator's ground truth (branch, latent time) — the best geometric cut there is.
Result at n = 100,000 (mean, max over ranks, fraction of n):
    identity     0.50 / 0.75 / 0.86
    rcm          0.205 / 0.238 (0.34) / 0.240 (0.36)
    tau          0.124 / 0.200 (0.25) / 0.226 (0.32)
    branch,tau   0.180 / 0.231 (0.27) / 0.174 (0.29)
=> the halo is a property of the graph class (k = 30 neighbours in 8 noisy dimensions: a cell's neighbourhood is
wide compared with a rank's slab), not of the ordering: a minimum-cut partitioner cannot make the exchange thin.

    python tools/halo_probe.py [n]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.sparse.csgraph import reverse_cuthill_mckee  # noqa: E402

from cy2path_b200.synth import knn_velocity_tpm  # noqa: E402


def halo(T, order, world):
    n = T.shape[0]
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    m = (n + world - 1) // world
    coo = T.tocoo()
    rd, rs = pos[coo.col] // m, pos[coo.row] // m  # owner of the destination row / of the source row
    fr = [len(np.unique(coo.row[(rd == r) & (rs != r)])) / n for r in range(world)]
    return float(np.mean(fr)), float(np.max(fr))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    seed = 5
    g = knn_velocity_tpm(n, k=30, seed=seed)
    T = g["T"]
    rng = np.random.default_rng(seed)  # replay the generator's first two draws: latent time, branch
    tau, branch = rng.random(n), rng.integers(0, 3, n)
    assert np.allclose(tau, g["tau"])
    orders = {
        "identity": np.arange(n),
        "rcm": reverse_cuthill_mckee((T + T.T).tocsr(), symmetric_mode=True),
        "tau": np.argsort(tau),
        "branch,tau": np.lexsort((tau, np.where(tau < 0.3, -1, branch))),
    }
    for name, o in orders.items():
        print(f"{name:12s}", "  ".join("%d ranks: mean %.3f max %.3f" % ((w,) + halo(T, o, w)) for w in (2, 4, 8)))


if __name__ == "__main__":
    main()
