"""Multi-GPU parity check of the row-partitioned propagation (run under torchrun, one rank per GPU):
the assembled result of cy2path_b200.rowpart.propagate on N GPUs against the unsharded float64 propagation of the same
start distributions on each rank's own GPU, and against scipy's x @ T (the reference's arithmetic) for one column.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 \
      tools/rowpart_check.py [--n 200000]
Prints one JSON line per case on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--iters", type=int, default=31)
    ap.add_argument("--bench-n", type=int, default=0, help="also time B = 1 float64 at this size, persistent kernel on / off")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bench import shared_graph
    from cy2path_b200 import _lib, rowpart
    from cy2path_b200.synth import batched_starts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    T, _ = shared_graph(a.n, 7, rank, barrier)
    rp = rowpart.RowPartition(T, world, rank, device=local, sweeps=8)
    full = _lib.Plan(T)  # the unsharded comparison only; the partitioned run never touches it
    for B, dt in ((1, torch.float64), (8, torch.float64), (64, torch.float32)):
        X0 = torch.from_numpy(batched_starts(T, B, seed=11)).cuda().to(dt)
        got = rowpart.propagate(rp, X0, a.iters)
        ref = _lib.propagate(full, X0 if B > 1 else X0.reshape(-1), a.iters)["final"].reshape(a.n, B)
        err = float((got - ref).abs().max())
        x = X0[:, 0].double().cpu().numpy()
        for _ in range(a.iters):
            x = x @ T
        err_scipy = float(np.abs(got[:, 0].double().cpu().numpy() - x).max())
        rel_scipy = float(np.abs(got[:, 0].double().cpu().numpy() - x).sum() / np.abs(x).sum())
        errs = torch.tensor([err, err_scipy, rel_scipy], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"world": world, "n": a.n, "B": B, "dtype": str(dt), "iters": a.iters,
                              "max_abs_vs_unsharded": float(errs[0]), "max_abs_vs_scipy_col0": float(errs[1]),
                              "rel_l1_vs_scipy_col0": float(errs[2]), "partition": rp.summary()["halo_rows_per_rank"],
                              "rows": rp.summary()["rows_per_rank"]}), flush=True)
    if a.bench_n:
        Tb, _ = shared_graph(a.bench_n, 5, rank, barrier)
        rpb = rowpart.RowPartition(Tb, world, rank, device=local, sweeps=8)
        x0 = torch.from_numpy(batched_starts(Tb, 1, seed=5)).cuda().double()
        for persist in ("1", "0"):
            os.environ["CY2P_ROWPART_PERSIST"] = persist
            rowpart.propagate(rpb, x0, 10, assemble=False)
            barrier()
            res = []
            for iters in (200, 0):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rowpart.propagate(rpb, x0, iters, assemble=False)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res.append(float(t))
            if rank == 0:
                print(json.dumps({"bench": "B=1 f64", "n": a.bench_n, "world": world, "persistent_kernel": persist == "1",
                                  "us_per_update": (res[0] - res[1]) * 1e3 / 200}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
