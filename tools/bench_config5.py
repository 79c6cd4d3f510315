"""BASELINE config 5: 1M-cell TPM, destination rows of T^T partitioned across ranks, per-iteration all-gather.

  torchrun --nproc-per-node G tools/bench_config5.py [--n 1000000] [--B 1|256] [--iters 200]
Prints one JSON line (rank 0): cell-iterations/s, us per iteration, all-gather bytes and achieved GB/s.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", dest="n", type=int, default=1_000_000)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--f64", type=int, default=0)
    ap.add_argument("--ordered", type=int, default=1)
    ap.add_argument("--fused", type=int, default=0, help="fused update + peer-store all-gather instead of NCCL")
    ap.add_argument("--halo", type=int, default=0, help="halo push (per-row peer masks) instead of the full exchange")
    ap.add_argument("--cache", default="/tmp/cy2p_config5_graph.npz", help="graph cache on the box (generated once)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cy2path_b200 import _lib, dist as c2dist
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm

    import scipy.sparse as sp

    cache = f"{a.cache}.{a.n}"
    if rank == 0 and not os.path.exists(cache):
        g = knn_velocity_tpm(a.n, seed=5)
        np.savez(cache + ".tmp.npz", indptr=g["T"].indptr, indices=g["T"].indices, data=g["T"].data)
        os.replace(cache + ".tmp.npz", cache)
    if world > 1:
        dist.barrier()
    z = np.load(cache)
    T = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(a.n, a.n))
    plan = _lib.Plan(T)
    X0 = torch.from_numpy(batched_starts(T, a.B, seed=5)).cuda()
    if a.f64:
        X0 = X0.double()
    def run(iters):
        if a.halo:
            return c2dist.propagate_row_partitioned_halo(plan, T, X0, iters)
        if a.fused:
            return c2dist.propagate_row_partitioned_fused(plan, X0, iters)
        return c2dist.propagate_row_partitioned_cuda(plan, X0, iters, ordered=bool(a.ordered))

    for _ in range(2):
        run(10)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run(a.iters)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # parity across partitionings: every rank holds the same full iterate; compare a checksum with the unsharded run
    chk = float(out.double().sum().item())
    ref = _lib.propagate(plan, X0, a.iters)["final"]
    err = float((out - ref).abs().max().item())
    if rank == 0:
        ms = float(ms.item())
        elem = 8 if a.f64 else 4
        ag_bytes = (world - 1) / world * elem * a.n * a.B  # received per rank per iteration
        print(json.dumps({"workload": "config5", "n": a.n, "nnz": int(T.nnz), "B": a.B, "iters": a.iters, "n_gpus": world,
                          "ms": ms, "us_per_iter": ms * 1e3 / a.iters,
                          "cell_iterations_per_s": a.n * a.B * a.iters / (ms * 1e-3),
                          "allgather_bytes_per_iter_per_rank": ag_bytes,
                          "allgather_GBps_if_all_time_were_comm": ag_bytes * a.iters / (ms * 1e-3) / 1e9,
                          "ordered": a.ordered, "fused": a.fused, "halo": a.halo, "mass_checksum": chk, "max_abs_vs_unsharded": err, "dtype": "f64" if a.f64 else "f32"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
