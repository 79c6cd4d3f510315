"""Timing probe (tuning aid): per-iteration cost of the single-start propagation kernels."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cy2path_b200 import _lib
from cy2path_b200.synth import knn_velocity_tpm

n = int(os.environ.get("PROBE_N", "3696"))
g = knn_velocity_tpm(n, k=30, seed=1)
plan = _lib.Plan(g["T"])
stat = g["end_points"] / g["end_points"].sum()
res = {"n": n, "cluster_env": os.environ.get("CY2P_CLUSTER_MAX_NNZ")}
dts = (torch.float64,) if os.environ.get("PROBE_ONE") else (torch.float32, torch.float64)
for dt in dts:
    x0 = torch.tensor(g["root_cells"] / g["root_cells"].sum(), dtype=dt, device="cuda")
    for eps in ((True,) if os.environ.get("PROBE_ONE") else (False, True)):
        ts = {}
        for iters in (200, 1000):
            fn = lambda: _lib.propagate(plan, x0, iters, stationary=stat if eps else None, want_final=False,
                                        history_stride=1, want_eps=eps)
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ts[iters] = e0.elapsed_time(e1) / 10 * 1e3
        res[f"{str(dt)[6:]}_eps{int(eps)}_us_per_iter"] = (ts[1000] - ts[200]) / 800
        res[f"{str(dt)[6:]}_eps{int(eps)}_us_1000"] = ts[1000]
print(json.dumps(res))
