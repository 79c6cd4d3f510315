"""Timing aid: where the end-to-end (host buffers) step of config 2 spends its time beyond the kernels."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cy2path_b200 import _lib
from cy2path_b200.sampling import batch_convergence, clear_plan_cache, get_plan
from cy2path_b200.simulation import _tables_device, _walk

w = bench.WORKLOADS["config2"]
inp = bench.make_inputs(w, 0)
T = inp["T"]
for key in ("X0", "uni", "init_s", "stat"):
    inp[key] = torch.from_numpy(inp[key]).pin_memory().numpy()
res = {}
def tic():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    clear_plan_cache()
    t = tic(); plan = get_plan(T); res["plan_create"] = tic() - t
    t = tic(); x0 = _lib.to_device(inp["X0"], torch.float32, plan.device); res["h2d_X0"] = tic() - t
    t = tic(); out = _lib.propagate(plan, x0, w["iters"], stationary=inp["stat"], want_eps=True); res["propagate(incl. reorder on rep0)"] = tic() - t
    t = tic(); fin = _lib.to_host(out["final"]); eps = _lib.to_host(out["eps"]); res["d2h_final_eps"] = tic() - t
    t = tic(); batch_convergence(eps, 1e-5, w["iters"]); res["convergence_host"] = tic() - t
    t = tic(); idx_d, cum_d = _tables_device(plan); res["cdf_tables"] = tic() - t
    if rep == 0:
        inp["uni"] *= float(cum_d.max(1).values.min())
    t = tic(); st, pr, hi = _walk(plan, idx_d, cum_d, inp["init_s"], inp["uni"], want_history_steps=w["max_iter"]); res["chains_h2d+kernels"] = tic() - t
    t = tic(); a = _lib.to_host(st); b = _lib.to_host(pr); c = _lib.to_host(hi); res["d2h_chains"] = tic() - t
    print(json.dumps({k: round(v * 1e3, 1) for k, v in res.items()}))
