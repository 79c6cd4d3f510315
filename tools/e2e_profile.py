"""cProfile of the end-to-end (host buffers) step of a bench workload, results dropped between steps as in bench.py."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import bench
import cy2path_b200 as c2p
import cy2path_b200.simulation as sim
from cy2path_b200.sampling import clear_plan_cache

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
w = bench.WORKLOADS[name]
inp = bench.make_inputs(w, 0)
T, n = inp["T"], inp["T"].shape[0]
sim.estimate_stationary_state = lambda adata, matrix_key="T_forward": inp["stat"]
for key in ("X0", "uni", "init_s", "stat"):
    if inp.get(key) is not None:
        inp[key] = torch.from_numpy(inp[key]).pin_memory().numpy()

class Adata:
    pass

def step():
    clear_plan_cache()
    r = c2p.propagate_batch(T, inp["X0"], w["iters"], stationary=inp["stat"])
    ad = Adata()
    ad.obsp, ad.shape, ad.uns = {"T_forward": T}, (n, 0), {}
    ad.obs = pd.DataFrame({"root_cells": inp["root_cells"], "end_points": inp["end"]})
    c2p.sample_markov_chains(ad, num_chains=w["chains"], max_iter=w["max_iter"], convergence=w["max_iter"],
                             uniforms=inp["uni"], init_states=inp["init_s"])
    return r

# scale the uniforms as bench.py does (below the smallest final CDF entry)
from cy2path_b200.sampling import get_plan
from cy2path_b200.simulation import _tables_device
idx_d, cum_d = _tables_device(get_plan(T))
inp["uni"] *= float(cum_d.max(1).values.min())
step(); step()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
print("step wall ms", (time.perf_counter() - t0) * 1e3)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())
