"""cProfile of the end-to-end config-1 step (single start, full float64 history) as bench.py runs it."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import cy2path_b200 as c2p
from cy2path_b200.sampling import clear_plan_cache

w = bench.WORKLOADS["config1"]
inp = bench.make_inputs(w, 0)
T, n = inp["T"], inp["T"].shape[0]

class Adata:
    pass

def step():
    clear_plan_cache()
    ad0 = Adata()
    ad0.obsp, ad0.shape, ad0.uns = {"T_forward": T}, (n, 0), {}
    return c2p.iterate_state_probability(ad0, init=inp["X0"][:, 0].astype(np.float64), stationary=inp["stat"],
                                         max_iter=w["iters"])

for _ in range(5):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("step ms: min %.2f median %.2f max %.2f" % (min(ts), sorted(ts)[10], max(ts)))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue())
