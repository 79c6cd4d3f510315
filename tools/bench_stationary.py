"""Time the stationary-state estimate: device subspace iteration (K1, B = 10, float64) vs host ARPACK."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cy2path_b200.synth import knn_velocity_tpm
from cy2path_b200.sampling import get_plan
from cy2path_b200.utils import eigs, eigs_device

for n, k in ((50_000, 30), (250_000, 30)):
    g = knn_velocity_tpm(n, k=k, seed=1)
    T = g["T"]
    get_plan(T)
    eigs_device(T)  # warm-up (plan, locality order)
    torch.cuda.synchronize()
    t = time.perf_counter(); vals, V = eigs_device(T); torch.cuda.synchronize(); td = time.perf_counter() - t
    t = time.perf_counter(); rv, rV = eigs(T); th = time.perf_counter() - t
    print(json.dumps({"n": n, "nnz": int(T.nnz), "device_s": round(td, 4), "arpack_s": round(th, 3),
                      "kept_device": len(vals), "kept_arpack": len(rv), "vals_device": np.round(vals, 8).tolist(),
                      "vals_arpack": np.round(rv, 8).tolist()}))
