import os, sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from cy2path_b200 import _lib, simulation
from cy2path_b200.sampling import get_plan
from cy2path_b200.synth import knn_velocity_tpm, chain_inputs
g = knn_velocity_tpm(50000, k=30, seed=2); T = g["T"]; plan = get_plan(T)
idx_d, cum_d = simulation._tables_device(plan)
C, steps = 100000, 1999
init, uni = chain_inputs(T, g["root_cells"], C, steps + 1, seed=1)
init_d = torch.from_numpy(init).cuda(); uni_d = torch.from_numpy(uni).cuda()
res = {}
for mode in ("1", "0", "1", "L4", "L16"):
    os.environ["CY2P_SAMPLER_PREFETCH"] = "0" if mode == "0" else "1"
    os.environ["CY2P_SAMPLER_PF_LANES"] = mode[1:] if mode.startswith("L") else "8"
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = simulation._walk(plan, idx_d, cum_d, init_d, uni_d)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res.setdefault(mode, []).append(dt)
    key = "states" if isinstance(out, dict) else 0
    st = out[key] if isinstance(out, dict) else out[0]
    res["sum" + mode] = int(st.long().sum().item())
print(json.dumps({"ms_4_lanes": [round(x * 1e3, 2) for x in res["L4"]], "ms_16_lanes": [round(x * 1e3, 2) for x in res["L16"]], "same4": res["sumL4"] == res["sum0"], "W": int(idx_d.shape[1]), "ms_prefetch": [round(x * 1e3, 2) for x in res["1"]], "ms_chunked": [round(x * 1e3, 2) for x in res["0"]],
                  "same_states": res["sum1"] == res["sum0"], "steps_per_s_prefetch": C * steps / min(res["1"])}))
