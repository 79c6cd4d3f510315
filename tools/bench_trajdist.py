"""Time K4 (DTW / Hausdorff distance matrices) at cytopath-like sizes and a bounded sample of the CPU restatement."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cy2path_b200 import _lib

for C, ln, d in ((200, 300, 30), (1000, 500, 30), (1000, 1000, 50)):
    rng = np.random.default_rng(0)
    s = torch.from_numpy(rng.standard_normal((C, ln, d)).cumsum(1)).cuda()
    out = {"C": C, "len": ln, "d": d, "pairs": C * (C - 1) // 2}
    for kind in ("dtw", "hausdorff"):
        _lib.trajectory_distance_matrix(s[:8], kind)
        torch.cuda.synchronize()
        t = time.perf_counter(); _lib.trajectory_distance_matrix(s, kind); torch.cuda.synchronize()
        dt = time.perf_counter() - t
        out[kind + "_s"] = round(dt, 4)
        out[kind + "_cell_dims_per_s"] = out["pairs"] * (1 if kind == "dtw" else 2) * ln * ln * d / dt
    if C == 200:
        from oracle import traj
        t = time.perf_counter(); traj.dtw_distance(s[0].cpu().numpy(), s[1].cpu().numpy()); one = time.perf_counter() - t
        out["oracle_python_one_pair_s"] = round(one, 3)
    print(json.dumps(out), flush=True)
