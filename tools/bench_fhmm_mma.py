"""Config-3 shapes: ms per cy2p_fhmm_loss_grad call for the fused FFMA2 data pass, the forward + FFMA slices split and
the forward + tensor-core gradients split (mma.sync 3xTF32), restricted (K = 10 emission rows) and unrestricted (K = 40).
One process; the library reads CY2P_FHMM_MMA / CY2P_FHMM_TWO_PASS at every call.   --only NAME runs one case (for ncu)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    "restricted_mma_overlap": (True, {"CY2P_FHMM_TWO_PASS": "1", "CY2P_FHMM_MMA": "1", "CY2P_FHMM_OVERLAP": "1"}),
    "restricted_mma_serial": (True, {"CY2P_FHMM_TWO_PASS": "1", "CY2P_FHMM_MMA": "1", "CY2P_FHMM_OVERLAP": "0"}),
    "unrestricted_mma_overlap": (False, {"CY2P_FHMM_TWO_PASS": "1", "CY2P_FHMM_MMA": "1", "CY2P_FHMM_OVERLAP": "1"}),
    "unrestricted_mma_serial": (False, {"CY2P_FHMM_TWO_PASS": "1", "CY2P_FHMM_MMA": "1", "CY2P_FHMM_OVERLAP": "0"}),
    "restricted_fused_ffma2": (True, {"CY2P_FHMM_TWO_PASS": "0", "CY2P_FHMM_MMA": "0", "CY2P_FHMM_OVERLAP": "0"}),
    "restricted_twopass_ffma": (True, {"CY2P_FHMM_TWO_PASS": "1", "CY2P_FHMM_MMA": "0"}),
    "restricted_twopass_mma": (True, {"CY2P_FHMM_TWO_PASS": "1", "CY2P_FHMM_MMA": "1"}),
    "unrestricted_twopass_ffma": (False, {"CY2P_FHMM_TWO_PASS": "0", "CY2P_FHMM_MMA": "0"}),
    "unrestricted_twopass_mma": (False, {"CY2P_FHMM_TWO_PASS": "0", "CY2P_FHMM_MMA": "1"}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=50000)
    ap.add_argument("--I", type=int, default=500)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    import torch

    from cy2path_b200 import _lib, models
    from cy2path_b200.sampling import get_plan
    from cy2path_b200.synth import knn_velocity_tpm

    S, L = 10, 4
    g = knn_velocity_tpm(a.N, seed=3)
    plan = get_plan(g["T"])
    x0 = torch.full((a.N,), 1.0 / a.N, dtype=torch.float64, device="cuda")
    D = _lib.propagate(plan, x0, a.I, history_stride=1, want_final=False)["history"].float().contiguous()
    w = (1.0, 0.1, 0.1, 0.05)
    for name, (restricted, env) in CASES.items():
        if a.only and name != a.only:
            continue
        os.environ.update(env)
        torch.manual_seed(3)
        m = models.FHMM(S, L, a.N, a.I, restricted=restricted, use_gpu=True)
        params = [p.data for p in m.parameters()]
        terms, grads, bufs = _lib.fhmm_loss_grad(plan, *params, D, restricted, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            _lib.fhmm_loss_grad(plan, *params, D, restricted, w, out=bufs)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"case": name, "N": a.N, "I": a.I, "S": S, "L": L, "K": S * (1 if restricted else L),
                          "loss_grad_ms": e0.elapsed_time(e1) / a.reps, "loss": float(terms[0]),
                          "g_E_absmax": float(grads[2].abs().max())}), flush=True)


if __name__ == "__main__":
    main()
