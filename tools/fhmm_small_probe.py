"""Timing probe (tuning aid): cost per recurrence step of the one-CTA FHMM kernels, from the slope over I."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cy2path_b200 import _lib
from cy2path_b200.models import FHMM
from cy2path_b200.sampling import get_plan
from cy2path_b200.synth import knn_velocity_tpm

g = knn_velocity_tpm(1024, k=8, seed=1)
plan = get_plan(g["T"])
res = {}
for I in (64, 256, 1024, 2048):
    torch.manual_seed(0)
    m = FHMM(10, 4, 1024, I, restricted=True, use_gpu=True)
    P = [p.data for p in m.parameters()]
    D = torch.rand(I, 1024, device="cuda"); D /= D.sum(1, keepdim=True)
    for name, fn in (("forward", lambda: _lib.fhmm_forward(*P, I, True, want_pred=False)),
                     ("loss_grad", lambda: _lib.fhmm_loss_grad(plan, *P, D, True, (1.0, 0.1, 0.1, 0.0)))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res[f"{name}_I{I}_us"] = e0.elapsed_time(e1) / 20 * 1e3
print(json.dumps(res))
f = res
print("forward us/step:", (f["forward_I2048_us"] - f["forward_I256_us"]) / (2048 - 256))
print("loss_grad us/step:", (f["loss_grad_I2048_us"] - f["loss_grad_I256_us"]) / (2048 - 256))
