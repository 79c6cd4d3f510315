"""Measured distances of the K3 outputs to the reference golden vectors (what the test tolerances are set from):
gradients as max|dg| / max|g|, decoder outputs as max-abs and relative to the magnitude of the output."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch

    from conftest import load_golden
    from cy2path_b200 import _lib
    from cy2path_b200.sampling import get_plan
    from test_fhmm_gpu import GOLDEN, NAMES, _cls, _model_from

    for name in GOLDEN:
        g = load_golden(name)
        m = _model_from(g, name)
        plan = get_plan(g["T"])
        D = torch.tensor(g["D"]).cuda()
        terms, grads, _ = _lib.fhmm_loss_grad(plan, *[p.data for p in m.parameters()], D, bool(g["restricted"]),
                                              tuple(g["weights"]), model=m._kind)
        rec = {"golden": name, "grad_rel": {}, "decoder_abs": {}, "decoder_mag": {}}
        for gq, nme in zip(grads, NAMES):
            rg = g["grad_" + nme]
            rec["grad_rel"][nme[13:]] = float(np.abs(gq.cpu().numpy() - rg).max() / max(np.abs(rg).max(), 1e-30))
        m2 = _cls(name)(int(g["S"]), int(g["L"]), int(g["N"]), int(g["I"]), restricted=bool(g["restricted"]), use_gpu=True)
        m2.load_state_dict({k: torch.tensor(g["after3_" + k]) for k in NAMES})
        Dh = torch.tensor(g["D"])
        la, lp = m2.filtering(Dh)
        lb, lg = m2.smoothing(Dh)
        ld, psi, lmax, best = m2.viterbi(Dh)
        for got, key in ((la, "log_alpha"), (lp, "log_probs"), (lb, "log_beta"), (lg, "log_gamma"), (ld, "log_delta"),
                         (lmax, "log_max")):
            ref = g[key]
            rec["decoder_abs"][key] = float(np.abs(got.cpu().numpy() - ref).max())
            rec["decoder_mag"][key] = float(np.abs(ref).max())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
