"""Tuning sweep for the extended-precision batched propagation (cy2p_propagate_x) and the legacy float32 / float64
paths on a config-2 shaped problem. One process, one graph, many settings: the library reads CY2PX_CPL / CY2PX_WARPS /
CY2PX_ROWS / CY2P_TILE from the environment at every call. Not part of the product.

  python tools/k1x_sweep.py --n 50000 --B 4096 --iters 40 --out gpurun_out/k1x_sweep.jsonl [--configs a,b,...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# name -> (kind, tail_bits, env)
def _x(tail, **env):
    return ("x", tail, {"CY2PX_" + k.upper(): str(v) for k, v in env.items()})


CONFIGS = {
    "fp32": ("f32", 0, {}),
    "fp32_vec2": ("f32", 0, {"CY2P_GATHER_VEC": "2"}),  # 256-byte row segments per warp
    "fp32_vec1": ("f32", 0, {"CY2P_GATHER_VEC": "1"}),  # 128-byte row segments
    "fp64": ("f64", 0, {}),
}
for _tail, _nm in ((16, "f48"), (8, "f40")):
    for _cpl in (4, 8):
        for _w, _mb in ((8, 2), (8, 3), (16, 1)):
            for _u in (2, 4):
                CONFIGS["%s_c%d_w%d_m%d_u%d" % (_nm, _cpl, _w, _mb, _u)] = _x(_tail, cpl=_cpl, warps=_w, minb=_mb, u=_u)
for _cpl, _w, _mb, _r in ((8, 8, 2, 64), (8, 8, 2, 128), (8, 8, 2, 256), (8, 16, 1, 128), (8, 16, 1, 256), (8, 16, 1, 512),
                         (4, 8, 3, 128), (4, 16, 1, 256)):
    CONFIGS["f48_persist_c%d_w%d_m%d_r%d" % (_cpl, _w, _mb, _r)] = _x(16, cpl=_cpl, warps=_w, minb=_mb, u=4, persist=1, rows=_r)
for _t in (128, 256, 512, 2048):
    CONFIGS["f48_multi_c4_t%d" % _t] = ("x", 16, {"CY2PX_MULTI": "1", "CY2PX_CPL": "4", "CY2PX_MINB": "2", "CY2P_TILE": str(_t)})
    CONFIGS["f48_plain_c4_t%d" % _t] = ("x", 16, {"CY2PX_CPL": "4", "CY2PX_MINB": "2", "CY2P_TILE": str(_t)})
CONFIGS["f48_default"] = ("x", 16, {})
for _d in (1, 2, 4, 8, 12):  # eps-epilogue cost breakdown (results are NOT valid in these modes)
    CONFIGS["f48_epsdbg%d" % _d] = ("x", 16, {"CY2PX_EPS_DBG": str(_d)})
CONFIGS["f40_default"] = ("x", 8, {})
for _t in (128, 512, 1024):
    CONFIGS["f48_default_t%d" % _t] = ("x", 16, {"CY2P_TILE": str(_t)})
for _r in (32, 128, 256):
    CONFIGS["f48_default_r%d" % _r] = ("x", 16, {"CY2PX_ROWS": str(_r)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--eps", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch

    from cy2path_b200 import _lib
    from cy2path_b200.synth import batched_starts, knn_velocity_tpm, stationary_by_power

    g = knn_velocity_tpm(a.n, seed=2)
    T = g["T"]
    stat = stationary_by_power(T, 50) if a.eps else None
    X0 = torch.from_numpy(batched_starts(T, a.B, seed=3)).cuda()
    X0d = None
    plan = _lib.Plan(T)
    names = [c for c in a.configs.split(",") if c] or list(CONFIGS)
    out = open(a.out, "a") if a.out else None
    for name in names:
        kind, tail, env = CONFIGS[name]
        saved = {k: os.environ.get(k) for k in ("CY2PX_CPL", "CY2PX_WARPS", "CY2PX_ROWS", "CY2P_TILE", "CY2PX_MINB", "CY2PX_U", "CY2PX_PERSIST", "CY2PX_MULTI")}
        for k in saved:
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            if kind == "x":
                run = lambda: _lib.propagate_x(plan, X0, a.iters, stationary=stat, tail_bits=tail, want_eps=bool(a.eps))  # noqa: E731
            elif kind == "f32":
                run = lambda: _lib.propagate(plan, X0, a.iters, stationary=stat, want_eps=bool(a.eps))  # noqa: E731
            else:
                if X0d is None:
                    X0d = X0.double()
                run = lambda: _lib.propagate(plan, X0d, a.iters, stationary=stat, want_eps=bool(a.eps))  # noqa: E731
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            rec = {"config": name, "n": a.n, "B": a.B, "iters": a.iters, "eps": a.eps, "env": env, "ms": ms,
                   "ms_per_iter": ms / a.iters, "cell_it_per_s": a.n * a.B * a.iters / (ms * 1e-3),
                   "hbm_frac": (8 * T.nnz + 4 * (a.n + 1) + 8 * a.n * a.B) * a.iters / (ms * 1e-3) / 6553.9e9}
        except Exception as e:  # keep sweeping
            rec = {"config": name, "error": repr(e)[:300]}
        finally:
            for k, v in saved.items():
                os.environ.pop(k, None)
                if v is not None:
                    os.environ[k] = v
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()
        if kind == "f64":
            X0d = None
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
