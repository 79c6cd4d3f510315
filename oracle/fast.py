"""TEST INFRASTRUCTURE — ctypes wrappers over oracle/csrc/oracle.c (same arithmetic as
oracle/msm.py and oracle/chains.py, fast enough for mid-size parity cases)."""
import ctypes

import numpy as np

from .build import lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _csr(T):
    return (np.ascontiguousarray(T.indptr, dtype=np.int32), np.ascontiguousarray(T.indices, dtype=np.int32),
            np.ascontiguousarray(T.data, dtype=np.float32))


def spmv_left(x, T):
    ip, ix, dt = _csr(T)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().orc_spmv_left(ctypes.c_int(T.shape[0]), _p(ip, ctypes.c_int32), _p(ix, ctypes.c_int32),
                        _p(dt, ctypes.c_float), _p(x, ctypes.c_double), _p(y, ctypes.c_double))
    return y


def propagate_batch(T, X0, iters):
    """X0 [n, B] cell-major -> X after iters steps, float64 [n, B]."""
    ip, ix, dt = _csr(T)
    n, B = X0.shape
    X = np.ascontiguousarray(np.asarray(X0, dtype=np.float64).T)
    Y = np.empty_like(X)
    for _ in range(iters):
        lib().orc_spmm_left(ctypes.c_int(n), ctypes.c_int(B), _p(ip, ctypes.c_int32), _p(ix, ctypes.c_int32),
                            _p(dt, ctypes.c_float), _p(X, ctypes.c_double), _p(Y, ctypes.c_double))
        X, Y = Y, X
    return np.ascontiguousarray(X.T)


def sample_chains(T, idx, cum, init_states, uniforms):
    ip, ix, dt = _csr(T)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    cum = np.ascontiguousarray(cum, dtype=np.float32)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    C, steps = uniforms.shape
    W = idx.shape[1]
    states = np.empty((C, steps + 1), dtype=np.int32)
    probs = np.empty((C, steps), dtype=np.float32)
    f = lib().orc_chain
    for c in range(C):
        rc = f(ctypes.c_int(W), _p(idx, ctypes.c_int32), _p(cum, ctypes.c_float), _p(ip, ctypes.c_int32),
               _p(ix, ctypes.c_int32), _p(dt, ctypes.c_float), ctypes.c_int(int(init_states[c])),
               ctypes.c_int(steps), _p(uniforms[c], ctypes.c_double), _p(states[c], ctypes.c_int32),
               _p(probs[c], ctypes.c_float))
        if rc != 0:
            raise IndexError(f"CDF overshoot: chain {c} step {-rc}")
    return states, probs
