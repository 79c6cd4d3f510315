"""ctypes binding of the C ABI in include/cy2path_b200.h.

PyTorch is used only for device memory, streams and host<->device copies; every arithmetic
step of the hot path runs in libcy2path_b200.so. There is NO CPU fallback: if the library is
missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcy2path_b200.so")

OK, ERR_INVALID_ARG, ERR_BAD_CSR, ERR_CDF_OVERSHOOT, ERR_CUDA, ERR_WORKSPACE, ERR_CDF_TABLE = range(7)

# every symbol the header declares (tests check the built library exports all of them)
SYMBOLS = [
    "cy2p_abi_version", "cy2p_last_error", "cy2p_launch_count", "cy2p_plan_create", "cy2p_plan_create_host",
    "cy2p_plan_create_host_f64", "cy2p_plan_destroy",
    "cy2p_plan_info", "cy2p_propagate_ws_bytes", "cy2p_propagate_column_tile", "cy2p_propagate", "cy2p_propagate_rows",
    "cy2p_propagate_f64_ws_bytes", "cy2p_propagate_f64", "cy2p_propagate_rows_f64", "cy2p_plan_order",
    "cy2p_propagate_x_ws_bytes", "cy2p_propagate_x_column_tile", "cy2p_propagate_x", "cy2p_order_host", "cy2p_xbarrier", "cy2p_propagate_rows_persist", "cy2p_propagate_rows_persist_f64",
    "cy2p_propagate_rows_ordered", "cy2p_propagate_rows_ordered_f64", "cy2p_propagate_rows_allgather",
    "cy2p_propagate_rows_allgather_f64", "cy2p_propagate_rows_push", "cy2p_propagate_rows_push_f64",
    "cy2p_enable_peer_access", "cy2p_cdf_build",
    "cy2p_sample_chains", "cy2p_chain_history", "cy2p_fhmm_ws_bytes", "cy2p_fhmm_small_floats",
    "cy2p_fhmm_forward", "cy2p_fhmm_loss_grad", "cy2p_fhmm_decode", "cy2p_lssm_ws_bytes", "cy2p_lssm_small_floats",
    "cy2p_lssm_forward", "cy2p_lssm_loss_grad", "cy2p_lssm_decode", "cy2p_nssm_ws_bytes", "cy2p_nssm_forward",
    "cy2p_nssm_loss_grad", "cy2p_nssm_decode", "cy2p_traj_ws_bytes", "cy2p_dtw_matrix", "cy2p_hausdorff_matrix",
]

_lib = None
_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t


class Cy2pError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises ImportError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. The CUDA library is required (there is no CPU fallback); build it with "
            "`python -m cy2path_b200._build` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.cy2p_abi_version.restype = _i32
    lib.cy2p_last_error.restype = ctypes.c_char_p
    lib.cy2p_launch_count.restype = _i64
    lib.cy2p_plan_create.argtypes = [_i32, _i64, _vp, _vp, _vp, _i32, _vp, ctypes.POINTER(_vp)]
    lib.cy2p_plan_create_host.argtypes = lib.cy2p_plan_create.argtypes
    lib.cy2p_plan_create_host_f64.argtypes = lib.cy2p_plan_create.argtypes
    lib.cy2p_plan_destroy.argtypes = [_vp]
    lib.cy2p_plan_info.argtypes = [_vp, ctypes.POINTER(_i64)]
    lib.cy2p_propagate_ws_bytes.argtypes = [_vp, _i32, _i32]
    lib.cy2p_propagate_ws_bytes.restype = _sz
    lib.cy2p_propagate_column_tile.argtypes = [_vp, _i32]
    lib.cy2p_propagate.argtypes = [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]
    lib.cy2p_propagate_rows.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _vp]
    lib.cy2p_propagate_f64_ws_bytes.argtypes = [_vp, _i32, _i32]
    lib.cy2p_propagate_f64_ws_bytes.restype = _sz
    lib.cy2p_propagate_f64.argtypes = lib.cy2p_propagate.argtypes
    lib.cy2p_propagate_rows_f64.argtypes = lib.cy2p_propagate_rows.argtypes
    lib.cy2p_plan_order.argtypes = [_vp, _vp, _vp]
    lib.cy2p_propagate_x_ws_bytes.argtypes = [_vp, _i32, _i32, _i32, _i32]
    lib.cy2p_propagate_x_ws_bytes.restype = _sz
    lib.cy2p_propagate_x_column_tile.argtypes = [_vp, _i32, _i32]
    lib.cy2p_order_host.argtypes = [_i32, _i64, _vp, _vp, _i32, _vp]
    lib.cy2p_xbarrier.argtypes = [_vp, _vp, _vp, _i32, _i32, _vp, _vp]
    lib.cy2p_propagate_rows_persist.argtypes = [_vp, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]
    lib.cy2p_propagate_rows_persist_f64.argtypes = lib.cy2p_propagate_rows_persist.argtypes
    lib.cy2p_propagate_x.argtypes = [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]
    lib.cy2p_propagate_rows_ordered.argtypes = lib.cy2p_propagate_rows.argtypes
    lib.cy2p_propagate_rows_ordered_f64.argtypes = lib.cy2p_propagate_rows.argtypes
    lib.cy2p_propagate_rows_allgather.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]
    lib.cy2p_propagate_rows_allgather_f64.argtypes = lib.cy2p_propagate_rows_allgather.argtypes
    lib.cy2p_propagate_rows_push.argtypes = [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp]
    lib.cy2p_propagate_rows_push_f64.argtypes = lib.cy2p_propagate_rows_push.argtypes
    lib.cy2p_enable_peer_access.argtypes = [_i32, _i32]
    lib.cy2p_cdf_build.argtypes = [_vp, _vp, _vp, _i32, _vp]
    lib.cy2p_sample_chains.argtypes = [_vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]
    lib.cy2p_chain_history.argtypes = [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]
    if hasattr(lib, "cy2p_fhmm_forward"):
        lib.cy2p_fhmm_ws_bytes.argtypes = [_i32, _i32, _i32, _i32]
        lib.cy2p_fhmm_ws_bytes.restype = _sz
        lib.cy2p_fhmm_small_floats.argtypes = [_i32, _i32]
        lib.cy2p_fhmm_small_floats.restype = _i32
        lib.cy2p_fhmm_forward.argtypes = [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                          _vp, _sz, _vp]
        lib.cy2p_fhmm_loss_grad.argtypes = [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                                            ctypes.POINTER(ctypes.c_float), _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
    if hasattr(lib, "cy2p_fhmm_decode"):
        lib.cy2p_fhmm_decode.argtypes = [_i32] * 5 + [_vp] * 14 + [_sz, _vp]
    if hasattr(lib, "cy2p_lssm_forward"):
        lib.cy2p_lssm_ws_bytes.argtypes = lib.cy2p_fhmm_ws_bytes.argtypes
        lib.cy2p_lssm_ws_bytes.restype = _sz
        lib.cy2p_lssm_small_floats.argtypes = [_i32, _i32]
        lib.cy2p_lssm_small_floats.restype = _i32
        lib.cy2p_lssm_forward.argtypes = lib.cy2p_fhmm_forward.argtypes
        lib.cy2p_lssm_loss_grad.argtypes = lib.cy2p_fhmm_loss_grad.argtypes
        lib.cy2p_lssm_decode.argtypes = lib.cy2p_fhmm_decode.argtypes
    if hasattr(lib, "cy2p_nssm_forward"):
        lib.cy2p_nssm_ws_bytes.argtypes = lib.cy2p_fhmm_ws_bytes.argtypes
        lib.cy2p_nssm_ws_bytes.restype = _sz
        lib.cy2p_nssm_forward.argtypes = [_i32] * 5 + [_vp] * 11 + [_sz, _vp]
        lib.cy2p_nssm_loss_grad.argtypes = lib.cy2p_fhmm_loss_grad.argtypes
        lib.cy2p_nssm_decode.argtypes = lib.cy2p_fhmm_decode.argtypes
    if hasattr(lib, "cy2p_dtw_matrix"):
        lib.cy2p_traj_ws_bytes.argtypes = [_i32, _i32]
        lib.cy2p_traj_ws_bytes.restype = _sz
        lib.cy2p_dtw_matrix.argtypes = [_vp, _i32, _i32, _i32, _vp, _vp, _sz, _vp]
        lib.cy2p_hausdorff_matrix.argtypes = lib.cy2p_dtw_matrix.argtypes
    for name in SYMBOLS:
        fn = getattr(lib, name, None)
        if fn is not None and name not in ("cy2p_last_error", "cy2p_propagate_ws_bytes", "cy2p_fhmm_ws_bytes",
                                           "cy2p_fhmm_small_floats", "cy2p_abi_version", "cy2p_launch_count",
                                           "cy2p_propagate_f64_ws_bytes", "cy2p_lssm_ws_bytes",
                                           "cy2p_lssm_small_floats", "cy2p_nssm_ws_bytes", "cy2p_traj_ws_bytes"):
            fn.restype = _i32
    _lib = lib
    return lib


def check(status):
    """Map a C status to the exception type the reference raises in the same situation."""
    if status == OK:
        return
    msg = load().cy2p_last_error().decode("utf-8", "replace")
    if status in (ERR_CDF_OVERSHOOT, ERR_CDF_TABLE):
        raise IndexError(msg)  # reference: simulation.py:78-81 / :43-46
    if status in (ERR_INVALID_ARG, ERR_BAD_CSR):
        raise ValueError(msg)
    raise Cy2pError(f"status {status}: {msg}")


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise Cy2pError("cy2path_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch


def stream_ptr(device=None):
    torch = _torch()
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def to_device(a, dtype, dev):
    """Host array (numpy) or tensor -> contiguous device tensor of `dtype`. Pinned host memory is copied
    asynchronously at PCIe speed; pageable memory goes through the driver's staging copy (no extra pinning pass)."""
    torch = _torch()
    if torch.is_tensor(a):
        return a.to(dev, dtype=dtype).contiguous()
    host = torch.from_numpy(np.ascontiguousarray(a))
    if host.dtype != dtype and not host.is_pinned():
        host = host.to(dtype)  # convert on the host once rather than moving the wider type
    return host.to(dev, non_blocking=host.is_pinned()).to(dtype).contiguous()


def to_host(t):
    """Device tensor -> numpy array backed by pinned host memory (torch caches the pinned blocks, so repeated
    calls copy at PCIe speed instead of through a pageable staging buffer)."""
    torch = _torch()
    if t.numel() * t.element_size() < (1 << 20):
        return t.cpu().numpy()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out.numpy()


def to_host_many(tensors):
    """to_host for several device tensors with one stream synchronisation (None entries pass through)."""
    torch = _torch()
    outs, dev = [], None
    for t in tensors:
        if t is None:
            outs.append(None)
            continue
        dev = t.device
        o = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        o.copy_(t, non_blocking=True)
        outs.append(o)
    if dev is not None:
        torch.cuda.current_stream(dev).synchronize()
    return [None if o is None else o.numpy() for o in outs]


# ------------------------------------------------------------------------------------------ guarded device buffers
# compute-sanitizer is not available on the GPU pool this was developed on, so out-of-bounds WRITES are hunted with
# canaries instead: with CY2P_GUARD=1 every device buffer the wrappers hand to the library (outputs and work spaces) is
# carved out of a larger allocation with 64 KB of 0xA5 on either side, and check_guards() (after a synchronise) raises
# if a single guard byte changed. tests/test_guards_gpu.py runs odd shapes of every entry point this way. Off by default:
# one environment lookup per allocation.
_GUARD_BYTES = 1 << 16
_GUARDS = []


def dev_empty(shape, dtype, dev, what="buffer"):
    """torch.empty on the device, or — under CY2P_GUARD=1 — a view into a canary-fenced allocation."""
    torch = _torch()
    shape = tuple(int(v) for v in shape)
    if os.environ.get("CY2P_GUARD", "0") in ("", "0"):
        return torch.empty(shape, dtype=dtype, device=dev)
    nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
    pad = (-nbytes) % 256
    flat = torch.full((2 * _GUARD_BYTES + nbytes + pad,), 0xA5, dtype=torch.uint8, device=dev)
    _GUARDS.append((what, flat[:_GUARD_BYTES], flat[_GUARD_BYTES + nbytes:], flat))
    return flat[_GUARD_BYTES:_GUARD_BYTES + nbytes].view(dtype).view(shape)


def check_guards():
    """Synchronise and verify every canary handed out since the last call; raises RuntimeError on a changed byte."""
    torch = _torch()
    bad = []
    for what, lo, hi, _keep in _GUARDS:
        torch.cuda.synchronize(lo.device)
        if not bool((lo == 0xA5).all()):
            bad.append("%s: bytes BEFORE the buffer were written" % what)
        if not bool((hi == 0xA5).all()):
            bad.append("%s: bytes AFTER the buffer were written" % what)
    n = len(_GUARDS)
    _GUARDS.clear()
    if bad:
        raise RuntimeError("out-of-bounds device writes: " + "; ".join(bad))
    return n


class Plan:
    """Device-resident T and T^T for one transition matrix on one device (cy2p_plan)."""

    def __init__(self, T, device=None):
        torch = _torch()
        lib = load()
        import scipy.sparse as sp

        if not sp.issparse(T):
            T = sp.csr_matrix(T)
        T = T.tocsr()
        self.n = int(T.shape[0])
        if T.shape[0] != T.shape[1]:
            raise ValueError("transition matrix must be square")
        self.nnz = int(T.nnz)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        indptr = np.ascontiguousarray(T.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(T.indices, dtype=np.int32)
        # A float64 matrix whose values are not float32-representable is held as hi + lo float32 pairs (48 mantissa
        # bits): the float64-arithmetic paths (single-start float64, batched f48 / f40) and the CDF tables then follow
        # the reference's float64 product / cumsum (sampling.py:40-42, simulation.py:24-25). The float32 / legacy fp64
        # batched kernels, the row-partitioned run and the FHMM operand use hi alone.
        self.float64_matrix = False
        handle = _vp(0)
        if np.asarray(T.data).dtype == np.float64:
            data64 = np.ascontiguousarray(T.data, dtype=np.float64)
            self.float64_matrix = not np.array_equal(data64.astype(np.float32).astype(np.float64), data64)
        if self.float64_matrix:
            with torch.cuda.device(self.device):
                check(lib.cy2p_plan_create_host_f64(self.n, self.nnz, indptr.ctypes.data, indices.ctypes.data,
                                                    data64.ctypes.data, self.device.index, stream_ptr(self.device),
                                                    ctypes.byref(handle)))
        else:
            data = np.ascontiguousarray(T.data, dtype=np.float32)
            with torch.cuda.device(self.device):
                check(lib.cy2p_plan_create_host(self.n, self.nnz, indptr.ctypes.data, indices.ctypes.data,
                                                data.ctypes.data, self.device.index, stream_ptr(self.device),
                                                ctypes.byref(handle)))
        self.handle = handle
        info = (_i64 * 8)()
        check(lib.cy2p_plan_info(self.handle, info))
        self.max_in_degree, self.max_out_degree, self.W = int(info[2]), int(info[3]), int(info[4])
        self.bytes_owned = int(info[7])

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.cy2p_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def propagate(plan, X0, iters, stationary=None, want_final=True, history_stride=0, want_eps=False):
    """Run cy2p_propagate (float32 X0) or cy2p_propagate_f64 (float64 X0). X0: cuda [n, B] (or [n]).
    Returns dict(final, history, eps) of cuda tensors in X0's dtype (eps is always float64)."""
    torch = _torch()
    lib = load()
    squeeze = X0.dim() == 1
    X0 = X0.reshape(plan.n, -1).contiguous()
    assert X0.dtype in (torch.float32, torch.float64) and X0.is_cuda
    f64 = X0.dtype == torch.float64
    fn, ws_fn = (lib.cy2p_propagate_f64, lib.cy2p_propagate_f64_ws_bytes) if f64 else \
        (lib.cy2p_propagate, lib.cy2p_propagate_ws_bytes)
    B = X0.shape[1]
    dev = X0.device
    with torch.cuda.device(dev):
        st = None
        if stationary is not None:
            st = torch.as_tensor(np.asarray(stationary, dtype=np.float64) if not torch.is_tensor(stationary)
                                 else stationary, dtype=torch.float64, device=dev).contiguous()
        final = dev_empty(X0.shape, X0.dtype, dev, "propagate final") if want_final else None
        hist = None
        if history_stride:
            slots = (iters + history_stride - 1) // history_stride
            hist = dev_empty((max(slots, 1), plan.n, B), X0.dtype, dev, "propagate history")
        eps = dev_empty((iters, B), torch.float64, dev, "propagate eps") if want_eps else None
        ws_bytes = ws_fn(plan.handle, B, iters)
        ws = dev_empty((ws_bytes,), torch.uint8, dev, "propagate work space")
        check(fn(plan.handle, ptr(X0), B, iters, ptr(st), ptr(final), ptr(hist), max(history_stride, 1), ptr(eps),
                 ptr(ws), ws_bytes, stream_ptr(dev)))
    if squeeze:
        final = final.reshape(-1) if final is not None else None
        hist = hist.reshape(hist.shape[0], plan.n) if hist is not None else None
        eps = eps.reshape(-1) if eps is not None else None
    return {"final": final, "history": hist, "eps": eps}


def propagate_x(plan, X0, iters, stationary=None, tail_bits=16, want_eps=False, out_dtype=None):
    """Run cy2p_propagate_x: batched propagation with the extended-precision ("f48" / "f40") iterate.
    X0: cuda float32 [n, B]. Returns dict(final, eps): final is float32 (or float64 with out_dtype=torch.float64),
    eps float64 [iters, B] (bit-reproducible)."""
    torch = _torch()
    lib = load()
    X0 = X0.reshape(plan.n, -1).contiguous()
    assert X0.dtype == torch.float32 and X0.is_cuda
    B = X0.shape[1]
    dev = X0.device
    f64_out = out_dtype == torch.float64
    with torch.cuda.device(dev):
        st = None
        if stationary is not None:
            st = torch.as_tensor(np.asarray(stationary, dtype=np.float64) if not torch.is_tensor(stationary)
                                 else stationary, dtype=torch.float64, device=dev).contiguous()
        want_eps = bool(want_eps and st is not None)
        final = dev_empty((plan.n, B), torch.float64 if f64_out else torch.float32, dev, "propagate_x final")
        eps = dev_empty((iters, B), torch.float64, dev, "propagate_x eps") if want_eps else None
        ws_bytes = lib.cy2p_propagate_x_ws_bytes(plan.handle, B, iters, int(want_eps), tail_bits)
        ws = dev_empty((ws_bytes,), torch.uint8, dev, "propagate_x work space")
        check(lib.cy2p_propagate_x(plan.handle, ptr(X0), B, iters, ptr(st), None if f64_out else ptr(final),
                                   ptr(final) if f64_out else None, ptr(eps), tail_bits, ptr(ws), ws_bytes,
                                   stream_ptr(dev)))
    return {"final": final, "history": None, "eps": eps}


# ------------------------------------------------------------------------------------------ K3 (FHMM)
def _model_dims(model, theta_pi, theta_w):
    """(S, L) of the FHMM (theta_pi [S,L], theta_w [L]) or the LSSM (theta_pi [S], theta_w [S,L])."""
    if model == "fhmm":
        return int(theta_pi.shape[0]), int(theta_pi.shape[1])
    if model == "lssm":
        return int(theta_w.shape[0]), int(theta_w.shape[1])
    if model == "nssm":  # theta_pi [S], theta_w [1|S, N, L]
        return int(theta_pi.shape[0]), int(theta_w.shape[-1])
    raise ValueError("model must be 'fhmm', 'lssm' or 'nssm'")


def _fhmm_ws(S, L, N, I, dev, model="fhmm"):
    torch = _torch()
    nbytes = getattr(load(), "cy2p_%s_ws_bytes" % model)(S, L, N, I)
    if nbytes == 0:
        raise ValueError("unsupported model size (supported: S <= 32, L <= 16, S*(1 if restricted else L) <= 64; "
                         "NSSM: S*L <= 64)")
    return dev_empty((nbytes,), torch.uint8, dev, "%s work space" % model), nbytes


def fhmm_forward(theta_pi, theta_w, theta_E, theta_A, num_iters, restricted, want_pred=True, want_joint=False,
                 model="fhmm"):
    """cy2p_fhmm_forward / cy2p_lssm_forward on cuda float32 parameter tensors. Returns dict of cuda tensors:
    pred [I,N] | None, hidden [I,S,L] ([I,S] for the LSSM), joint [I,S,L,N] | None, log_E [S,Le,N], log_pi, log_w,
    log_A, log_A_mix (the LSSM has one transition matrix: log_A_mix is log_A)."""
    torch = _torch()
    lib = load()
    S, L = _model_dims(model, theta_pi, theta_w)
    I = int(num_iters)
    dev = theta_pi.device
    tp, tw, tE, tA = (t.detach().contiguous().float() for t in (theta_pi, theta_w, theta_E, theta_A))
    if model == "nssm":
        N = theta_E.shape[1]
        with torch.cuda.device(dev):
            ws, nbytes = _fhmm_ws(S, L, N, I, dev, model)
            f = lambda *shape: dev_empty(shape, torch.float32, dev, "model output")  # noqa: E731
            pred = f(I, N) if want_pred else None
            hidden, joint = f(I, S), (f(I, S, L, N) if want_joint else None)
            logE, logcw, small = f(S, N), f(*tw.shape), f(lib.cy2p_lssm_small_floats(S, L))
            check(lib.cy2p_nssm_forward(S, L, N, I, int(bool(restricted)), ptr(tp), ptr(tw), ptr(tE), ptr(tA), ptr(pred),
                                        ptr(hidden), ptr(joint), ptr(logE), ptr(logcw), ptr(small), ptr(ws), nbytes,
                                        stream_ptr(dev)))
        return {"pred": pred, "hidden": hidden, "joint": joint, "log_E": logE, "log_pi": small[:S], "log_w": logcw,
                "log_r": small[S:S + S * L].view(S, L), "log_A": small[S + S * L:].view(S, S),
                "log_A_mix": small[S + S * L:].view(S, S)}
    Le, N = theta_E.shape[1], theta_E.shape[2]
    with torch.cuda.device(dev):
        ws, nbytes = _fhmm_ws(S, L, N, I, dev, model)
        pred = dev_empty((I, N), torch.float32, dev, "forward pred") if want_pred else None
        hidden = dev_empty((I, S, L) if model == "fhmm" else (I, S), torch.float32, dev, "forward hidden")
        joint = dev_empty((I, S, L, N), torch.float32, dev, "forward joint") if want_joint else None
        logE = dev_empty((S, Le, N), torch.float32, dev, "forward logE")
        small = dev_empty((getattr(lib, "cy2p_%s_small_floats" % model)(S, L),), torch.float32, dev, "forward small")
        fwd = getattr(lib, "cy2p_%s_forward" % model)
        check(fwd(S, L, N, I, int(bool(restricted)), ptr(tp), ptr(tw), ptr(tE), ptr(tA), ptr(pred),
                  ptr(hidden), ptr(joint), ptr(logE), ptr(small), ptr(ws), nbytes, stream_ptr(dev)))
    if model == "lssm":
        log_pi, log_w, log_A = small[:S], small[S:S + S * L].view(S, L), small[S + S * L:].view(S, S)
        return {"pred": pred, "hidden": hidden, "joint": joint, "log_E": logE, "log_pi": log_pi, "log_w": log_w,
                "log_A": log_A, "log_A_mix": log_A}
    o = 0
    log_pi = small[o:o + S * L].view(S, L); o += S * L
    log_w = small[o:o + L]; o += L
    log_A = small[o:o + L * S * S].view(L, S, S); o += L * S * S
    log_A_mix = small[o:o + S * S].view(S, S)
    return {"pred": pred, "hidden": hidden, "joint": joint, "log_E": logE, "log_pi": log_pi, "log_w": log_w,
            "log_A": log_A, "log_A_mix": log_A_mix}


TERM_NAMES = ("loss", "divergence", "log_likelihood", "sparsity", "orthogonality", "exclusivity", "regularisation")


def fhmm_loss_grad(plan, theta_pi, theta_w, theta_E, theta_A, D, restricted, weights, out=None, terms=None,
                   model="fhmm"):
    """cy2p_fhmm_loss_grad / cy2p_lssm_loss_grad: one epoch's forward + loss + analytic gradients.
    weights = (sparsity, exclusivity, orthogonality, TPM). Returns (terms [8] cuda, (g_pi, g_w, g_E, g_A))."""
    torch = _torch()
    lib = load()
    S, L = _model_dims(model, theta_pi, theta_w)
    N = theta_E.shape[-1]
    I = int(D.shape[0])
    dev = theta_pi.device
    tp, tw, tE, tA = (t.detach().contiguous() for t in (theta_pi, theta_w, theta_E, theta_A))
    assert D.is_cuda and D.dtype == torch.float32 and D.is_contiguous() and D.shape[1] == N
    with torch.cuda.device(dev):
        if out is None:
            ws, nbytes = _fhmm_ws(S, L, N, I, dev, model)
            out = {"ws": ws, "nbytes": nbytes, "terms": dev_empty((8,), torch.float32, dev, "loss terms"),
                   "g": tuple(dev_empty(t.shape, t.dtype, dev, "gradient %d" % i) for i, t in enumerate((tp, tw, tE, tA)))}
        hw = (ctypes.c_float * 4)(*[float(x) for x in weights])
        g = out["g"]
        tdst = out["terms"] if terms is None else terms  # 8 contiguous float32 (e.g. a row of the epoch history)
        fn = getattr(lib, "cy2p_%s_loss_grad" % model)
        check(fn(plan.handle, S, L, N, I, int(bool(restricted)), ptr(tp), ptr(tw), ptr(tE), ptr(tA), ptr(D), hw,
                 ptr(tdst), ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(out["ws"]), out["nbytes"],
                 stream_ptr(dev)))
    return tdst, g, out


def fhmm_decode(theta_pi, theta_w, theta_E, theta_A, D, restricted, model="fhmm"):
    """cy2p_fhmm_decode / cy2p_lssm_decode: filtering, smoothing and viterbi in one call. Returns a dict of cuda
    tensors."""
    torch = _torch()
    lib = load()
    S, L = _model_dims(model, theta_pi, theta_w)
    N = theta_E.shape[-1]
    I = int(D.shape[0])
    dev = theta_pi.device
    tp, tw, tE, tA = (t.detach().contiguous().float() for t in (theta_pi, theta_w, theta_E, theta_A))
    D = D.to(dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        ws, nbytes = _fhmm_ws(S, L, N, I, dev, model)
        f = lambda *shape: dev_empty(shape, torch.float32, dev, "model output")  # noqa: E731
        out = {"log_alpha": f(I, S, L), "log_probs": f(I, L), "log_beta": f(I, S, L), "log_gamma": f(I, S, L),
               "log_delta": f(I, S, L), "psi": f(I, S, L), "log_max": f(I, L), "best_path": f(L, I)}
        fn = getattr(lib, "cy2p_%s_decode" % model)
        check(fn(S, L, N, I, int(bool(restricted)), ptr(tp), ptr(tw), ptr(tE), ptr(tA), ptr(D),
                 ptr(out["log_alpha"]), ptr(out["log_probs"]), ptr(out["log_beta"]), ptr(out["log_gamma"]),
                 ptr(out["log_delta"]), ptr(out["psi"]), ptr(out["log_max"]), ptr(out["best_path"]), ptr(ws), nbytes,
                 stream_ptr(dev)))
    return out


# ------------------------------------------------------------------------------------------ K4 (trajectory distances)
def trajectory_distance_matrix(series, kind="dtw"):
    """cy2p_dtw_matrix / cy2p_hausdorff_matrix. ``series``: cuda float64 tensor [C, len, d] (d <= 64).
    Returns a cuda float64 tensor [C, C] (symmetric, zero diagonal)."""
    torch = _torch()
    lib = load()
    if kind not in ("dtw", "hausdorff"):
        raise ValueError("kind must be 'dtw' or 'hausdorff'")
    if series.dim() != 3 or not series.is_cuda:
        raise ValueError("series must be a cuda tensor [C, len, d]")
    C, ln, d = (int(v) for v in series.shape)
    if d > 64:
        raise ValueError("at most 64 dimensions per point are supported (reduce the basis, e.g. fewer principal components)")
    dpad = max(8, (d + 7) // 8 * 8)
    dev = series.device
    with torch.cuda.device(dev):
        s = torch.zeros((C, ln, dpad), dtype=torch.float64, device=dev)
        s[:, :, :d] = series.to(torch.float64)
        out = dev_empty((C, C), torch.float64, dev, "distance matrix")
        if C == 0 or ln == 0:
            return out.zero_()
        nbytes = lib.cy2p_traj_ws_bytes(C, ln)
        ws = dev_empty((nbytes,), torch.uint8, dev, "trajectory work space")
        fn = lib.cy2p_dtw_matrix if kind == "dtw" else lib.cy2p_hausdorff_matrix
        check(fn(ptr(s), C, ln, dpad, ptr(out), ptr(ws), nbytes, stream_ptr(dev)))
    return out
