"""TEST INFRASTRUCTURE — builds oracle/csrc/oracle.c with gcc into oracle/_build/liboracle.so."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "oracle.c")
_OUT = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        # -ffp-contract=off: no FMA contraction, the reference's numpy/scipy code has none
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _OUT, _SRC, "-lm"])
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_cosine.restype = ctypes.c_double
    return _lib
