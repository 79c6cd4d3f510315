"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports every declared symbol."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cy2path_b200 import _build, _lib

    _build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cy2path_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cy2p_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from cy2path_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"library does not export: {missing}"
    assert sorted(_lib.SYMBOLS) == declared


def test_abi_version_and_error_string(lib):
    assert lib.cy2p_abi_version() == 2
    assert isinstance(lib.cy2p_last_error(), bytes)


def test_null_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call
    out = ctypes.c_void_p(0)
    rc = lib.cy2p_plan_create_host(0, 0, None, None, None, 0, None, ctypes.byref(out))
    assert rc == 1 and b"invalid argument" in lib.cy2p_last_error()
    assert lib.cy2p_propagate_ws_bytes(None, 4, 10) == 0


def test_sass_is_sm100a(lib):
    import subprocess

    from cy2path_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cy2path_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
