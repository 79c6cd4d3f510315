"""TEST INFRASTRUCTURE — loads the *unmodified* reference hot-path modules.

Only usable where ``/root/reference`` exists (the build container). It is used
by ``tests/golden/make_golden.py`` to produce the committed golden vectors and
by the ``not gpu`` tests that pin ``oracle/`` against the reference when the
tree is present. Nothing in the product package, ``bench.py`` or the ``-m gpu``
tests imports this file: the reference does not travel to the GPU box.

How it works: ``import cy2path`` fails here because the package
``__init__`` pulls in plotting / dtaidistance / scvelo (not installed). The
hot-path modules themselves only need numpy / scipy / torch / joblib / tqdm, so
we register empty stubs for the ``scvelo`` names that ``cy2path/utils.py:9-11``
imports, pre-register a bare ``cy2path`` package object whose ``__path__``
points at the reference directory (so ``cy2path/__init__.py`` never runs), and
import ``cy2path.sampling`` / ``.simulation`` / ``.models`` from the source
files where they lie.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CY2PATH_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "cy2path")):
    # the GPU box: the unmodified copies oracle/build_ref.py made in the build container (git-ignored build output)
    REFERENCE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cy2path"))


class DuckAnnData:
    """Minimal AnnData stand-in: the hot path only touches .obsp/.obs/.uns/.shape/.copy()."""

    def __init__(self, T, obs=None):
        import pandas as pd

        self.obsp = {"T_forward": T}
        self.shape = (T.shape[0], 0)
        self.obs = obs if obs is not None else pd.DataFrame(index=range(T.shape[0]))
        self.uns = {}

    def copy(self):
        import copy as _copy

        return _copy.deepcopy(self)


def load():
    """Return a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "cy2path" in sys.modules and getattr(sys.modules["cy2path"], "_b200_shim", False):
        pkg = sys.modules["cy2path"]
    else:
        def _stub(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules.setdefault(name, m)
            return sys.modules[name]

        def _missing(*a, **k):
            raise RuntimeError("scvelo is not installed; this branch of the reference cannot run here")

        sv = _stub("scvelo")
        tools = _stub("scvelo.tools", terminal_states=_missing)
        ts = _stub("scvelo.tools.terminal_states", eigs=_missing)
        ut = _stub("scvelo.utils", get_transition_matrix=_missing)
        sv.tools, sv.utils = tools, ut
        tools.terminal_states_module = ts

        pkg = types.ModuleType("cy2path")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "cy2path")]
        pkg._b200_shim = True
        sys.modules["cy2path"] = pkg

    ns = types.SimpleNamespace()
    ns.utils = importlib.import_module("cy2path.utils")
    ns.sampling = importlib.import_module("cy2path.sampling")
    ns.simulation = importlib.import_module("cy2path.simulation")
    ns.models = importlib.import_module("cy2path.models")
    ns.methods = importlib.import_module("cy2path.models.methods")
    ns.trainer = importlib.import_module("cy2path.models.trainer")
    ns.DuckAnnData = DuckAnnData
    return ns
