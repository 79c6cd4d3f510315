"""Drop-in mirror of cy2path/models — the latent dynamic models on the B200.

``FHMM``, ``LSSM`` and ``NSSM`` keep the reference's constructors, parameter names/shapes (state_dict compatible),
``forward_model`` / ``filtering`` / ``smoothing`` / ``viterbi`` / ``train`` methods and training-history
attributes; the arithmetic runs in csrc/fhmm.cu.
"""
from ._FHMM import FHMM  # noqa: F401
from ._LSSM import LSSM  # noqa: F401
from ._NSSM import NSSM  # noqa: F401

__all__ = ["FHMM", "LSSM", "NSSM"]

