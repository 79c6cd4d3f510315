"""Drop-in mirror of cy2path/models — the latent dynamic models on the B200.

``FHMM`` and ``LSSM`` keep the reference's constructors, parameter names/shapes (state_dict compatible),
``forward_model`` / ``filtering`` / ``smoothing`` / ``viterbi`` / ``train`` methods and training-history
attributes; the arithmetic runs in csrc/fhmm.cu. NSSM is the remaining row of SURVEY.md §8(f-3).
"""
from ._FHMM import FHMM  # noqa: F401
from ._LSSM import LSSM  # noqa: F401

__all__ = ["FHMM", "LSSM"]

