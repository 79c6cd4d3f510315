"""Drop-in mirror of cy2path/models — the factorial latent dynamic model on the B200.

``FHMM`` keeps the reference's constructor, parameter names/shapes (state_dict compatible),
``forward_model`` / ``train`` methods and training-history attributes; the arithmetic runs in
csrc/fhmm.cu. LSSM / NSSM and the decoders' CUDA kernels are the next rows of SURVEY.md §8(f).
"""
from ._FHMM import FHMM  # noqa: F401

__all__ = ["FHMM"]

