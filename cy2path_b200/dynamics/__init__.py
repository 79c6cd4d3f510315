"""Mirror of cy2path/dynamics for the rows of SURVEY.md §8 f-2: ``infer_dynamics`` / ``extract_model_outputs``
(reference dynamics/infer_dynamics.py) and the conditional / state-selection helpers of
dynamics/infer_posthoc.py:10-125. Joint training, kinetic clusters and latent paths are out of scope."""
from .infer_dynamics import extract_model_outputs, infer_dynamics  # noqa: F401
from .infer_posthoc import compute_conditionals, latent_state_selection  # noqa: F401
from .lazy import JointProbabilities  # noqa: F401

__all__ = ["infer_dynamics", "extract_model_outputs", "compute_conditionals", "latent_state_selection",
           "JointProbabilities"]
