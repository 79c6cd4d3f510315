"""cy2path_b200 — B200-native implementation of cy2path's data-parallel hot path.

Drop-in for the reference's ``cy2path.sampling`` / ``cy2path.simulation`` / ``cy2path.models`` API
(``import cy2path_b200 as cy2path``). All arithmetic runs in hand-written sm_100a CUDA
(cy2path_b200/csrc, C ABI in include/cy2path_b200.h); there is no CPU fallback.
"""
from .sampling import (check_convergence_criteria, iterate_state_probability, propagate_batch,  # noqa: F401
                       sample_state_probability)
from .simulation import extract_nonzero_entries, iterate_markov_chain, sample_markov_chains  # noqa: F401

__version__ = "0.1.0"

__all__ = [
    "iterate_state_probability", "sample_state_probability", "check_convergence_criteria", "propagate_batch",
    "extract_nonzero_entries", "iterate_markov_chain", "sample_markov_chains", "FHMM", "LSSM", "NSSM", "infer_dynamics", "cluster_markov_chains",
]


def __getattr__(name):  # lazy: torch.nn is only imported when the model is used
    if name in ("FHMM", "LSSM", "NSSM"):
        from . import models

        return getattr(models, name)
    if name == "infer_dynamics":
        from .dynamics import infer_dynamics

        return infer_dynamics
    if name == "cluster_markov_chains":
        from .cytopath import cluster_markov_chains

        return cluster_markov_chains
    raise AttributeError(name)
