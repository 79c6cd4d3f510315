"""Lazily evaluated ``joint_probabilities``.

The reference stores ``exp(joint)`` with joint [I, S, L, N] on the host (infer_dynamics.py:37-39): 4 GB at
BASELINE config 3, only for ``.mean(0)`` and ``.sum(-1).argmax(1)`` to be taken later (infer_posthoc.py:27-31,
75-80). Every model class factorises the joint as

    joint[t, s, l, n] = static[s, l, n] + hidden[t, s, l]

(static = log E + log chain weights, hidden = the latent recurrence), so the object below keeps the two factors
(8 MB + 80 KB at config 3) and evaluates the reductions in closed form; indexing / ``np.asarray`` materialise on
demand. It quacks like the ndarray the reference's post-hoc functions expect.
"""
from __future__ import annotations

import numpy as np


class JointProbabilities:
    """exp(static[None] + hidden[..., None]) without storing it. ``static`` [S, L, N], ``hidden`` [I, S, L], both
    natural logs (float32 numpy)."""

    ndim = 4
    dtype = np.dtype(np.float32)

    def __init__(self, static_log, hidden_log):
        self.static_log = np.ascontiguousarray(static_log, dtype=np.float32)
        self.hidden_log = np.ascontiguousarray(hidden_log, dtype=np.float32)
        assert self.static_log.ndim == 3 and self.hidden_log.ndim == 3
        assert self.static_log.shape[:2] == self.hidden_log.shape[1:]

    @property
    def shape(self):
        return (self.hidden_log.shape[0],) + self.static_log.shape

    @property
    def nbytes(self):
        return int(np.prod(self.shape)) * 4

    def __len__(self):
        return self.shape[0]

    def _rows(self, t_index):
        h = self.hidden_log[t_index]
        return np.exp(self.static_log[None] + h[..., None]) if h.ndim == 3 else np.exp(self.static_log + h[..., None])

    def __array__(self, dtype=None, copy=None):
        out = self._rows(slice(None))
        return out if dtype is None else out.astype(dtype, copy=False)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if idx and idx[0] is Ellipsis:
            return np.asarray(self)[idx]
        first, rest = idx[0], idx[1:]
        block = self._rows(first)
        if not rest:
            return block
        if isinstance(first, (int, np.integer)):
            return block[rest]
        return block[(slice(None),) + rest]

    # the two reductions the reference's post-hoc code takes (infer_posthoc.py:29-31, 75-80), in closed form
    def mean(self, axis=None, **kw):
        if axis in (0, -4):
            hbar = np.exp(self.hidden_log.astype(np.float64)).mean(0)  # [S, L]
            return (np.exp(self.static_log.astype(np.float64)) * hbar[..., None]).astype(np.float32)
        return np.asarray(self).mean(axis=axis, **kw)

    def sum(self, axis=None, **kw):
        if axis in (-1, 3):
            rows = np.exp(self.static_log.astype(np.float64)).sum(-1)  # [S, L]
            return (np.exp(self.hidden_log.astype(np.float64)) * rows[None]).astype(np.float32)
        return np.asarray(self).sum(axis=axis, **kw)

    def __repr__(self):
        return "JointProbabilities(shape=%s, lazy; %.1f MB if materialised)" % (self.shape, self.nbytes / 1e6)
