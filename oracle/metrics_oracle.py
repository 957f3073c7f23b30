"""Oracle: evaluation metrics (numpy float64).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Restatement of ``scripts/validation_metrics.py``: TM-score
(:23-54), kabsch_align (:57-85), lDDT (:92-149), GDT (:156-199), RMSF (:206-241); pinned on ``tests/golden/metrics.npz``,
which ``tests/golden/make_golden.py::gen_metrics`` produces by importing that reference module."""
from __future__ import annotations

import numpy as np


def kabsch_align(mobile, target):
    """:57-85 (all residues, proper rotation)."""
    mc, tc = mobile - mobile.mean(0), target - target.mean(0)
    U, S, Vt = np.linalg.svd(mc.T @ tc)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    return mc @ R.T + target.mean(0)


def tm_score(pred, true):
    """:42-52."""
    L = len(true)
    d0 = 1.24 * np.cbrt(L - 15) - 1.8
    d = np.linalg.norm(kabsch_align(pred, true) - true, axis=1)
    return float(np.mean(1.0 / (1.0 + (d / d0) ** 2)))


def gdt(pred, true, mask=None):
    """:176-199."""
    mask = np.ones(len(true), bool) if mask is None else np.asarray(mask).astype(bool)
    d = np.linalg.norm(kabsch_align(pred, true) - true, axis=1)[mask]
    if len(d) == 0:
        return 0.0, 0.0
    f = lambda t: (d < t).mean() * 100  # noqa: E731
    return float((f(1) + f(2) + f(4) + f(8)) / 4), float((f(0.5) + f(1) + f(2) + f(4)) / 4)


def lddt(pred, true, mask=None, cutoff=15.0):
    """:111-149."""
    L = len(true)
    mask = np.ones(L, bool) if mask is None else np.asarray(mask).astype(bool)
    dt = np.linalg.norm(true[:, None] - true[None], axis=-1)
    dp = np.linalg.norm(pred[:, None] - pred[None], axis=-1)
    out = np.zeros(L)
    for i in range(L):
        if not mask[i]:
            continue
        nb = (dt[i] < cutoff) & (dt[i] > 0) & mask
        if nb.sum() == 0:
            continue
        dd = np.abs(dt[i, nb] - dp[i, nb])
        out[i] = ((dd < 0.5).sum() + (dd < 1.0).sum() + (dd < 2.0).sum() + (dd < 4.0).sum()) / (4 * nb.sum())
    return (float(out[mask].mean()) if mask.sum() > 0 else 0.0), out


def rmsf(ensemble):
    """:224-241."""
    N = ensemble.shape[0]
    if N == 1:
        return np.zeros(ensemble.shape[1])
    al = np.stack([kabsch_align(ensemble[i], ensemble[0]) for i in range(N)])
    return np.sqrt(((al - al.mean(0)) ** 2).sum(-1).mean(0))
