"""Kabsch RMSD oracle (numpy).  TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

``kabsch_rmsd_ref_compat`` restates ``generate_ensemble_pdbs.py:343-373`` of the
reference as written, including its use of the inverse rotation (``c1 @ R`` with
``R = V D U^T``).  ``kabsch_rmsd`` is the optimal-superposition
RMSD, equal to ``scripts/validation_metrics.py:57-85`` (``kabsch_align``)
followed by a per-residue RMSD, and to the singular-value closed form.
"""
from __future__ import annotations

import numpy as np


def _centered(c1, c2, mask):
    valid = np.asarray(mask).astype(bool)
    a = np.asarray(c1)[valid]
    b = np.asarray(c2)[valid]
    if len(a) == 0:
        return None, None
    return a - a.mean(axis=0), b - b.mean(axis=0)


def kabsch_rmsd_ref_compat(coords1, coords2, mask) -> float:
    """As the reference computes it (``generate_ensemble_pdbs.py:343-373``)."""
    a, b = _centered(coords1, coords2, mask)
    if a is None:
        return 0.0                                            # :350-351
    U, _, Vt = np.linalg.svd(a.T @ b)                         # :358-359
    V = Vt.T
    d = np.sign(np.linalg.det(V @ U.T))                       # :364
    R = V @ np.diag([1.0, 1.0, d]) @ U.T                      # :365-366
    moved = a @ R                                             # :369 (inverse rotation, F6)
    return float(np.sqrt(((moved - b) ** 2).sum() / len(a)))  # :372


def kabsch_rmsd(coords1, coords2, mask) -> float:
    """Minimum RMSD over proper rotations (``scripts/validation_metrics.py:57-85``)."""
    a, b = _centered(coords1, coords2, mask)
    if a is None:
        return 0.0
    H = a.T @ b
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:                                  # validation_metrics.py:75-77
        Vt = Vt.copy()
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    moved = a @ R.T
    return float(np.sqrt(((moved - b) ** 2).sum() / len(a)))


def kabsch_rmsd_closed_form(coords1, coords2, mask) -> float:
    """sqrt((|X|^2 + |Y|^2 - 2 (s1 + s2 + d s3)) / n), d = sign det(H)."""
    a, b = _centered(coords1, coords2, mask)
    if a is None:
        return 0.0
    H = a.T @ b
    S = np.linalg.svd(H, compute_uv=False)
    d = 1.0 if np.linalg.det(H) >= 0 else -1.0
    e = (a * a).sum() + (b * b).sum() - 2.0 * (S[0] + S[1] + d * S[2])
    return float(np.sqrt(max(e, 0.0) / len(a)))
