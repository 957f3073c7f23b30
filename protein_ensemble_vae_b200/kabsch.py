"""Kabsch RMSD (K4) with the reference's signature plus a batched device form.

``kabsch_rmsd(coords1[L,3], coords2[L,3], mask[L]) -> float`` mirrors
``generate_ensemble_pdbs.py:343-373``.  The reference applies the inverse rotation
(``c1 @ R`` with ``R = V D U^T``), so its value is not the superposition optimum;
``ref_compat=True`` reproduces that convention, the default returns the true minimum
(equal to ``scripts/validation_metrics.py:57-85`` + per-residue RMSD).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32c, ptr, stream


def kabsch_rmsd_batch(coords, ref, mask=None, ref_compat: bool = False) -> torch.Tensor:
    """``coords[S,L,3]`` against ``ref[L,3]`` (shared) or ``ref[S,L,3]``; ``mask[L]``, ``[S,L]`` or None -> ``[S]``.

    Stays on the device (no host sync); an 8-lane slot per conformer, the 3 x 3 solves of a warp's group lane-parallel.
    """
    a = f32c(coords)
    if a.dim() != 3 or a.shape[-1] != 3:
        raise ValueError("coords must be [S,L,3]")
    S, L, _ = a.shape
    b = f32c(ref)
    b_batch = int(b.dim() == 3)
    if b.shape[-2:] != (L, 3) or (b_batch and b.shape[0] != S):
        raise ValueError("ref must be [L,3] or [S,L,3]")
    m, m_batch = None, 0
    if mask is not None:
        m = f32c(mask)
        m_batch = int(m.dim() == 2)
        if m.shape[-1] != L or (m_batch and m.shape[0] != S):
            raise ValueError("mask must be [L] or [S,L]")
    with torch.cuda.device_of(a):
        out = torch.empty(S, dtype=torch.float32, device=a.device)
        _lib.lib().call("pev_kabsch_rmsd", ptr(a), ptr(b), ptr(m), S, L, b_batch, m_batch, int(ref_compat),
                        ptr(out), stream(a))
    return out


def kabsch_rmsd_pairs(coords, mask=None, ref_compat: bool = False) -> torch.Tensor:
    """All-pairs RMSD matrix ``[S,S]`` of one ensemble ``coords[S,L,3]`` (``mask[L]`` or None): symmetric, zero
    diagonal, entry ``(i,j)``, ``i < j``, = ``kabsch_rmsd(coords[i], coords[j], mask)``.  One launch, an 8-lane slot per
    pair, instead of the ``S(S-1)/2`` host calls of ``generate_ensemble_pdbs.py:591-595``."""
    a = f32c(coords)
    if a.dim() != 3 or a.shape[-1] != 3:
        raise ValueError("coords must be [S,L,3]")
    S, L, _ = a.shape
    m = None
    if mask is not None:
        m = f32c(mask)
        if m.shape != (L,):
            raise ValueError("mask must be [L]")
    with torch.cuda.device_of(a):
        out = torch.empty(S, S, dtype=torch.float32, device=a.device)
        _lib.lib().call("pev_kabsch_rmsd_pairs", ptr(a), ptr(m), S, L, int(ref_compat), ptr(out), stream(a))
    return out


def ensemble_diversity(coords, mask=None, ref_compat: bool = False) -> torch.Tensor:
    """Mean pairwise Kabsch RMSD of an ensemble (0-d device tensor; 0 for fewer than two members):
    ``avg_diversity`` of ``generate_ensemble_pdbs.py:591-597``."""
    S = coords.shape[0]
    if S < 2:
        return torch.zeros((), device=coords.device)
    return kabsch_rmsd_pairs(coords, mask, ref_compat=ref_compat).sum() / (S * (S - 1))


def kabsch_rmsd(coords1, coords2, mask, ref_compat: bool = False) -> float:
    """RMSD after Kabsch alignment of one pair; python float, 0.0 for an empty mask (``:350-351``)."""
    return float(kabsch_rmsd_batch(coords1.unsqueeze(0), coords2, mask, ref_compat=ref_compat)[0])
