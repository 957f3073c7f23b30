"""Evaluation metrics on the device with the reference's function names
(``scripts/validation_metrics.py``): TM-score (``:23-54``), lDDT (``:92-149``), GDT-TS / GDT-HA (``:156-199``) and RMSF
(``:206-241``).  The reference evaluates one numpy structure pair per call; here every function also accepts a leading
batch dimension ``[S,L,3]`` (``coords_true`` shared ``[L,3]`` or per structure) and stays on the GPU.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32c, ptr, stream


def _batch(pred, true, mask):
    single = pred.dim() == 2
    a = f32c(pred.unsqueeze(0) if single else pred)
    b = f32c(true)
    if a.dim() != 3 or a.shape[-1] != 3 or b.shape[-2:] != a.shape[-2:]:
        raise ValueError("coords must be [L,3] or [S,L,3] with matching L")
    m = None if mask is None else f32c(mask.float() if mask.dtype == torch.bool else mask)
    return single, a, b, int(b.dim() == 3), m, int(m is not None and m.dim() == 2)


def superpose(coords_pred, coords_true, mask=None):
    """``kabsch_align`` (``:57-85``) + scores: dict with ``aligned [S,L,3]``, ``dist [S,L]``, ``tm``, ``gdt_ts``,
    ``gdt_ha`` ``[S]`` (the mask enters GDT only, as in the reference)."""
    single, a, b, b_batch, m, m_batch = _batch(coords_pred, coords_true, mask)
    S, L, _ = a.shape
    with torch.cuda.device_of(a):
        dev = a.device
        out = {"aligned": torch.empty(S, L, 3, device=dev), "dist": torch.empty(S, L, device=dev),
               "tm": torch.empty(S, device=dev), "gdt_ts": torch.empty(S, device=dev), "gdt_ha": torch.empty(S, device=dev)}
        _lib.lib().call("pev_superpose_scores", ptr(a), ptr(b), ptr(m), S, L, b_batch, m_batch, ptr(out["aligned"]),
                        ptr(out["dist"]), ptr(out["tm"]), ptr(out["gdt_ts"]), ptr(out["gdt_ha"]), stream(a))
    return {k: (v[0] if single else v) for k, v in out.items()}


def compute_tm_score(coords_pred, coords_true):
    """``compute_tm_score_python`` (``:23-54``): 0-d tensor for one pair, ``[S]`` for a batch."""
    return superpose(coords_pred, coords_true)["tm"]


def compute_gdt(coords_pred, coords_true, mask=None):
    """``compute_gdt`` (``:156-199``): ``(gdt_ts, gdt_ha)`` in percent."""
    r = superpose(coords_pred, coords_true, mask)
    return r["gdt_ts"], r["gdt_ha"]


def compute_lddt(coords_pred, coords_true, mask=None, cutoff: float = 15.0):
    """``compute_lddt`` (``:92-149``): ``(lddt_global, lddt_per_residue)``."""
    single, a, b, b_batch, m, m_batch = _batch(coords_pred, coords_true, mask)
    S, L, _ = a.shape
    with torch.cuda.device_of(a):
        per = torch.empty(S, L, device=a.device)
        glob = torch.empty(S, device=a.device)
        _lib.lib().call("pev_lddt", ptr(a), ptr(b), ptr(m), S, L, b_batch, m_batch, float(cutoff), ptr(per), ptr(glob),
                        stream(a))
    return (glob[0], per[0]) if single else (glob, per)


def compute_rmsf(ensemble_coords, mask=None):
    """``compute_rmsf`` (``:206-241``): per-residue fluctuation ``[L]`` of ``ensemble_coords [N,L,3]`` after aligning every
    member to the first (``mask`` is accepted and unused, as in the reference)."""
    a = f32c(ensemble_coords)
    N, L, _ = a.shape
    if N == 1:
        return torch.zeros(L, device=a.device)
    aligned = superpose(a, a[0])["aligned"]
    with torch.cuda.device_of(a):
        out = torch.empty(L, device=a.device)
        _lib.lib().call("pev_rmsf", ptr(aligned), N, L, ptr(out), stream(a))
    return out
