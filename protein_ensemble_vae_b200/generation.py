"""Batched ensemble generation on the device: what ``generate_ensembles``
(``generate_ensemble_pdbs.py:376-672``) does one sample and one Python loop at a time.

* :func:`validate_protein_geometry` -- the reference's signature and messages (``:290-340``);
  :func:`validate_geometry_batch` -- the same filter for ``S`` conformers in one launch;
* :func:`generate_ensemble` -- decode ``S`` latent samples in chunks (sharded across ranks when ``torch.distributed``
  is initialised), filter them, score them against a reference structure (Kabsch RMSD, ``:500``, ``:594``) and
  measure the ensemble's diversity (mean pairwise RMSD, ``:591-597``), all without leaving the device.

The PDB writer (``:148-288``) is host-side text formatting and stays with the caller.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32c, ptr, stream

STATUS = ("Valid geometry", "No valid residues", "Extreme CA-CA distance", "Abnormal average CA-CA distance",
          "Abnormal average CA-CA-CA angle")


def validate_geometry_batch(coords_ca, mask=None):
    """``coords_ca[S,L,3]``, ``mask[L] | [S,L] | None`` -> ``(status int32 [S], stats float32 [S,3])``; status 0 = valid,
    1..4 index :data:`STATUS`; stats = (max CA-CA distance, mean CA-CA distance, mean CA-CA-CA angle in degrees)."""
    a = f32c(coords_ca)
    if a.dim() != 3 or a.shape[-1] != 3:
        raise ValueError("coords_ca must be [S,L,3]")
    S, L, _ = a.shape
    m, m_batch = None, 0
    if mask is not None:
        m = f32c(mask)
        m_batch = int(m.dim() == 2)
        if m.shape[-1] != L or (m_batch and m.shape[0] != S):
            raise ValueError("mask must be [L] or [S,L]")
    with torch.cuda.device_of(a):
        status = torch.empty(S, dtype=torch.int32, device=a.device)
        stats = torch.empty(S, 3, dtype=torch.float32, device=a.device)
        _lib.lib().call("pev_validate_geometry", ptr(a), ptr(m), S, L, m_batch, ptr(status), ptr(stats), stream(a))
    return status, stats


def validate_protein_geometry(coords_ca, mask):
    """``(is_valid, reason)`` for one structure, as ``generate_ensemble_pdbs.py:290-340``."""
    status, stats = validate_geometry_batch(coords_ca.unsqueeze(0), mask)
    code = int(status[0])
    mx, avg, ang = (float(v) for v in stats[0])
    if code == 2:
        return False, f"Extreme CA-CA distance {mx:.3f}Å"
    if code == 3:
        return False, f"Abnormal average CA-CA distance {avg:.3f}Å"
    if code == 4:
        return False, f"Abnormal average CA-CA-CA angle {ang:.1f}°"
    return code == 0, STATUS[code]


@torch.no_grad()
def generate_ensemble(decoder, z_g, z_l, mask=None, reference_ca=None, chunk: int = 2048, ref_compat: bool = False,
                      diversity_samples: int = 256):
    """Decode the latent samples ``z_g[S,zg], z_l[S,L,zl]`` (this rank's contiguous share when distributed) and return
    a dict of device tensors: ``N, CA, C`` ``[S_local,L,3]``, ``logits``, ``valid`` (bool), ``status``, ``stats``,
    ``rmsd`` (against ``reference_ca``, if given) and ``diversity`` (mean pairwise Kabsch RMSD over the first
    ``diversity_samples`` valid members)."""
    from .distributed import shard_range, world
    from .kabsch import ensemble_diversity, kabsch_rmsd_batch
    rank, ws = world()
    lo, hi = shard_range(z_l.shape[0], rank, ws)
    outs = ([], [], [], [])
    for a in range(lo, hi, chunk):
        b = min(hi, a + chunk)
        m = None if mask is None else (mask[a:b] if mask.dim() == 2 else mask.unsqueeze(0).expand(b - a, -1))
        for acc, o in zip(outs, decoder(z_g[a:b], z_l[a:b], m)):
            acc.append(o)
    n, ca, c, lg = (torch.cat(o) if o else None for o in outs)
    res = {"N": n, "CA": ca, "C": c, "logits": lg}
    if ca is None:
        return res
    km = None if mask is None else (mask[lo:hi] if mask.dim() == 2 else mask)
    res["status"], res["stats"] = validate_geometry_batch(ca, km)
    res["valid"] = res["status"] == 0
    if reference_ca is not None:
        res["rmsd"] = kabsch_rmsd_batch(ca, reference_ca, km, ref_compat=ref_compat)
    keep = torch.nonzero(res["valid"]).squeeze(-1)[:diversity_samples]
    dmask = None if km is None else (km if km.dim() == 1 else km[0])
    res["diversity"] = ensemble_diversity(ca.index_select(0, keep), dmask, ref_compat=ref_compat)
    return res
