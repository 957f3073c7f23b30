"""Batched ensemble generation on the device: what ``generate_ensembles``
(``generate_ensemble_pdbs.py:376-672``) does one sample and one Python loop at a time.

* :func:`validate_protein_geometry` -- the reference's signature and messages (``:290-340``);
  :func:`validate_geometry_batch` -- the same filter for ``S`` conformers in one launch;
* :func:`generate_ensemble` -- decode ``S`` latent samples in chunks (sharded across ranks when ``torch.distributed``
  is initialised), filter them, score them against a reference structure (Kabsch RMSD, ``:500``, ``:594``) and
  measure the ensemble's diversity (mean pairwise RMSD, ``:591-597``), all without leaving the device.

* :func:`write_ensemble_pdb` -- the multi-model PDB text of ``write_pdb`` (``:148-288``) formatted on the device and
  streamed to the file through pinned buffers, byte-identical to the reference's output.
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
                      diversity_samples: int = 256, pdb_path=None, sequence=None, pdb_id=None, chain_id: str = "A"):
    """Decode the latent samples ``z_g[S,zg], z_l[S,L,zl]`` (this rank's contiguous share when distributed) and return
    a dict of device tensors: ``N, CA, C`` ``[S_local,L,3]``, ``logits``, ``valid`` (bool), ``status``, ``stats``,
    ``rmsd`` (against ``reference_ca``, if given) and ``diversity`` (mean pairwise Kabsch RMSD over the first
    ``diversity_samples`` valid members).  ``pdb_path``: also write the geometrically valid members (one shared ``mask``) as
    a multi-model PDB file (:func:`write_ensemble_pdb`; the reference writes model by model, ``:598-625``)."""
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
    if pdb_path is not None:
        ok = torch.nonzero(res["valid"]).squeeze(-1)
        wm = dmask if dmask is not None else torch.ones(ca.shape[1], device=ca.device)
        res["pdb_bytes"] = write_ensemble_pdb(pdb_path, n.index_select(0, ok), ca.index_select(0, ok), c.index_select(0, ok), wm,
                                              sequence=sequence, pdb_id=pdb_id, chain_id=chain_id)
    return res


# ------------------------------------------------------------------------------------------------ streaming PDB writer
_AA3 = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY", "H": "HIS", "I": "ILE",
        "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER", "T": "THR", "W": "TRP", "Y": "TYR", "V": "VAL"}


def pdb_header(num_models: int, pdb_id=None, chain_id: str = "A", title=None) -> str:
    """The header ``write_pdb`` emits before the first model (``generate_ensemble_pdbs.py:177-216``)."""
    h = [f"HEADER    PROTEIN STRUCTURE                    {pdb_id.upper():>4}              \n" if pdb_id else
         "HEADER    PROTEIN STRUCTURE                    UNKN              \n",
         f"TITLE     {title[:70]:<70}\n" if title else "TITLE     GENERATED PROTEIN STRUCTURE BY ENHANCED VAE MODEL\n",
         "COMPND    MOL_ID: 1;\n", "COMPND   2 MOLECULE: GENERATED PROTEIN STRUCTURE;\n", f"COMPND   3 CHAIN: {chain_id};\n",
         "COMPND   4 SYNONYM: VAE-GENERATED STRUCTURE;\n", "COMPND   5 ENGINEERED: YES\n", "SOURCE    MOL_ID: 1;\n",
         "SOURCE   2 ORGANISM_SCIENTIFIC: SYNTHETIC;\n", "SOURCE   3 ORGANISM_COMMON: COMPUTER-GENERATED;\n",
         "SOURCE   4 ORGANISM_TAXID: 0;\n", "SOURCE   5 GENE: VAE-GENERATED;\n",
         "SOURCE   6 EXPRESSION_SYSTEM: COMPUTATIONAL MODEL;\n", "SOURCE   7 EXPRESSION_SYSTEM_TAXID: 0;\n",
         "KEYWDS    VAE, GENERATED, PROTEIN, STRUCTURE\n", "EXPDTA    COMPUTATIONAL MODELING\n",
         f"NUMMDL    {num_models or 1:4d}\n", "AUTHOR    ENHANCED PROTEIN VAE MODEL\n", "REVDAT   1   01-JAN-24 UNKN    0\n",
         "REMARK   2\n", "REMARK   2 RESOLUTION. NOT APPLICABLE.\n", "REMARK   3\n", "REMARK   3 REFINEMENT.\n",
         "REMARK   3   PROGRAM     : ENHANCED PROTEIN VAE\n", "REMARK   3   AUTHORS     : COMPUTATIONAL MODEL\n", "REMARK   4\n",
         "REMARK   4 GENERATED STRUCTURE COMPLIES WITH FORMAT V. 3.30\n", "REMARK 100\n",
         "REMARK 100 THIS ENTRY WAS GENERATED BY ENHANCED PROTEIN VAE\n",
         "REMARK 100 COMPLETE BACKBONE WITH PROPER CONNECTIVITY\n", "\n"]
    return "".join(h)


def write_ensemble_pdb(path, coords_n, coords_ca, coords_c, mask, sequence=None, pdb_id=None, chain_id: str = "A", title=None,
                       chunk: int = 4096) -> int:
    """Write a whole ensemble ``[S,L,3]`` (device tensors, one shared ``mask [L]``) as a multi-model PDB file, byte for byte
    what the reference produces by calling ``write_pdb(..., model_num=m, num_models=S)`` for ``m = 1..S``
    (``generate_ensemble_pdbs.py:148-288``, ``:598-625``): complete backbone with the O atom placed as in
    ``compute_backbone_oxygen``, CONECT records, TER / ENDMDL.  The text is formatted on the device, ``chunk`` models at a
    time, and streamed to the file through two pinned host buffers (the copy of chunk k+1 runs while chunk k is written).
    Returns the number of bytes written."""
    from . import _lib
    from ._lib import f32c, ptr, stream
    n, ca, c = f32c(coords_n), f32c(coords_ca), f32c(coords_c)
    if n.dim() == 2:
        n, ca, c = n[None], ca[None], c[None]
    S, L, _ = ca.shape
    dev = ca.device
    m = (mask.detach().float().cpu() > 0.5).tolist()
    valid = [i for i in range(L) if m[i]]
    nv = len(valid)
    prev_ok = [1 if (i > 0 and m[i - 1]) else 0 for i in valid]
    names = "".join(_AA3.get(sequence[i], "ALA") if (sequence and i < len(sequence)) else "ALA" for i in valid)
    lib = _lib.lib()
    with torch.cuda.device_of(ca):
        vidx = torch.tensor(valid, dtype=torch.int32, device=dev)
        pok = torch.tensor(prev_ok, dtype=torch.uint8, device=dev)
        rname = torch.tensor(list(names.encode("ascii")), dtype=torch.uint8, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        total = 0
        pinned = [None, None]
        pending = None                                    # (host buffer, event) of the chunk in flight

        def flush(f):
            nonlocal pending, total
            if pending is not None:
                buf, ev = pending
                ev.synchronize()
                f.write(memoryview(buf.numpy()))
                total += buf.numel()
                pending = None

        with open(path, "wb") as f:
            head = pdb_header(S, pdb_id, chain_id, title).encode("ascii")
            f.write(head)
            total += len(head)
            for k, s0 in enumerate(range(0, S, chunk)):
                s1 = min(S, s0 + chunk)
                nbytes = int(lib.cdll.pev_pdb_models_bytes(s0 + 1, s1 - s0, nv))
                out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                lib.call("pev_pdb_format_models", ptr(n[s0:s1]), ptr(ca[s0:s1]), ptr(c[s0:s1]), ptr(vidx), ptr(pok), ptr(rname),
                         s1 - s0, L, nv, s0 + 1, ord(chain_id[0]), ptr(out), ptr(flag), stream(ca))
                if dev.type == "cuda":
                    hb = pinned[k & 1]
                    if hb is None or hb.numel() < nbytes:
                        hb = pinned[k & 1] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                    view = hb[:nbytes]
                    view.copy_(out, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    flush(f)                               # write chunk k-1 while chunk k is formatted / copied
                    pending = (view, ev)
                else:                                      # host-compiled kernel bodies (tests)
                    f.write(memoryview(out.numpy()))
                    total += nbytes
            flush(f)
        if int(flag.item()):
            raise ValueError("a coordinate does not fit the PDB %8.3f field (|x| >= 9999.9995)")
    return total
