"""Data parallelism over conformers: one process per GPU, ``torch.distributed`` (NCCL).

Conformers never interact, so the path shards without any data-path collective:

* **train** -- each rank holds its own slice of the conformer batch and a replica of the parameters;
  the only exchange is one gradient all-reduce per step (decoder: 4.45 M fp32 = 17.8 MB), issued on a
  flat bucket so it is a single NCCL call over NVLink/NVSwitch.
* **decode** -- the ``S`` latent samples are split into contiguous per-rank ranges; every rank decodes
  and scores (Kabsch RMSD) its range locally and a single final gather collects the ``[S]`` RMSDs
  (and optionally the coordinates).

Loss normalisation: every term of ``compute_total_loss`` is a mean over the rank's conformers /
valid residues, so averaging gradients over ranks reproduces the global-batch gradient exactly when
all ranks hold the same number of conformers and of valid residues (the benchmark's case); with
ragged shards scale each rank's loss by its share of the denominator before ``backward()``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """``(rank, world_size)``; ``(0, 1)`` when ``torch.distributed`` is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous ``[lo, hi)`` slice of ``n`` units owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_gradients(params, average: bool = True) -> None:
    """All-reduce the ``.grad`` of ``params`` in place through one flat bucket (one collective per step)."""
    rank, ws = world()
    if ws == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= ws
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


@torch.no_grad()
def decode_ensemble(decoder, z_g, z_l, mask=None, reference_ca=None, chunk: int = 2048, ref_compat: bool = False,
                    gather: bool = True, return_coords: bool = False):
    """Decode this rank's contiguous share of ``S`` latent samples and score them against ``reference_ca``.

    ``z_g[S,zg], z_l[S,L,zl], mask[S,L]|[L]|None``; all ranks pass the same full tensors (or tensors whose
    leading dimension is the global ``S``) and only their own slice is touched.  Returns ``rmsd[S]`` on every
    rank when ``gather`` (one ``all_gather``), else the local slice; with ``return_coords`` also the local CA
    coordinates ``[S_local, L, 3]``.
    """
    from .kabsch import kabsch_rmsd_batch
    rank, ws = world()
    S = z_l.shape[0]
    lo, hi = shard_range(S, rank, ws)
    rm, cas = [], []
    for a in range(lo, hi, chunk):
        b = min(hi, a + chunk)
        m = None if mask is None else (mask[a:b] if mask.dim() == 2 else mask.unsqueeze(0).expand(b - a, -1))
        n, ca, c, _ = decoder(z_g[a:b], z_l[a:b], m)
        if reference_ca is not None:
            km = None if mask is None else (mask[a:b] if mask.dim() == 2 else mask)
            rm.append(kabsch_rmsd_batch(ca, reference_ca, km, ref_compat=ref_compat))
        if return_coords:
            cas.append(ca)
    local = torch.cat(rm) if rm else torch.zeros(0, device=z_l.device)
    out = local
    if gather and ws > 1 and reference_ca is not None:
        sizes = [shard_range(S, r, ws) for r in range(ws)]
        width = max(h - l for l, h in sizes)
        pad = torch.zeros(width, device=local.device, dtype=local.dtype)
        pad[:local.numel()] = local
        parts = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(parts, pad)
        out = torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)])
    if return_coords:
        return out, (torch.cat(cas) if cas else None)
    return out
