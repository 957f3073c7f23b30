"""Data parallelism over conformers: one process per GPU, ``torch.distributed`` (NCCL).

Conformers never interact, so the path shards without any data-path collective:

* **train** -- each rank holds its own slice of the conformer batch and a replica of the parameters;
  the only exchange is one gradient all-reduce per step (decoder: 4.45 M fp32 = 17.8 MB), issued on a
  flat bucket so it is a single NCCL call over NVLink/NVSwitch.
* **decode** -- the ``S`` latent samples are split into contiguous per-rank ranges; every rank decodes
  and scores (Kabsch RMSD) its range locally and a single final gather collects the ``[S]`` RMSDs
  (and optionally the coordinates).

Loss normalisation: every term of ``compute_total_loss`` is a mean over the rank's conformers /
valid residues (``models/losses.py:12-21``, ``:54-57``), so averaging gradients over ranks reproduces the
global-batch gradient exactly only when all ranks hold the same denominators.  For ragged shards (BASELINE
config 3: mixed lengths) ``compute_total_loss(..., dp_normalize=True)`` all-reduces the 17 term denominators and
rescales each rank's terms by its share (:func:`shard_term_scale`); :func:`balanced_shards` assigns conformers to
ranks by edge count so that every rank gets the same amount of edge work; :class:`GradBuckets` all-reduces the
gradients bucket by bucket while backward is still running.
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


def shard_term_scale(den_inv: torch.Tensor):
    """``world * den_local / sum_ranks(den_local)`` per loss term from the reciprocal local denominators
    (``None`` when not distributed); 0 where the local denominator is 0 (an empty shard)."""
    rank, ws = world()
    if ws == 1:
        return None
    den = torch.where(torch.isfinite(den_inv) & (den_inv > 0), 1.0 / den_inv, torch.zeros_like(den_inv))
    tot = den.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return torch.where(tot > 0, ws * den / tot.clamp_min(1e-30), torch.zeros_like(den))


def band_edges(L: int, W: int) -> int:
    """Edges of the banded residue graph of ``L`` valid residues (``models/en_gnn_decoder.py:180-189``)."""
    if L < 2:
        return 0
    return L * (L - 1) if W >= L - 1 else 2 * W * L - W * (W + 1)


def balanced_shards(lengths, world_size: int, max_neighbors: int = 40):
    """Assign conformers (by index) to ranks so that every rank gets nearly the same number of EDGES -- the unit of
    work of the EGNN layers -- instead of the same number of conformers: longest-processing-time greedy on
    ``band_edges(L)``.  Returns ``world_size`` index lists (each sorted); deterministic."""
    order = sorted(range(len(lengths)), key=lambda i: (-band_edges(int(lengths[i]), max_neighbors), i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += band_edges(int(lengths[i]), max_neighbors)
    return [sorted(o) for o in out]


class GradBuckets:
    """Bucketed gradient all-reduce overlapped with backward.

    ``buckets``: lists of parameters (e.g. one list per EGNN layer, in the order backward reaches them).  Every
    parameter's ``.grad`` becomes a view into its bucket's flat fp32 buffer, so autograd accumulates straight into
    the buffer; when the last gradient of a bucket has been accumulated (post-accumulate hook) the bucket's
    all-reduce is launched asynchronously and runs on the communication stream under the rest of backward.
    :meth:`finish` waits for the outstanding collectives (and averages); :meth:`zero` clears the buffers for the next
    step -- use it instead of ``optimizer.zero_grad(set_to_none=True)``, which would drop the views.
    """

    def __init__(self, buckets, average: bool = True):
        self.average = average
        self.buckets = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.flat, self.pending, self.works, self.handles = [], [], [], []
        self.launched = [False] * len(self.buckets)
        for bi, b in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device)
            off = 0
            for p in b:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self.handles.append(p.register_post_accumulate_grad_hook(self._hook(bi)))
            self.flat.append(flat)
            self.pending.append(len(b))

    def _hook(self, bi):
        def fn(_p):
            self.pending[bi] -= 1
            if self.pending[bi] == 0:
                self._launch(bi)
        return fn

    def _launch(self, bi):
        rank, ws = world()
        self.launched[bi] = True
        if ws > 1:
            self.works.append(dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        """Call after ``backward()``: reduces the buckets whose hooks did not all fire (parameters without a gradient
        this step), waits for every collective and averages."""
        rank, ws = world()
        for bi in range(len(self.buckets)):
            if not self.launched[bi]:
                self._launch(bi)
        for w in self.works:
            w.wait()
        self.works = []
        if self.average and ws > 1:
            for f in self.flat:
                f.div_(ws)
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

    def zero(self):
        for f in self.flat:
            f.zero_()

    def close(self):
        for h in self.handles:
            h.remove()
        self.handles = []


def decoder_buckets(decoder):
    """Gradient buckets of an ``EGNNDecoder`` in backward order: heads first, then the layers last to first, then the
    input projections."""
    heads = list(decoder.sequence_head.parameters()) + list(decoder.n_offset_head.parameters()) \
        + list(decoder.c_offset_head.parameters())
    out = [heads] + [list(l.parameters()) for l in reversed(decoder.layers)]
    out.append(list(decoder.input_embedding.parameters()) + list(decoder.latent_to_coords.parameters()))
    return out


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
