"""Host -> device batch staging for the training loop (the reference moves each batch with blocking
``.to(device)`` calls inside ``run_epoch``, ``models/training.py:78-88``).

:class:`DevicePrefetcher` wraps any iterable of batches (dicts of pinned host tensors of fixed shapes) and
yields device batches, copying batch ``i+1`` on a side CUDA stream into the other half of a fixed double
buffer while batch ``i`` is being consumed, so the H2D copy (206 MB per step at config 2) overlaps the
previous step's kernels instead of preceding every step, and no device memory is allocated per step.
"""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Iterate over ``batches`` (dicts of host tensors, ideally pinned), one step ahead on a copy stream.

    A yielded batch is valid until the next one is requested (its buffers are then reused two steps later).
    """

    def __init__(self, batches, device):
        self.batches = batches
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [None, None]
        self.free = [None, None]          # main-stream events: the consumer is done with buffer k

    def _stage(self, host, k):
        if self.bufs[k] is None or any(self.bufs[k][n].shape != v.shape or self.bufs[k][n].dtype != v.dtype
                                       for n, v in host.items()):
            self.bufs[k] = {n: torch.empty(v.shape, dtype=v.dtype, device=self.device) for n, v in host.items()}
        with torch.cuda.stream(self.stream):
            if self.free[k] is not None:
                self.stream.wait_event(self.free[k])
            for n, v in host.items():
                self.bufs[k][n].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return self.bufs[k], ev

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it), 0)
        except StopIteration:
            return
        k = 0
        while nxt is not None:
            dev, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            try:
                nxt = self._stage(next(it), k ^ 1)      # next batch's copy overlaps this batch's compute
            except StopIteration:
                nxt = None
            yield dev
            done = torch.cuda.Event()
            done.record(cur)                            # kernels that read buffer k were enqueued before this point
            self.free[k] = done
            k ^= 1


# --------------------------------------------------------------------------------------------- ragged packed batches
FIELDS = ("n", "ca", "c", "mask", "dih", "labels", "emb")


def collate_packed(batch, pin: bool = True) -> dict:
    """Collate conformers WITHOUT padding: ``batch`` is a list of the reference dataset's 7-tuples
    ``(n[L,3], ca, c, mask[L], seq_emb[L,D]|None, dih[L,6], seq_labels[L])`` (``models/data.py:153-194``; centring not
    required), the result one contiguous (pinned) host tensor per field holding only the real rows, plus
    ``cu_seqlens [B+1]`` (int32) and ``lmax``.  The reference's ``_collate_single_batch`` (``:219-266``) ships
    ``B x Lmax`` rows instead; at configs[2] (lengths 64..512) that is 1.8x the bytes."""
    lens = [int(b[0].shape[0]) for b in batch]
    cu = torch.zeros(len(batch) + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0).to(torch.int32)
    T = int(cu[-1])
    D = next((int(b[4].shape[-1]) for b in batch if b[4] is not None), 0)

    def buf(shape, dtype=torch.float32):
        t = torch.zeros(shape, dtype=dtype)
        return t.pin_memory() if (pin and torch.cuda.is_available()) else t
    out = {"n": buf((T, 3)), "ca": buf((T, 3)), "c": buf((T, 3)), "mask": buf((T,)), "dih": buf((T, 6)),
           "labels": buf((T,), torch.int64), "emb": buf((T, D)) if D else None, "cu_seqlens": cu, "lmax": max(lens, default=0)}
    for i, (n, ca, c, m, emb, dih, lbl) in enumerate(batch):
        s = slice(int(cu[i]), int(cu[i + 1]))
        out["n"][s], out["ca"][s], out["c"][s], out["mask"][s], out["dih"][s], out["labels"][s] = n, ca, c, m, dih, lbl
        if D and emb is not None:
            out["emb"][s] = emb
    return out


def unpack_batch(packed: dict, device=None, center: bool = True):
    """Packed host (or device) batch -> the reference's padded 7-tuple ``(n, ca, c, mask, seq_emb, dih, seq_labels)`` on
    the device, centred on each conformer's valid-CA centroid (``models/data.py:166-172``) and zero-padded to ``lmax``
    (``:238-262``) by ONE kernel (``pev_unpack_center``); only the real rows cross PCIe."""
    from . import _lib
    from ._lib import ptr, stream
    dev = torch.device(device) if device is not None else packed["n"].device
    d = {k: (packed[k].to(dev, non_blocking=True) if packed.get(k) is not None else None) for k in FIELDS + ("cu_seqlens",)}
    B, Lmax = d["cu_seqlens"].numel() - 1, int(packed["lmax"])
    D = 0 if d["emb"] is None else d["emb"].shape[1]
    with torch.cuda.device_of(d["n"]):
        o = {"n": torch.empty(B, Lmax, 3, device=dev), "ca": torch.empty(B, Lmax, 3, device=dev),
             "c": torch.empty(B, Lmax, 3, device=dev), "mask": torch.empty(B, Lmax, device=dev),
             "dih": torch.empty(B, Lmax, 6, device=dev), "labels": torch.empty(B, Lmax, dtype=torch.int64, device=dev),
             "emb": torch.empty(B, Lmax, D, device=dev) if D else None}
        _lib.lib().call("pev_unpack_center", ptr(d["n"]), ptr(d["ca"]), ptr(d["c"]), ptr(d["mask"]), ptr(d["dih"]),
                        ptr(d["labels"]), ptr(d["emb"]), ptr(d["cu_seqlens"]), B, Lmax, D, int(center), ptr(o["n"]),
                        ptr(o["ca"]), ptr(o["c"]), ptr(o["mask"]), ptr(o["dih"]), ptr(o["labels"]), ptr(o["emb"]),
                        stream(d["n"]))
    return o["n"], o["ca"], o["c"], o["mask"], o["emb"], o["dih"], o["labels"]
