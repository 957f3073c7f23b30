"""``run_epoch`` -- the caller of the hot path in ``models/training.py:22-199`` with the host synchronisations taken out.

Same positional signature and the same 12-key result (``loss, rec, pair, klg, kll, dihedral, rama, bond, angle, seq,
seq_acc, clash``: batch-size-weighted epoch means).  What the reference does per batch and this does not:

* 13 ``.item()`` read-backs (``:161-172`` + the accuracy) -- every one a device synchronisation that drains the stream the
  kernels of this package are queued on.  Here the 12 statistics are accumulated in ONE device vector and read back once
  per epoch;
* blocking ``.to(device)`` copies (``:65-86``) -- batches are staged by :class:`~.data.DevicePrefetcher` on a copy stream
  under the previous step.  A loader may yield the reference's ``(input_7tuple, target_7tuple)`` pairs (``collate_pad``)
  or pairs of ragged packed batches (``data.collate_packed``), which are padded and centred on the device;
* the per-batch ``torch.isfinite(loss)`` branch (``:134-145``) is a synchronisation too: the check is made on the
  accumulated statistics at the end of the epoch, or every ``check_finite_every`` batches when asked, and raises the
  reference's ``ValueError`` either way.

Gradient clipping (``:148``, max norm 10) stays, through ``torch.nn.utils.clip_grad_norm_`` with ``foreach`` kernels and no
read-back; ``wandb`` logging is replaced by an optional ``log_fn(batch_index, {"grad_norm": tensor, "loss": tensor})`` that
receives device tensors.  Data-parallel runs pass ``grad_sync`` (e.g. ``GradBuckets.finish`` or
``lambda: allreduce_gradients(params)``) and ``dp_normalize=True``.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from .data import DevicePrefetcher, unpack_batch
from .losses import compute_total_loss

STAT_KEYS = ("loss", "rec", "pair", "klg", "kll", "dihedral", "rama", "bond", "angle", "seq", "seq_acc", "clash")
_LOSS_OF_STAT = {"loss": "total", "rec": "reconstruction", "pair": "pair_distance", "klg": "kl_global", "kll": "kl_local",
                 "dihedral": "dihedral_total", "rama": "ramachandran", "bond": "bond_length", "angle": "bond_angle",
                 "seq": "sequence", "clash": "clash"}
_FIELDS = ("n", "ca", "c", "mask", "emb", "dih", "labels")


def _as_host_dict(batch_data) -> dict:
    """One loader item -> flat dict of host tensors (``in.*`` / ``tgt.*``) for the prefetcher; packed batches pass through
    with their ``cu_seqlens`` / ``lmax``."""
    input_data, target_data = batch_data
    out = {}
    for side, d in (("in", input_data), ("tgt", target_data)):
        if isinstance(d, dict):                                  # data.collate_packed
            for k, v in d.items():
                if v is not None:
                    out[f"{side}.{k}"] = v
        else:                                                    # the reference's padded 7-tuple (models/data.py:219-266)
            for k, v in zip(_FIELDS, d):
                if v is not None:
                    out[f"{side}.{k}"] = v
    return out


def _side(d: dict, side: str, device):
    sub = {k[len(side) + 1:]: v for k, v in d.items() if k.startswith(side + ".")}
    if "cu_seqlens" in sub:
        sub.setdefault("emb", None)
        return unpack_batch(sub, device)
    return tuple(sub.get(k) for k in _FIELDS)


class _HostScalars:
    """Non-tensor entries (``lmax`` of a packed batch) ride along the prefetcher untouched."""

    @staticmethod
    def split(d: dict):
        tensors = {k: v for k, v in d.items() if torch.is_tensor(v)}
        return tensors, {k: v for k, v in d.items() if not torch.is_tensor(v)}


def run_epoch(model, loader: Iterable, opt, device, klw_g, klw_l, w_pair, pair_stride, train, w_dihedral, w_rama, w_bond,
              w_angle, w_rec, w_seq, w_clash, epoch, *, max_grad_norm: float = 10.0, grad_sync: Optional[Callable] = None,
              dp_normalize: bool = False, check_finite_every: int = 0, log_fn: Optional[Callable] = None) -> dict:
    """``models/training.py:22-199``.  Returns the epoch means as python floats (one device read-back)."""
    device = torch.device(device)
    model.train(bool(train))
    stats = torch.zeros(len(STAT_KEYS) + 1, dtype=torch.float64, device=device)        # 12 weighted sums + the sample count
    params = [p for p in model.parameters() if p.requires_grad]

    extras = []

    def host_batches():
        for item in loader:
            tensors, scalars = _HostScalars.split(_as_host_dict(item))
            extras.append(scalars)
            yield tensors

    for batch_idx, dev_batch in enumerate(DevicePrefetcher(host_batches(), device)):
        d = dict(dev_batch)
        d.update(extras[batch_idx])
        n_in, ca_in, c_in, mask_in, seqemb_in, dih_in, _ = _side(d, "in", device)
        n_tgt, ca_tgt, c_tgt, mask_tgt, _, dih_tgt, seq_lbl_tgt = _side(d, "tgt", device)
        mask = mask_tgt                                                                 # :88
        with torch.set_grad_enabled(bool(train)):
            pred_N, pred_CA, pred_C, pred_seq, mu_g, lv_g, mu_l, lv_l = model(seqemb_in, n_in, ca_in, c_in, dih_in, mask)
            loss_dict = compute_total_loss(
                pred_N=pred_N, pred_CA=pred_CA, pred_C=pred_C, pred_seq=pred_seq, target_N=n_tgt, target_CA=ca_tgt,
                target_C=c_tgt, target_seq_labels=seq_lbl_tgt, mask=mask, mu_g=mu_g, lv_g=lv_g, mu_l=mu_l, lv_l=lv_l,
                target_dihedrals=dih_tgt, klw_g=klw_g, klw_l=klw_l, w_pair=w_pair, pair_stride=pair_stride,
                w_dihedral=w_dihedral, w_rama=w_rama, w_bond=w_bond, w_angle=w_angle, w_rec=w_rec, w_seq=w_seq,
                w_clash=w_clash, **({"dp_normalize": True} if dp_normalize else {}))
            loss = loss_dict["total"]
            with torch.no_grad():                                                       # :107-110
                mb = mask.bool()
                correct = (torch.argmax(pred_seq, dim=-1) == seq_lbl_tgt) & mb
                seq_accuracy = correct.sum().float() / mb.sum().float()
            if train:
                opt.zero_grad()
                loss.backward()
                if grad_sync is not None:
                    grad_sync()
                grad_norm = torch.nn.utils.clip_grad_norm_(params, max_norm=max_grad_norm)   # :148 (no read-back)
                opt.step()
                if log_fn is not None:
                    log_fn(batch_idx, {"grad_norm": grad_norm, "loss": loss.detach()})
        with torch.no_grad():                                                           # :160-174 without the 13 .item()
            bs = float(ca_tgt.size(0))
            vals = [loss_dict[_LOSS_OF_STAT[k]].detach() if k != "seq_acc" else seq_accuracy for k in STAT_KEYS]
            stats[:-1] += torch.stack([v.reshape(()).double() for v in vals]) * bs
            stats[-1] += bs
        if check_finite_every and (batch_idx + 1) % check_finite_every == 0 and not bool(torch.isfinite(stats[0])):
            raise ValueError(f"Training collapsed - NaN detected (epoch {epoch}, by batch {batch_idx})")
    host = stats.cpu()                                                                  # the epoch's one synchronisation
    n = float(host[-1])
    if n == 0.0:
        raise ValueError("run_epoch: the loader yielded no batch")
    if train and not bool(torch.isfinite(host[0])):
        raise ValueError(f"Training collapsed - NaN detected (epoch {epoch})")           # :134-140, deferred to the epoch's end
    return {k: float(host[i]) / n for i, k in enumerate(STAT_KEYS)}
