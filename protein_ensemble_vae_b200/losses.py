"""Geometric / KL / sequence losses with the reference's signatures (``models/losses.py``),
evaluated by the fused CUDA loss kernels (K3) through the C ABI.

Every public function of ``models/losses.py`` is here with the same positional order, the
same return convention (0-d tensors; ``compute_total_loss`` returns the same 16-key dict,
``models/losses.py:596-613``) and autograd support for the same inputs.  All base terms come
from one forward pass (``pev_loss_fwd`` + ``pev_loss_finalize``) and one backward pass
(``pev_loss_bwd``); the weighted sums on top are ordinary torch scalar ops.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import (NUM_TERMS, T_ANG_CACN, T_ANG_CNCA, T_ANG_NCAC, T_BOND_CAC, T_BOND_CN, T_BOND_NCA,
                   T_CLASH, T_DIH_CONS, T_KL_G, T_KL_L, T_OMEGA, T_PAIR, T_RAMA, T_REC_C, T_REC_CA,
                   T_REC_N, T_SEQ, LossArgs, f32c, ptr, stream)

_DIFF = ("pred_N", "pred_CA", "pred_C", "logits", "mu_l", "lv_l", "mu_g", "lv_g")
_CONST = ("target_N", "target_CA", "target_C", "mask", "target_dih", "labels")


def _void(t):
    p = ptr(t)
    return p.value if p is not None else None


def _make_args(t, cfg):
    a = LossArgs()
    for name in _DIFF + _CONST:
        setattr(a, name, _void(t.get(name)))
    mask = t["mask"]
    a.B, a.L = mask.shape
    a.C = t["logits"].shape[-1] if t.get("logits") is not None else 0
    a.D = t["mu_l"].shape[-1] if t.get("mu_l") is not None else 0
    a.G = t["mu_g"].shape[-1] if t.get("mu_g") is not None else 0
    a.pair_stride = int(cfg.get("pair_stride", 0))
    a.enable_clash = int(cfg.get("clash", False))
    a.enable_geometry = int(cfg.get("geometry", False))
    a.clash_dist = float(cfg.get("clash_dist", 3.2))
    a.soft_margin = float(cfg.get("soft_margin", 0.5))
    return a


class _LossTerms(torch.autograd.Function):
    """All 17 base terms (``enum pev_term``) in one pass; see ``include/pev_b200.h``."""

    @staticmethod
    def forward(ctx, cfg, consts, *diff):
        t = {k: f32c(v) for k, v in zip(_DIFF, diff)}
        for k in _CONST:
            v = consts.get(k)
            t[k] = (v.contiguous() if k == "labels" else f32c(v)) if v is not None else None
        if t["labels"] is not None and t["labels"].dtype != torch.int64:
            t["labels"] = t["labels"].long()
        mask = t["mask"]
        B = mask.shape[0]
        dev = mask.device
        with torch.cuda.device_of(mask):
            acc_g = torch.zeros(2 * NUM_TERMS, dtype=torch.float64, device=dev)
            acc_s = torch.zeros(B * 8, dtype=torch.float64, device=dev)
            terms = torch.empty(NUM_TERMS, dtype=torch.float32, device=dev)
            inv_den = torch.empty(NUM_TERMS + 2 * B, dtype=torch.float32, device=dev)
            args = _make_args(t, cfg)
            L = _lib.lib()
            L.call("pev_loss_fwd", ctypes.byref(args), ptr(acc_g), ptr(acc_s), stream(mask))
            L.call("pev_loss_finalize", ptr(acc_g), ptr(acc_s), B, ptr(terms), ptr(inv_den), stream(mask))
        ctx.cfg = cfg
        ctx.tensors = t          # keeps the (possibly converted) inputs alive for backward
        ctx.inv_den = inv_den
        den_inv = inv_den[:NUM_TERMS].clone()
        ctx.mark_non_differentiable(den_inv)
        return terms, den_inv

    @staticmethod
    def backward(ctx, gterms, _gden=None):
        t, cfg = ctx.tensors, ctx.cfg
        need = dict(zip(_DIFF, ctx.needs_input_grad[2:]))
        mask = t["mask"]
        grads = {}
        with torch.cuda.device_of(mask):
            for k in _DIFF:
                grads[k] = torch.empty_like(t[k]) if (need[k] and t[k] is not None) else None
            # the KL kernels write mu and lv gradients as pairs
            for a, b in (("mu_l", "lv_l"), ("mu_g", "lv_g")):
                if (grads[a] is None) != (grads[b] is None):
                    for k in (a, b):
                        if grads[k] is None and t[k] is not None:
                            grads[k] = torch.empty_like(t[k])
            coef = f32c(gterms)
            args = _make_args(t, cfg)
            _lib.lib().call("pev_loss_bwd", ctypes.byref(args), ptr(coef), ptr(ctx.inv_den),
                            *[ptr(grads[k]) for k in _DIFF], stream(mask))
        return (None, None) + tuple(grads[k] if need[k] else None for k in _DIFF)


def _terms(cfg, mask, dp_normalize=False, **kw):
    """The 17 base terms.  ``dp_normalize``: this process holds one shard of a data-parallel global batch; every
    term is rescaled by ``world * den_local / sum_ranks(den_local)`` (one all-reduce of the 17 denominators), so that
    the plain average over ranks of the returned terms -- and of their gradients -- is the term of the GLOBAL batch
    even when ranks hold different numbers of conformers / valid residues (masked means, models/losses.py:12-21,
    :54-57)."""
    consts = {k: kw.get(k) for k in _CONST}
    consts["mask"] = mask
    t, den_inv = _LossTerms.apply(cfg, consts, *[kw.get(k) for k in _DIFF])
    if dp_normalize:
        from .distributed import shard_term_scale
        scale = shard_term_scale(den_inv)
        if scale is not None:
            t = torch.where(scale > 0, t * scale, torch.zeros_like(t))    # an empty shard contributes 0, not 0/0
    return t


# ------------------------------------------------------------------------------------ public API
def rmsd_loss(pred, target, mask):
    """Masked per-conformer MSE, mean over conformers (``models/losses.py:12-21``)."""
    return _terms({}, mask, pred_CA=pred, target_CA=target)[T_REC_CA]


def pair_distance_loss(pred, target, mask, stride=4, min_sep=2):
    """``models/losses.py:24-37``; ``min_sep`` is accepted and ignored exactly as there."""
    return _terms({"pair_stride": stride}, mask, pred_CA=pred, target_CA=target)[T_PAIR]


def _kl_unit_gauss(mu, lv, reduce_dims=None):
    """Element-wise KL to N(0,I) (``models/losses.py:40-46``); a plain torch helper."""
    kl = 0.5 * (lv.exp() + mu.pow(2) - 1.0 - lv)
    return kl if reduce_dims is None else kl.sum(dim=reduce_dims)


def kl_global(mu, lv):
    """``models/losses.py:49-51``."""
    mask = torch.ones(mu.shape[0], 1, dtype=torch.float32, device=mu.device)
    return _terms({}, mask, mu_g=mu, lv_g=lv)[T_KL_G]


def kl_local(mu, lv, mask):
    """``models/losses.py:54-57``."""
    return _terms({}, mask, mu_l=mu, lv_l=lv)[T_KL_L]


class _Dihedrals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, N, CA, C, mask):
        N, CA, C, mask = f32c(N), f32c(CA), f32c(C), f32c(mask)
        B, L = mask.shape
        with torch.cuda.device_of(CA):
            out = torch.empty(B, L, 6, dtype=torch.float32, device=CA.device)
            _lib.lib().call("pev_dihedrals_fwd", ptr(N), ptr(CA), ptr(C), ptr(mask), B, L, ptr(out), stream(CA))
        ctx.save_for_backward(N, CA, C, mask)
        return out

    @staticmethod
    def backward(ctx, gout):
        N, CA, C, mask = ctx.saved_tensors
        B, L = mask.shape
        with torch.cuda.device_of(CA):
            g = [torch.empty_like(N) for _ in range(3)]
            _lib.lib().call("pev_dihedrals_bwd", ptr(N), ptr(CA), ptr(C), ptr(mask), ptr(f32c(gout)), B, L,
                            ptr(g[0]), ptr(g[1]), ptr(g[2]), stream(CA))
        return g[0], g[1], g[2], None


def compute_dihedrals_from_coords(N, CA, C, mask):
    """``[B,L,6]`` sin/cos of phi, psi, omega; unset slots are 0 (``models/losses.py:235-308``)."""
    return _Dihedrals.apply(N, CA, C, mask)


class _DihedralTerms(torch.autograd.Function):
    """consistency / ramachandran / omega on an explicit dihedral tensor -> float32[3]."""

    @staticmethod
    def forward(ctx, dih, target, mask):
        dih, target, mask = f32c(dih), f32c(target), f32c(mask)
        B, L = mask.shape
        with torch.cuda.device_of(dih):
            sums = torch.zeros(6, dtype=torch.float64, device=dih.device)
            _lib.lib().call("pev_dihedral_terms_fwd", ptr(dih), ptr(target), ptr(mask), B, L, ptr(sums),
                            stream(dih))
        den = torch.stack([sums[1], sums[4], sums[4]])
        ctx.save_for_backward(dih, target, mask, den)
        return (torch.stack([sums[0], sums[2], sums[3]]) / den).float()

    @staticmethod
    def backward(ctx, g):
        dih, target, mask, den = ctx.saved_tensors
        B, L = mask.shape
        with torch.cuda.device_of(dih):
            coef = (g.double() / den).float().contiguous()
            gd = torch.empty_like(dih)
            _lib.lib().call("pev_dihedral_terms_bwd", ptr(dih), ptr(target), ptr(mask), ptr(coef), B, L,
                            ptr(gd), stream(dih))
        return gd, None, None


def dihedral_consistency_loss(pred_dihedrals, target_dihedrals, mask):
    """``models/losses.py:60-69``."""
    if pred_dihedrals is None or target_dihedrals is None:
        return torch.tensor(0.0, device=mask.device)
    return _DihedralTerms.apply(pred_dihedrals, target_dihedrals, mask)[0]


def ramachandran_loss(dihedrals, mask, aa_types=None):
    """``models/losses.py:72-131``."""
    if dihedrals.numel() == 0:
        return torch.tensor(0.0, device=mask.device)
    return _DihedralTerms.apply(dihedrals, None, mask)[1]


def ang_wrap(x):
    """Maps to (-pi, pi] (``models/losses.py:133-134``)."""
    return torch.atan2(torch.sin(x), torch.cos(x))


def omega_trans_loss(dihedrals, mask):
    """``models/losses.py:136-155``."""
    if dihedrals.numel() == 0:
        return torch.tensor(0.0, device=mask.device)
    return _DihedralTerms.apply(dihedrals, None, mask)[2]


def huber_loss(x, delta=0.2):
    """Element-wise Huber (``models/losses.py:311-316``); a plain torch helper."""
    a = torch.abs(x)
    return torch.where(a < delta, 0.5 * x ** 2, delta * (a - 0.5 * delta))


def bond_length_loss(pred_N, pred_CA, pred_C, mask):
    """``models/losses.py:318-355``."""
    t = _terms({"geometry": True}, mask, pred_N=pred_N, pred_CA=pred_CA, pred_C=pred_C)
    if pred_N.shape[1] > 1:
        return t[T_BOND_NCA] + t[T_BOND_CAC] + 2 * t[T_BOND_CN]
    return t[T_BOND_NCA] + t[T_BOND_CAC]


def bond_angle_loss(pred_N, pred_CA, pred_C, mask):
    """``models/losses.py:371-408``."""
    t = _terms({"geometry": True}, mask, pred_N=pred_N, pred_CA=pred_CA, pred_C=pred_C)
    if pred_N.shape[1] > 1:
        return t[T_ANG_NCAC] + 2.0 * (t[T_ANG_CNCA] + t[T_ANG_CACN])
    return t[T_ANG_NCAC]


def sequence_classification_loss(pred_seq_logits, target_seq_labels, mask):
    """``models/losses.py:411-437``."""
    return _terms({}, mask, logits=pred_seq_logits, labels=target_seq_labels)[T_SEQ]


def clash_loss(pred_N, pred_CA, pred_C, mask, clash_dist=3.2, soft_margin=0.5):
    """``models/losses.py:439-517``; never materialises the ``[B,3L,3L]`` matrices."""
    cfg = {"clash": True, "clash_dist": clash_dist, "soft_margin": soft_margin}
    return _terms(cfg, mask, pred_N=pred_N, pred_CA=pred_CA, pred_C=pred_C)[T_CLASH]


_COEF_CACHE: dict = {}


def _coef_vector(values: tuple, device):
    """Device copy of a coefficient vector and the mask of its non-zero entries, cached per (values, device)."""
    key = (values, str(device))
    hit = _COEF_CACHE.get(key)
    if hit is None:
        if len(_COEF_CACHE) > 256:
            _COEF_CACHE.clear()
        coef = torch.tensor(values, dtype=torch.float32, device=device)
        hit = _COEF_CACHE[key] = (coef, coef != 0)
    return hit


def compute_total_loss(pred_N, pred_CA, pred_C, pred_seq, target_N, target_CA, target_C, target_seq_labels,
                       mask, mu_g, lv_g, mu_l, lv_l,
                       target_dihedrals, klw_g, klw_l, w_pair, pair_stride,
                       w_dihedral, w_rama, w_bond, w_angle, w_rec, w_seq, w_clash, *, dp_normalize=False):
    """Weighted total and its 16 components (``models/losses.py:520-613``).

    ``dp_normalize=True`` (keyword only, not in the reference): the batch is one rank's shard of a data-parallel
    global batch -- see :func:`_terms`; averaging the returned values / gradients over ranks then equals the
    single-process result on the concatenated batch, also for ragged shards."""
    t = _terms({"pair_stride": pair_stride, "clash": True, "geometry": True}, mask, dp_normalize,
               pred_N=pred_N, pred_CA=pred_CA, pred_C=pred_C, logits=pred_seq, mu_l=mu_l, lv_l=lv_l,
               mu_g=mu_g, lv_g=lv_g, target_N=target_N, target_CA=target_CA, target_C=target_C,
               target_dih=target_dihedrals, labels=target_seq_labels)
    multi = pred_N.shape[1] > 1
    cons = t[T_DIH_CONS] if target_dihedrals is not None else torch.zeros((), device=mask.device)
    loss_rec = t[T_REC_CA] + 0.5 * (t[T_REC_N] + t[T_REC_C])
    loss_dihedral = cons + t[T_OMEGA]
    loss_bond = t[T_BOND_NCA] + t[T_BOND_CAC] + (2 * t[T_BOND_CN] if multi else 0.0)
    loss_angle = t[T_ANG_NCAC] + (2.0 * (t[T_ANG_CNCA] + t[T_ANG_CACN]) if multi else 0.0)
    ws = (w_rec, w_pair, klw_g, klw_l, w_dihedral, w_rama, w_bond, w_angle, w_seq, w_clash)
    if all(isinstance(w, (int, float)) for w in ws):
        # Scalar weights: the same weighted sum as one dot product with a cached coefficient vector.  The component
        # expressions above stay available in the dict, but `total.backward()` no longer walks ~20 select / mul / add
        # nodes (each a zero-fill + copy + add of a [NUM_TERMS] tensor: ~100 launch-bound kernels per step).
        c = [0.0] * NUM_TERMS
        c[T_REC_CA], c[T_REC_N], c[T_REC_C] = w_rec, 0.5 * w_rec, 0.5 * w_rec
        c[T_PAIR], c[T_KL_G], c[T_KL_L] = w_pair, klw_g, klw_l
        c[T_OMEGA] = w_dihedral
        if target_dihedrals is not None:
            c[T_DIH_CONS] = w_dihedral
        c[T_RAMA], c[T_SEQ], c[T_CLASH] = w_rama, w_seq, w_clash
        c[T_BOND_NCA] = c[T_BOND_CAC] = w_bond
        c[T_ANG_NCAC] = w_angle
        if multi:
            c[T_BOND_CN] = 2 * w_bond
            c[T_ANG_CNCA] = c[T_ANG_CACN] = 2.0 * w_angle
        coef, used = _coef_vector(tuple(float(v) for v in c), t.device)
        loss = torch.dot(torch.where(used, t, t.new_zeros(())), coef)     # unused terms may be undefined (0/0)
    else:
        loss = (w_rec * loss_rec + w_pair * t[T_PAIR] + klw_g * t[T_KL_G] + klw_l * t[T_KL_L]
                + w_dihedral * loss_dihedral + w_rama * t[T_RAMA] + w_bond * loss_bond
                + w_angle * loss_angle + w_seq * t[T_SEQ] + w_clash * t[T_CLASH])
    return {
        "total": loss,
        "reconstruction": loss_rec,
        "reconstruction_ca": t[T_REC_CA],
        "reconstruction_n": t[T_REC_N],
        "reconstruction_c": t[T_REC_C],
        "pair_distance": t[T_PAIR],
        "kl_global": t[T_KL_G],
        "kl_local": t[T_KL_L],
        "dihedral_consistency": cons,
        "omega_trans": t[T_OMEGA],
        "ramachandran": t[T_RAMA],
        "dihedral_total": loss_dihedral,
        "bond_length": loss_bond,
        "bond_angle": loss_angle,
        "sequence": t[T_SEQ],
        "clash": t[T_CLASH],
    }
