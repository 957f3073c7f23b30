"""Floating-point oracle: geometric / KL / sequence losses (torch on CPU).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Restates
``models/losses.py`` of the reference.  Distances are taken as
``sqrt(sum(diff**2))`` rather than through ``torch.cdist``'s matmul path
(SURVEY.md F7): in float64 the two agree to ~1e-8, in float32 only this form is
accurate, and it is the form the CUDA kernels evaluate.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

RAMA_BASINS = (  # (phi0, psi0, width)  models/losses.py:92-113
    (-1.05, -0.79, 0.6),
    (-2.09, 2.09, 0.9),
    (1.05, 0.79, 0.6),
    (-1.31, 2.53, 0.5),
)


def _pdist(a):
    # zero distances (the diagonal) get value 0 and gradient 0, as torch.cdist's backward gives
    d = a[:, :, None, :] - a[:, None, :, :]
    sq = (d * d).sum(-1)
    pos = sq > 0
    return torch.where(pos, sq, torch.ones_like(sq)).sqrt() * pos.to(sq.dtype)


def rmsd_loss(pred, target, mask):
    """Masked MSE per conformer, mean over conformers (``models/losses.py:12-21``)."""
    sq = ((pred - target) ** 2).sum(-1)
    return ((sq * mask).sum(1) / mask.sum(1)).mean()


def pair_distance_loss(pred, target, mask, stride=4, min_sep=2):
    """Strided all-pairs |d_pred - d_target| (``models/losses.py:24-37``); ``min_sep`` unused there too."""
    sel = torch.arange(0, pred.shape[1], stride)
    m = mask[:, sel]
    pm = (m[:, :, None] * m[:, None, :]).to(pred.dtype)
    dp, dt = _pdist(pred[:, sel]), _pdist(target[:, sel])
    return ((dp - dt).abs() * pm).sum() / pm.sum()


def _kl(mu, lv):
    return 0.5 * (lv.exp() + mu * mu - 1.0 - lv)      # models/losses.py:43


def kl_global(mu, lv):
    return _kl(mu, lv).sum(1).mean()                  # models/losses.py:49-51


def kl_local(mu, lv, mask):
    return (_kl(mu, lv).sum(-1) * mask).sum() / mask.sum()   # models/losses.py:54-57


def dihedral_sincos(p0, p1, p2, p3, eps=1e-8):
    """sin/cos of the torsion p0-p1-p2-p3 (``models/losses.py:158-232``)."""
    b1, b2, b3 = p1 - p0, p2 - p1, p3 - p2
    n1 = torch.linalg.cross(b1, b2)
    n2 = torch.linalg.cross(b2, b3)
    l1 = n1.norm(dim=-1, keepdim=True)
    l2 = n2.norm(dim=-1, keepdim=True)
    ok = (l1.squeeze(-1) > eps) & (l2.squeeze(-1) > eps)
    okv = ok.unsqueeze(-1)
    zero = torch.zeros_like(n1)
    u1 = torch.where(okv, n1 / (l1 + eps), zero)
    u2 = torch.where(okv, n2 / (l2 + eps), zero)
    ub = torch.where(okv, b2 / (b2.norm(dim=-1, keepdim=True) + eps), zero)
    c = (u1 * u2).sum(-1).clamp(-1.0 + eps, 1.0 - eps)
    s = torch.sign((torch.linalg.cross(u1, u2) * ub).sum(-1)) * torch.sqrt(1.0 - c * c + eps)
    return torch.where(ok, s, torch.zeros_like(s)), torch.where(ok, c, torch.ones_like(c))


def compute_dihedrals_from_coords(N, CA, C, mask):
    """[B,L,6] = (sin,cos) of phi, psi, omega; unset slots stay 0 (``models/losses.py:235-308``)."""
    B, L, _ = CA.shape
    out = torch.zeros(B, L, 6, dtype=CA.dtype)
    if L < 2:
        return out
    m = mask.bool()
    pair = m[:, :-1] & m[:, 1:]
    z = torch.zeros(B, L - 1, dtype=CA.dtype)
    sp, cp = dihedral_sincos(C[:, :-1], N[:, 1:], CA[:, 1:], C[:, 1:])       # phi(i), i>=1
    ss, cs = dihedral_sincos(N[:, :-1], CA[:, :-1], C[:, :-1], N[:, 1:])     # psi(i), i<=L-2
    so, co = dihedral_sincos(CA[:, :-1], C[:, :-1], N[:, 1:], CA[:, 1:])     # omega(i), i>=1
    pad = torch.zeros(B, 1, dtype=CA.dtype)
    cols = [
        torch.cat([pad, torch.where(pair, sp, z)], 1),
        torch.cat([pad, torch.where(pair, cp, z)], 1),
        torch.cat([torch.where(pair, ss, z), pad], 1),
        torch.cat([torch.where(pair, cs, z), pad], 1),
        torch.cat([pad, torch.where(pair, so, z)], 1),
        torch.cat([pad, torch.where(pair, co, z)], 1),
    ]
    return torch.stack(cols, -1)


def dihedral_consistency_loss(pred_dih, target_dih, mask):
    """``models/losses.py:60-69``."""
    valid = mask.unsqueeze(-1).bool() & torch.isfinite(pred_dih) & torch.isfinite(target_dih)
    diff = torch.where(valid, pred_dih - target_dih, torch.zeros_like(pred_dih))
    return (diff ** 2).sum() / valid.to(pred_dih.dtype).sum()


def ramachandran_loss(dih, mask):
    """``models/losses.py:72-131``."""
    phi = torch.atan2(dih[..., 0], dih[..., 1])
    psi = torch.atan2(dih[..., 2], dih[..., 3])
    best = None
    for p0, s0, wdt in RAMA_BASINS:
        g = torch.exp(-((phi - p0) ** 2 / wdt + (psi - s0) ** 2 / wdt))
        best = g if best is None else torch.maximum(best, g)
    pen = 1.0 - best + 5.0 * ((phi > 0) & (psi < 0)).to(dih.dtype)
    return (pen * mask).sum() / mask.sum()


def omega_trans_loss(dih, mask):
    """``models/losses.py:136-155`` (``ang_wrap`` at ``:133-134``)."""
    om = torch.atan2(dih[..., 4], dih[..., 5])
    wrapped = torch.atan2(torch.sin(om), torch.cos(om))
    pen = 2.0 * (1.0 - torch.cos(om - math.pi)) + 3.0 * (wrapped.abs() < 0.5).to(dih.dtype)
    return (pen * mask).sum() / mask.sum()


def huber(x, delta):
    """``models/losses.py:311-316``."""
    a = x.abs()
    return torch.where(a < delta, 0.5 * x * x, delta * (a - 0.5 * delta))


def bond_length_loss(N, CA, C, mask):
    """``models/losses.py:318-355``."""
    nca = (huber((CA - N).norm(dim=-1) - 1.46, 0.02) * mask).sum() / mask.sum()
    cac = (huber((C - CA).norm(dim=-1) - 1.52, 0.02) * mask).sum() / mask.sum()
    if N.shape[1] > 1:
        pm = mask[:, :-1] * mask[:, 1:]
        cn = (huber((N[:, 1:] - C[:, :-1]).norm(dim=-1) - 1.33, 0.01) * pm).sum() / pm.sum()
    else:
        cn = torch.zeros((), dtype=N.dtype)
    return nca + cac + 2 * cn


def _angle(a, b, c, eps=1e-8):
    """Angle at b (``models/losses.py:358-368`` + ``acos``)."""
    u = a - b
    v = c - b
    u = u / (u.norm(dim=-1, keepdim=True) + eps)
    v = v / (v.norm(dim=-1, keepdim=True) + eps)
    return torch.acos((u * v).sum(-1).clamp(-1.0, 1.0))


def bond_angle_loss(N, CA, C, mask):
    """``models/losses.py:371-408``."""
    mask = mask.to(N.dtype)
    rad = math.pi / 180.0
    l1 = (huber(_angle(N, CA, C) - 110.0 * rad, 0.1) * mask).sum() / mask.sum()
    if N.shape[1] > 1:
        pm = mask[:, :-1] * mask[:, 1:]
        l2 = (huber(_angle(C[:, :-1], N[:, 1:], CA[:, 1:]) - 121.0 * rad, 0.1) * pm).sum() / pm.sum()
        l3 = (huber(_angle(CA[:, :-1], C[:, :-1], N[:, 1:]) - 116.0 * rad, 0.1) * pm).sum() / pm.sum()
    else:
        l2 = l3 = torch.zeros((), dtype=N.dtype)
    return l1 + 2.0 * (l2 + l3)


def sequence_classification_loss(logits, labels, mask):
    """``models/losses.py:411-437``."""
    ce = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), reduction="none")
    mf = mask.reshape(-1)
    return (ce * mf).sum() / (mf.sum() + 1e-8)


def clash_loss(N, CA, C, mask, clash_dist=3.2, soft_margin=0.5):
    """``models/losses.py:439-517``: interleaved N,CA,C atoms, residue separation >= 2, i<j."""
    B, L = CA.shape[:2]
    atoms = torch.stack([N, CA, C], 2).reshape(B, 3 * L, 3)
    am = torch.stack([mask, mask, mask], 2).reshape(B, 3 * L).to(N.dtype)
    res = torch.arange(3 * L) // 3
    sep = ((res[:, None] - res[None, :]).abs() >= 2).to(N.dtype)
    upper = torch.triu(torch.ones(3 * L, 3 * L, dtype=N.dtype), diagonal=1)
    pm = am[:, :, None] * am[:, None, :] * sep[None] * upper[None]
    v = F.relu(clash_dist - _pdist(atoms))
    pen = torch.where(v < soft_margin, 0.5 * v * v, v * v)
    return ((pen * pm).sum((1, 2)) / (pm.sum((1, 2)) + 1e-8)).mean()


LOSS_KEYS = ("total", "reconstruction", "reconstruction_ca", "reconstruction_n", "reconstruction_c",
             "pair_distance", "kl_global", "kl_local", "dihedral_consistency", "omega_trans",
             "ramachandran", "dihedral_total", "bond_length", "bond_angle", "sequence", "clash")


def compute_total_loss(pred_N, pred_CA, pred_C, pred_seq, target_N, target_CA, target_C,
                       target_seq_labels, mask, mu_g, lv_g, mu_l, lv_l, target_dihedrals,
                       klw_g, klw_l, w_pair, pair_stride, w_dihedral, w_rama, w_bond, w_angle,
                       w_rec, w_seq, w_clash):
    """Weighted total and the 16-key dict (``models/losses.py:520-613``)."""
    ca = rmsd_loss(pred_CA, target_CA, mask)
    n = rmsd_loss(pred_N, target_N, mask)
    c = rmsd_loss(pred_C, target_C, mask)
    rec = ca + 0.5 * (n + c)
    pair = pair_distance_loss(pred_CA, target_CA, mask, stride=pair_stride)
    kg = kl_global(mu_g, lv_g)
    kl = kl_local(mu_l, lv_l, mask)
    dih = compute_dihedrals_from_coords(pred_N, pred_CA, pred_C, mask)
    cons = dihedral_consistency_loss(dih, target_dihedrals, mask)
    rama = ramachandran_loss(dih, mask)
    omega = omega_trans_loss(dih, mask)
    dtot = cons + omega
    bond = bond_length_loss(pred_N, pred_CA, pred_C, mask)
    angle = bond_angle_loss(pred_N, pred_CA, pred_C, mask)
    seq = sequence_classification_loss(pred_seq, target_seq_labels, mask)
    clash = clash_loss(pred_N, pred_CA, pred_C, mask)
    total = (w_rec * rec + w_pair * pair + klw_g * kg + klw_l * kl + w_dihedral * dtot
             + w_rama * rama + w_bond * bond + w_angle * angle + w_seq * seq + w_clash * clash)
    vals = (total, rec, ca, n, c, pair, kg, kl, cons, omega, rama, dtot, bond, angle, seq, clash)
    return dict(zip(LOSS_KEYS, vals))
