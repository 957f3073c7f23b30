"""Oracle: ``ProteinEncoder`` forward (numpy float64).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Restatement of ``models/encoder.py``: SinusoidalPE (:14-27),
DihedralAwareEncoder.forward (:83-143; ``nn.TransformerEncoderLayer(norm_first=True)`` = pre-norm self-attention + ReLU
feed-forward blocks, ``nn.MultiheadAttention`` = scaled dot-product attention over the un-masked keys), HierLatent.forward
(:177-216).  Pinned on ``tests/golden/encoders.npz`` (outputs of the reference itself, ``make_golden.py::gen_encoders``).
Dropout is not modelled (evaluation / p = 0)."""
from __future__ import annotations

import numpy as np


def _ln(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _lin(sd, name, x):
    return x @ sd[name + ".weight"].T + sd[name + ".bias"]


def _mha(sd, prefix, xq, xkv, keymask, nheads):
    """nn.MultiheadAttention(batch_first=True)(xq, xkv, xkv, key_padding_mask=~keymask) for ONE sample: [Lq,d],[Lk,d]."""
    d = xq.shape[-1]
    W, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q, k, v = xq @ W[:d].T + b[:d], xkv @ W[d:2 * d].T + b[d:2 * d], xkv @ W[2 * d:].T + b[2 * d:]
    hd = d // nheads
    out = np.zeros_like(q)
    for h in range(nheads):
        s = (q[:, h * hd:(h + 1) * hd] @ k[:, h * hd:(h + 1) * hd].T) / np.sqrt(hd)
        s = np.where(keymask[None, :], s, -np.inf)
        p = np.exp(s - s.max(-1, keepdims=True))
        out[:, h * hd:(h + 1) * hd] = (p / p.sum(-1, keepdims=True)) @ v[:, h * hd:(h + 1) * hd]
    return out @ sd[prefix + ".out_proj.weight"].T + sd[prefix + ".out_proj.bias"]


def sinusoidal_pe(L, d):
    """:17-22."""
    pe = np.zeros((L, d))
    pos = np.arange(L)[:, None]
    div = np.exp(np.arange(0, d, 2) * (-np.log(10000.0) / d))
    pe[:, 0::2], pe[:, 1::2] = np.sin(pos * div), np.cos(pos * div)
    return pe


def encoder(sd, seq_emb, n, ca, c, dih, mask, nhead=8, pe=None):
    """-> (H [B,L,d], mu_g, lv_g [B,zg], mu_l, lv_l [B,L,zl]); every row is computed as in the reference (padded ones too).
    ``pe``: the positional table to use (the reference builds its buffer in float32; default: float64 values)."""
    sd = {k: np.asarray(v, np.float64) for k, v in sd.items()}
    B, L = ca.shape[:2]
    d = sd["enc.ln.weight"].shape[0]
    nlayers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("enc.transformer_layers."))
    Hs, outs = [], []
    for b in range(B):
        km = mask[b].astype(bool)
        coord = _ln(_lin(sd, "enc.coord_proj", np.concatenate([n[b], ca[b], c[b]], -1)), sd["enc.coord_norm.weight"],
                    sd["enc.coord_norm.bias"])                                                        # :108-109
        dfe = _ln(_lin(sd, "enc.dihedral_proj", dih[b]), sd["enc.dihedral_norm.weight"], sd["enc.dihedral_norm.bias"])
        f = np.concatenate([_lin(sd, "enc.seq_proj", seq_emb[b]), coord, dfe], -1)                    # :114-115
        f = np.maximum(_ln(_lin(sd, "enc.feature_fusion.0", f), sd["enc.feature_fusion.1.weight"],
                           sd["enc.feature_fusion.1.bias"]), 0.0)                                     # :118
        f = f + (sinusoidal_pe(L, d) if pe is None else np.asarray(pe[:L], np.float64))                                                                   # :121
        f = f + sd["enc.geom_res_scale"] * _mha(sd, "enc.geometric_attention", f, f, km, nhead // 2)  # :126-132
        for i in range(nlayers):                                                                      # :139-140
            p = f"enc.transformer_layers.{i}"
            x = _ln(f, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
            f = f + _mha(sd, p + ".self_attn", x, x, km, nhead)
            x = _ln(f, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
            f = f + _lin(sd, p + ".linear2", np.maximum(_lin(sd, p + ".linear1", x), 0.0))
        H = _ln(f, sd["enc.ln.weight"], sd["enc.ln.bias"])                                            # :143
        Hs.append(H)
        pooled = _mha(sd, "latent.global_attention", sd["latent.global_query"].reshape(1, d), H, km, 4)   # :188-197
        g = _lin(sd, "latent.global_head.2", np.maximum(_lin(sd, "latent.global_head.0", pooled), 0.0))[0]
        loc = _lin(sd, "latent.local_head.2", np.maximum(_lin(sd, "latent.local_head.0", H), 0.0))     # :207
        outs.append((g, loc))
    g = np.stack([o[0] for o in outs])
    loc = np.stack([o[1] for o in outs])
    zg, zl = g.shape[-1] // 2, loc.shape[-1] // 2
    return np.stack(Hs), g[:, :zg], g[:, zg:], loc[..., :zl], loc[..., zl:]
