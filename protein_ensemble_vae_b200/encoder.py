"""``ProteinEncoder`` -- ``models/encoder.py:14-262`` on the device, same classes, constructor arguments,
``state_dict`` keys and return values.

The reference runs the encoder on zero-padded ``[B,L,...]`` tensors through ``nn.TransformerEncoderLayer`` /
``nn.MultiheadAttention``.  Here the valid residues of the batch are PACKED (``N = sum Lb`` rows, as in the decoder), every
linear with 256-multiple widths runs on the repo's tcgen05 GEMM kernels (``tc_linear.py``: TF32, or fp32-accurate 3xTF32
with ``precision="fp32"``) with bias / ReLU / residual in the epilogue, LayerNorms on ``pev_add_layernorm_fwd``, and
attention runs per conformer over its own residues (``attention.py``), which is what the key-padding mask of the reference
(``:125-137``) computes for the valid rows.  The ``nn.TransformerEncoderLayer`` / ``nn.MultiheadAttention`` sub-modules are
kept as parameter holders so checkpoints load unchanged (``models/training.py:456``).

Differences, all at PADDED positions only: the reference returns whatever its padded rows computed there (they never reach
a valid row or a masked loss); here ``mu_l`` / ``lv_l`` / ``z_l`` are exact zeros at padding.  A conformer with no valid
residue gives NaN in the reference (softmax over an empty set) and zeros here.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import attention as pattn
from .egnn_tc import layer_norm
from .tc_linear import linear


class SinusoidalPE(nn.Module):
    """``models/encoder.py:14-27``."""

    def __init__(self, d_model: int, max_len: int = 4096):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return x + self.pe[:x.size(1)]


class _Packing:
    """Valid-row bookkeeping of one batch: flat indices of the valid ``(b, l)``, per-conformer lengths and offsets."""

    def __init__(self, mask, B, L, device):
        if mask is None:
            self.lengths = [L] * B
            self.idx = None
        else:
            mb = mask.bool()
            self.lengths = mb.sum(1).tolist()                              # one host sync per forward (as the decoder)
            self.idx = torch.nonzero(mb.reshape(-1)).squeeze(-1) if sum(self.lengths) != B * L else None
        self.B, self.L, self.N = B, L, int(sum(self.lengths))
        cu = [0]
        for n in self.lengths:
            cu.append(cu[-1] + n)
        self.cu_host = cu
        self.cu = torch.tensor(cu, dtype=torch.int32, device=device)
        self.lmax = max(self.lengths, default=0)
        flat = torch.arange(B * L, device=device) if self.idx is None else self.idx
        self.pos = flat % L                                                  # position inside the conformer (for the PE)
        self.conf = flat // L

    def pack(self, t):
        flat = t.reshape(self.B * self.L, *t.shape[2:])
        return flat if self.idx is None else flat.index_select(0, self.idx)

    def unpack(self, t):
        if self.idx is None:
            return t.reshape(self.B, self.L, *t.shape[1:])
        full = torch.zeros(self.B * self.L, *t.shape[1:], device=t.device, dtype=t.dtype)
        return full.index_copy(0, self.idx, t).reshape(self.B, self.L, *t.shape[1:])


class DihedralAwareEncoder(nn.Module):
    """``models/encoder.py:30-143``."""

    def __init__(self, seq_dim: int, dihedral_dim: int, d_model: int, nhead: int, ff: int, nlayers: int,
                 dropout: float = 0.1, d_pair: int = 128, precision: str = "tf32"):
        super().__init__()
        self.seq_dim, self.dihedral_dim, self.d_model = seq_dim, dihedral_dim, d_model
        self.seq_proj = nn.Linear(seq_dim, d_model // 2)
        self.dihedral_proj = nn.Linear(dihedral_dim, d_model // 4)
        self.coord_proj = nn.Linear(9, d_model // 4)
        self.coord_norm = nn.LayerNorm(d_model // 4)
        self.dihedral_norm = nn.LayerNorm(d_model // 4)
        self.feature_fusion = nn.Sequential(nn.Linear(d_model, d_model), nn.LayerNorm(d_model), nn.ReLU(), nn.Dropout(dropout))
        self.pe = SinusoidalPE(d_model)
        self.nhead, self.nlayers = nhead, nlayers
        self.transformer_layers = nn.ModuleList([
            nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=ff, dropout=dropout, norm_first=True,
                                       batch_first=True) for _ in range(nlayers)])
        self.ln = nn.LayerNorm(d_model)
        self.geom_res_scale = nn.Parameter(torch.tensor(0.1))
        self.geometric_attention = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead // 2, dropout=dropout,
                                                         batch_first=True)
        self.precision = precision

    def _mha(self, mha: nn.MultiheadAttention, x, pk: _Packing, res=None, p_drop=0.0):
        """Self-attention of one ``nn.MultiheadAttention`` parameter set over the packed rows: QKV projection (one GEMM),
        per-conformer attention, output projection with dropout and the residual in its epilogue when ``res`` is given."""
        precise = self.precision == "fp32"
        qkv = linear(x, mha.in_proj_weight, mha.in_proj_bias, precise=precise)            # [N, 3 d]
        a = pattn.self_attention(qkv, pk, mha.num_heads, mha.dropout if self.training else 0.0, precise)
        return linear(a, mha.out_proj.weight, mha.out_proj.bias, precise=precise, res=res, p_drop=p_drop)

    def forward_packed(self, sequence_emb, n_coords, ca_coords, c_coords, dihedrals, pk: _Packing):
        precise = self.precision == "fp32"
        # per-residue features (:104-121)
        backbone = pk.pack(torch.cat([n_coords, ca_coords, c_coords], dim=-1))               # [N, 9]
        coord_feat = layer_norm(self.coord_norm, self.coord_proj(backbone))
        dihedral_feat = layer_norm(self.dihedral_norm, self.dihedral_proj(pk.pack(dihedrals)))
        seq_feat = linear(pk.pack(sequence_emb), self.seq_proj.weight, self.seq_proj.bias, precise=precise)
        combined = torch.cat([seq_feat, coord_feat, dihedral_feat], dim=-1)
        f = linear(combined, self.feature_fusion[0].weight, self.feature_fusion[0].bias, precise=precise)
        f = self.feature_fusion[3](torch.relu(layer_norm(self.feature_fusion[1], f)))
        f = f + self.pe.pe.index_select(0, pk.pos)                                           # (:124)
        # geometric attention (:126-134)
        f = f + self.geom_res_scale * self._mha(self.geometric_attention, f, pk)
        # transformer layers, pre-norm (:139-140; nn.TransformerEncoderLayer with norm_first=True, ReLU)
        for layer in self.transformer_layers:
            p1, p, p2 = ((layer.dropout1.p, layer.dropout.p, layer.dropout2.p) if self.training else (0.0, 0.0, 0.0))
            f = self._mha(layer.self_attn, layer_norm(layer.norm1, f), pk, res=f, p_drop=p1)          # x + drop1(SA(norm1 x))
            h1 = linear(layer_norm(layer.norm2, f), layer.linear1.weight, layer.linear1.bias, relu=True, precise=precise,
                        p_drop=p)                                                                    # drop(relu(linear1))
            f = linear(h1, layer.linear2.weight, layer.linear2.bias, precise=precise, res=f, p_drop=p2)
        return layer_norm(self.ln, f)                                                        # (:143)

    def forward(self, sequence_emb, n_coords, ca_coords, c_coords, dihedrals, mask):
        """``[B,L,d_model]`` encoded features (zeros at padding)."""
        B, L = ca_coords.shape[:2]
        pk = _Packing(mask, B, L, ca_coords.device)
        return pk.unpack(self.forward_packed(sequence_emb, n_coords, ca_coords, c_coords, dihedrals, pk))


class HierLatent(nn.Module):
    """``models/encoder.py:149-216``."""

    def __init__(self, d_model: int, z_g: int = 64, z_l: int = 32, precision: str = "tf32"):
        super().__init__()
        self.d_model, self.z_g, self.z_l = d_model, z_g, z_l
        self.global_attention = nn.MultiheadAttention(embed_dim=d_model, num_heads=4, dropout=0.1, batch_first=True)
        self.global_query = nn.Parameter(torch.randn(1, 1, d_model))
        self.global_head = nn.Sequential(nn.Linear(d_model, 256), nn.ReLU(), nn.Linear(256, 2 * z_g))
        self.local_head = nn.Sequential(nn.Linear(d_model, 256), nn.ReLU(), nn.Linear(256, 2 * z_l))
        with torch.no_grad():
            self.global_head[-1].bias[z_g:] = -2.0
            self.local_head[-1].bias[z_l:] = -2.0
            self.global_query.data.normal_(0, 0.02)
        self.precision = precision

    def forward_packed(self, H, pk: _Packing):
        """``H`` packed ``[N,d]`` -> ``mu_g, lv_g [B,zg]``, ``mu_l, lv_l [B,L,zl]``."""
        precise = self.precision == "fp32"
        d = self.d_model
        ga = self.global_attention
        # attention pooling with one learned query per conformer (:188-197): K / V projections in one GEMM
        kv = linear(H, ga.in_proj_weight[d:], ga.in_proj_bias[d:], precise=precise)                    # [N, 2 d]
        q = F.linear(self.global_query.reshape(1, d), ga.in_proj_weight[:d], ga.in_proj_bias[:d])      # [1, d]
        pooled = pattn.pooled_attention(q, kv, pk, ga.num_heads, ga.dropout if self.training else 0.0)  # [B, d]
        g = self.global_head(F.linear(pooled, ga.out_proj.weight, ga.out_proj.bias))
        mu_g, lv_g = torch.chunk(g, 2, dim=-1)
        # local latents (:207-209)
        hid = linear(H, self.local_head[0].weight, self.local_head[0].bias, relu=True, precise=precise)
        loc = pk.unpack(linear(hid, self.local_head[2].weight, self.local_head[2].bias, precise=precise))
        mu_l, lv_l = torch.chunk(loc, 2, dim=-1)
        return mu_g, lv_g, mu_l, lv_l

    def forward(self, H, mask):
        B, L, _ = H.shape
        pk = _Packing(mask, B, L, H.device)
        return self.forward_packed(pk.pack(H), pk)


class ProteinEncoder(nn.Module):
    """``models/encoder.py:219-262``.  ``precision``: ``"tf32"`` (tensor-core TF32 linears, default) or ``"fp32"``
    (3xTF32 linears, exact attention).  ``forward(..., eps_g=None, eps_l=None)`` accepts the reparameterisation noise
    (tests); by default it is drawn as in the reference (``:231-236``)."""

    def __init__(self, seqemb_dim: int, d_model: int = 512, nhead: int = 8, ff: int = 1024, nlayers: int = 6,
                 z_g: int = 512, z_l: int = 256, dropout: float = 0.1, use_dihedrals: bool = True, precision: str = "tf32"):
        super().__init__()
        if precision not in ("tf32", "fp32"):
            raise ValueError("precision must be 'tf32' or 'fp32'")
        self.seqemb_dim, self.use_dihedrals = seqemb_dim, use_dihedrals
        self.enc = DihedralAwareEncoder(seqemb_dim, dihedral_dim=6, d_model=d_model, nhead=nhead, ff=ff, nlayers=nlayers,
                                        dropout=dropout, precision=precision)
        self.latent = HierLatent(d_model, z_g, z_l, precision=precision)

    def reparam(self, mu, lv, eps=None):
        std = torch.exp(0.5 * lv)
        return mu + (torch.randn_like(std) if eps is None else eps) * std

    def forward(self, seqemb, n_coords, ca_coords, c_coords, dihedrals, mask, eps_g=None, eps_l=None):
        if not ca_coords.is_cuda:
            raise RuntimeError("ProteinEncoder runs on CUDA tensors only")
        B, L = ca_coords.shape[:2]
        pk = _Packing(mask, B, L, ca_coords.device)
        H = self.enc.forward_packed(seqemb, n_coords, ca_coords, c_coords, dihedrals, pk)
        mu_g, lv_g, mu_l, lv_l = self.latent.forward_packed(H, pk)
        z_g = self.reparam(mu_g, lv_g, eps_g)
        z_l = self.reparam(mu_l, lv_l, eps_l)
        if pk.idx is not None:
            z_l = z_l * mask.to(z_l.dtype).unsqueeze(-1)                   # zeros at padding (the reference leaves noise)
        return z_g, z_l, mu_g, lv_g, mu_l, lv_l
