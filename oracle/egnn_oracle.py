"""Floating-point oracle: EGNN layer and decoder (torch on CPU, any dtype).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Functional
restatement of ``models/en_gnn_decoder.py`` of the reference, operating on a
plain ``state_dict`` with the reference's parameter names so that a module
under test and the oracle share weights through ``state_dict()``.
Run in float64 it is the parity oracle; run in float32 it is
the CPU baseline that ``bench.py`` times.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import graph_oracle

N_CA_LENGTH = 1.46   # models/en_gnn_decoder.py:274
CA_C_LENGTH = 1.52   # models/en_gnn_decoder.py:275
PEPTIDE_LENGTH = 1.33  # models/en_gnn_decoder.py:304


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def egn_layer(sd, prefix, h, x, edge_index, degree_inv=None):
    """One EGNN layer.  Reference: ``EGNLayer.forward`` (``models/en_gnn_decoder.py:53-87``).

    ``sd[prefix + 'phi_e.0.weight']`` etc.  ``edge_index`` int64 ``[2,E]``, rows
    are destinations.  Sums run in edge order through ``index_add_`` exactly as
    the reference does (``:68-69``, ``:78-79``).
    """
    p = prefix
    row, col = edge_index[0], edge_index[1]
    rel = x[row] - x[col]                                             # :61
    d2 = (rel * rel).sum(-1, keepdim=True)                            # :62
    e_in = torch.cat([h[row], h[col], d2], dim=-1)                    # :65
    m = F.silu(_lin(sd, p + "phi_e.2", F.silu(_lin(sd, p + "phi_e.0", e_in))))
    agg = torch.zeros(h.shape[0], m.shape[1], dtype=h.dtype)
    agg.index_add_(0, row, m)                                         # :68-69
    upd = _lin(sd, p + "phi_h.2", F.silu(_lin(sd, p + "phi_h.0", torch.cat([h, agg], -1))))
    h_new = F.layer_norm(h + upd, (h.shape[1],), sd[p + "norm_h.weight"],
                         sd[p + "norm_h.bias"], 1e-5)                  # :72-73
    w = _lin(sd, p + "phi_x.2", F.silu(_lin(sd, p + "phi_x.0", m)))   # :76
    delta = torch.zeros_like(x)
    delta.index_add_(0, row, w * rel)                                 # :78-79
    if degree_inv is not None:
        delta = delta * degree_inv[:, None]                           # :81-82
    delta = delta * 0.2                                               # :85
    return h_new, x + delta                                           # :86


def _latent_to_coords(sd, z, drop=None):
    # models/en_gnn_decoder.py:124-132 (dropout :128 is identity in eval; `drop` injects a train-mode mask)
    y = _lin(sd, "latent_to_coords.0", z)
    y = F.relu(F.layer_norm(y, (y.shape[-1],), sd["latent_to_coords.1.weight"],
                            sd["latent_to_coords.1.bias"], 1e-5))
    if drop is not None:
        y = drop(y, 0.5)
    y = F.relu(_lin(sd, "latent_to_coords.4", y))
    return _lin(sd, "latent_to_coords.6", y)


def _sequence_head(sd, h, drop=None):
    # models/en_gnn_decoder.py:162-172 (dropouts :166, :170)
    y = _lin(sd, "sequence_head.0", h)
    y = F.relu(F.layer_norm(y, (y.shape[-1],), sd["sequence_head.1.weight"],
                            sd["sequence_head.1.bias"], 1e-5))
    if drop is not None:
        y = drop(y, 0.5)
    y = _lin(sd, "sequence_head.4", y)
    y = F.relu(F.layer_norm(y, (y.shape[-1],), sd["sequence_head.5.weight"],
                            sd["sequence_head.5.bias"], 1e-5))
    if drop is not None:
        y = drop(y, 0.5)
    return _lin(sd, "sequence_head.8", y)


def _offset(sd, name, h, length):
    # models/en_gnn_decoder.py:260-290: only the first three output channels are used
    y = _lin(sd, name + ".2", F.relu(_lin(sd, name + ".0", h)))[:, :3]
    return F.normalize(y, dim=-1) * length


def backbone_from_ca(sd, h, x_ca):
    """N / C placement and the 3-iteration peptide pull (``:260-310``)."""
    x_n = x_ca + _offset(sd, "n_offset_head", h, N_CA_LENGTH)
    x_c = x_ca + _offset(sd, "c_offset_head", h, CA_C_LENGTH)
    if x_ca.shape[0] > 1:
        head = x_n[:1]
        tail = x_n[1:]
        anchor = x_c[:-1]
        for _ in range(3):                                            # :299
            vec = tail - anchor
            dist = vec.norm(dim=-1, keepdim=True)
            scale = (1.0 + 0.15 * (PEPTIDE_LENGTH / (dist + 1e-8) - 1.0)).clamp(0.90, 1.10)
            tail = anchor + vec * scale                               # :310
        x_n = torch.cat([head, tail], 0)
    return x_n, x_c


def num_layers_of(sd) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))


def egnn_decoder(sd, z_g, z_l, mask=None, max_neighbors=20, degree_normalize=True,
                 edge_cache=None, dropout=None):
    """Decoder forward in eval mode.  Reference: ``EGNNDecoder.forward`` (``:200-333``).

    Keeps the reference's per-conformer loop (F2).  ``edge_cache`` (dict) lets a
    caller reuse the banded edge list between conformers of equal length -- the
    reference rebuilds it with a Python double loop on every call.

    ``dropout=(p, keep)`` restates train mode with injected masks: ``keep(site, b, rows, dim, p_site)``
    returns the 0/1 keep mask of the ``site``-th dropout call of conformer ``b`` (call order :128, :250 per
    layer, :166, :170); a kept value is scaled by ``1 / (1 - p_site)`` as ``nn.Dropout`` does, with
    ``p_site = p`` after a layer (:114) and ``p / 2`` inside the heads (:128, :166, :170).
    """
    B, L, _ = z_l.shape
    dt = z_l.dtype
    nl = num_layers_of(sd)
    outs = ([], [], [], [])
    for b in range(B):
        if mask is not None:
            idx = torch.nonzero(mask[b].bool()).squeeze(-1)           # :217-218
        else:
            idx = torch.arange(L)
        Lb = idx.numel()
        full = [torch.zeros(L, 3, dtype=dt) for _ in range(3)] + [torch.zeros(L, 20, dtype=dt)]
        if Lb > 0:
            z = torch.cat([z_g[b].unsqueeze(0).expand(Lb, -1), z_l[b, idx]], -1)   # :233-234
            drop = None
            if dropout is not None:
                p_drop, keep_fn = dropout
                site = [0]

                def drop(y, frac, _b=b, _site=site):
                    ps = p_drop * frac
                    k = keep_fn(_site[0], _b, y.shape[0], y.shape[1], ps)
                    _site[0] += 1
                    return y * torch.as_tensor(k, dtype=y.dtype) / (1.0 - ps)
            x = _latent_to_coords(sd, z, drop)                        # :237
            h = _lin(sd, "input_embedding", z)                        # :240
            key = (Lb, max_neighbors)
            if edge_cache is not None and key in edge_cache:
                ei = edge_cache[key]
            else:
                ei = torch.from_numpy(graph_oracle.build_edge_index(Lb, max_neighbors))
                if edge_cache is not None:
                    edge_cache[key] = ei
            dinv = None
            if degree_normalize:
                deg = torch.bincount(ei[0], minlength=Lb)             # :244
                dinv = (1.0 / deg.float()).to(dt)                     # :245 (float32 reciprocal)
            for l in range(nl):                                       # :248-250
                h, x = egn_layer(sd, f"layers.{l}.", h, x, ei, dinv)
                if drop is not None:
                    h = drop(h, 1.0)                                  # :250
            logits = _sequence_head(sd, h, drop)                      # :253
            x_n, x_c = backbone_from_ca(sd, h, x)
            full[0] = full[0].index_put((idx,), x_n)                  # :313-328
            full[1] = full[1].index_put((idx,), x)
            full[2] = full[2].index_put((idx,), x_c)
            full[3] = full[3].index_put((idx,), logits)
        for o, f in zip(outs, full):
            o.append(f)
    return tuple(torch.stack(o, 0) for o in outs)                     # :330-333


def cast_state_dict(sd, dtype):
    return {k: v.detach().to(dtype).clone() for k, v in sd.items()}
