"""Yardstick for 'bf16 edge MLP' parity: the fp32 exact-order decoder with the edge MLP evaluated the way
``torch.autocast(dtype=bfloat16)`` would evaluate the reference's ``phi_e`` / ``phi_x`` linears
(``models/en_gnn_decoder.py:34-50``): bf16 operands and bf16 outputs, fp32 accumulation, fp32 everywhere else.

Test infrastructure only (used by tests/test_gpu_configs.py and tools/bf16_drift.py).
"""
import torch
import torch.nn.functional as F

from protein_ensemble_vae_b200 import egnn_ops, en_gnn_decoder

BF = torch.bfloat16


def _r(t):
    return t.to(BF).float()


def autocast_layer(layer, h, x, g, dinv):
    """One EGNN layer in plain torch with autocast-style rounding of the three 256-wide edge linears."""
    D = layer.node_dim
    row, col = g.row.long(), g.col.long()
    W1, b1 = layer.phi_e[0].weight, layer.phi_e[0].bias
    rel = x[row] - x[col]
    d2 = (rel * rel).sum(-1, keepdim=True)
    hb = _r(h)
    A = hb @ _r(W1[:, :D]).t()
    Bm = hb @ _r(W1[:, D:2 * D]).t()
    u = _r(A[row] + Bm[col] + _r(d2) * _r(W1[:, 2 * D]) + b1)                 # Linear(513 -> 256) under autocast
    a = F.silu(u)
    v = _r(F.linear(_r(a), _r(layer.phi_e[2].weight), layer.phi_e[2].bias))   # Linear(256 -> 256)
    m = F.silu(v)
    s = _r(F.linear(_r(m), _r(layer.phi_x[0].weight), layer.phi_x[0].bias))   # Linear(256 -> 256)
    w = F.linear(F.silu(s), layer.phi_x[2].weight, layer.phi_x[2].bias).squeeze(-1)
    agg, x_new = egnn_ops.ScatterCoord.apply(m, w, x, dinv, g)
    h_new = layer.norm_h(h + layer.phi_h(torch.cat([h, agg], -1)))
    return h_new, x_new


class emulate_autocast_edge_mlp:
    """Context manager: decoders constructed with ``precision="autocast-emu"`` run :func:`autocast_layer`."""

    def __enter__(self):
        self.prev = en_gnn_decoder._layer_forward

        def fwd(layer, h, x, g, dinv, precision):
            if precision == "autocast-emu":
                return autocast_layer(layer, h, x, g, dinv)
            return self.prev(layer, h, x, g, dinv, precision)

        en_gnn_decoder._layer_forward = fwd
        return self

    def __exit__(self, *exc):
        en_gnn_decoder._layer_forward = self.prev
        return False
