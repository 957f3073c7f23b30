"""bf16 tensor-core (tcgen05) path of one EGNN layer -- filled in by the K1 tcgen05 kernels."""
from __future__ import annotations

import torch.nn as nn


def supports(layer) -> bool:
    return (layer.node_dim == 256 and layer.hidden_dim == 256
            and all(isinstance(layer.phi_e[i], nn.SiLU) for i in (1, 3)) and isinstance(layer.phi_x[1], nn.SiLU))


def egn_layer_bf16(layer, h, x, g, dinv):
    raise NotImplementedError
