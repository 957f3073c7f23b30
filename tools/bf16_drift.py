#!/usr/bin/env python3
"""How far does ANY bf16 edge MLP drift from fp32 with depth?  Compares, against the fp32 exact-order path,
(A) the fp32 path with autocast-style rounding of the three edge linears (bf16 operands AND bf16 outputs, fp32
accumulate; tests/bf16_yardstick.py) and (B) this package's tcgen05 path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from protein_ensemble_vae_b200 import EGNNDecoder
from bf16_yardstick import emulate_autocast_edge_mlp

dev = "cuda"
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())


for layers in (1, 2, 4, 6, 8):
    torch.manual_seed(5)
    d32 = EGNNDecoder(64, 32, hidden_dim=256, num_layers=layers, max_neighbors=40, dropout=0.0, precision="fp32").to(dev).eval()
    dac = EGNNDecoder(64, 32, hidden_dim=256, num_layers=layers, max_neighbors=40, dropout=0.0, precision="fp32").to(dev).eval()
    d16 = EGNNDecoder(64, 32, hidden_dim=256, num_layers=layers, max_neighbors=40, dropout=0.0, precision="bf16").to(dev).eval()
    dac.load_state_dict(d32.state_dict()); d16.load_state_dict(d32.state_dict())
    dac.precision = "autocast-emu"
    zg, zl = torch.randn(4, 64, device=dev), torch.randn(4, 256, 32, device=dev)
    with torch.no_grad(), emulate_autocast_edge_mlp():
        a, b, c = d32(zg, zl), dac(zg, zl), d16(zg, zl)
    print(f"layers={layers}: autocast-emulation vs fp32 " + " ".join(f"{rel(y, x):.2e}" for x, y in zip(a, b))
          + " | tcgen05 path vs fp32 " + " ".join(f"{rel(y, x):.2e}" for x, y in zip(a, c)) + "   (N CA C logits)")
