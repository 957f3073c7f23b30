#!/usr/bin/env python3
"""Output errors of the bf16 path against the float64 golden fixtures (tests/golden/decoders.npz)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import numpy as np
import torch
import cases, synth
from protein_ensemble_vae_b200 import EGNNDecoder

gold = np.load(os.path.join(ROOT, "tests", "golden", "decoders.npz"))
for tag in ("h256_gaps", "refdims"):
    case = cases.DECODER_CASES[tag]
    z_g, z_l, Hd, nl, W, B, L, mkind, pseed, dseed = case
    dec = EGNNDecoder(z_g, z_l, hidden_dim=Hd, num_layers=nl, max_neighbors=W, dropout=0.0, precision="bf16").cuda().eval()
    params = synth.make_params(synth.decoder_param_shapes(z_g, z_l, Hd, nl), pseed)
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device="cuda")
    dec.load_state_dict({k: t(v) for k, v in params.items()})
    zg, zl, mask, coef = cases.decoder_inputs(case)
    with torch.no_grad():
        outs = dec(t(zg), t(zl), None if mask is None else t(mask))
    errs = []
    for name, o in zip(("N", "CA", "C", "logits"), outs):
        g = torch.tensor(gold[f"{tag}.{name}"], device="cuda")
        errs.append(f"{name}={float((o.double() - g).abs().max() / g.abs().max()):.2e}")
    print(tag, f"layers={nl} L={L} B={B}", " ".join(errs))
