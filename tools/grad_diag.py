#!/usr/bin/env python3
"""bf16-path vs fp32-path gradients on the GPU (per-parameter relative L2 error)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch, cases, synth
from protein_ensemble_vae_b200 import EGNNDecoder
tag = sys.argv[1] if len(sys.argv) > 1 else "h256_gaps"
case = cases.DECODER_CASES[tag]
z_g, z_l, Hd, nl, W, B, L, mkind, pseed, dseed = case
params = synth.make_params(synth.decoder_param_shapes(z_g, z_l, Hd, nl), pseed)
t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device="cuda")
zg, zl, mask, coef = cases.decoder_inputs(case)
res = {}
for prec in ("fp32", "bf16"):
    dec = EGNNDecoder(z_g, z_l, hidden_dim=Hd, num_layers=nl, max_neighbors=W, dropout=0.0, precision=prec).cuda().eval()
    dec.load_state_dict({k: t(v) for k, v in params.items()})
    zl_t = t(zl).requires_grad_()
    outs = dec(t(zg), zl_t, None if mask is None else t(mask))
    sum((o * t(c)).sum() for o, c in zip(outs, coef)).backward()
    res[prec] = {k: p.grad.double() for k, p in dec.named_parameters() if p.grad is not None}
    res[prec]["z_l"] = zl_t.grad.double()
    res[prec]["_out"] = [o.detach().double() for o in outs]
for k in sorted(res["fp32"]):
    if k == "_out": continue
    a, b = res["bf16"][k], res["fp32"][k]
    print(f"{k:40s} relL2={float((a-b).norm()/b.norm()):.4f} maxrel={float((a-b).abs().max()/b.abs().max()):.4f}")
for i, n in enumerate(("N", "CA", "C", "logits")):
    a, b = res["bf16"]["_out"][i], res["fp32"]["_out"][i]
    print("out", n, float((a - b).abs().max() / b.abs().max()))
