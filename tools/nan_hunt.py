#!/usr/bin/env python3
"""Where does the synthetic training run lose finiteness?  Per-step loss terms and gradient norms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import bench
from protein_ensemble_vae_b200 import EGNNDecoder, compute_total_loss
from protein_ensemble_vae_b200 import losses as pl

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
C = bench.CFG
torch.manual_seed(0)
dec = EGNNDecoder(C["z_g"], C["z_l"], hidden_dim=256, num_layers=C["layers"], max_neighbors=40, dropout=0.1, precision=prec).cuda().train()
d = bench.synth_batch(B, C["L"], C["z_g"], C["z_l"], 0, device="cuda")
tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])
opt = torch.optim.Adam(dec.parameters(), lr=1e-4, fused=True)
quiet = len(sys.argv) > 3
for i in range(40):
    seen = {}
    hooks = []
    def watch(name):
        def fwd_hook(mod, inp, out):
            if torch.is_tensor(out) and out.requires_grad:
                out.register_hook(lambda g, name=name: seen.__setitem__(name, bool(torch.isfinite(g).all())))
        return fwd_hook
    for name, mod in (("n_offset_head", dec.n_offset_head), ("c_offset_head", dec.c_offset_head), ("sequence_head", dec.sequence_head),
                      ("latent_to_coords", dec.latent_to_coords), ("layer5", dec.layers[5]), ("layer0", dec.layers[0])):
        hooks.append(mod.register_forward_hook(watch(name)))
    outs = dec(d["z_g"], d["z_l"], d["mask"])
    for nm, o in zip(("N", "CA", "C", "logits"), outs):
        o.register_hook(lambda g, nm=nm: seen.__setitem__("out_" + nm, bool(torch.isfinite(g).all())))
    res = compute_total_loss(outs[0], outs[1], outs[2], outs[3], d["target_N"], d["target_CA"], d["target_C"], d["labels"], d["mask"],
                             d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **bench.LOSS_W)
    res["total"].backward()
    bad = [n for n, p in dec.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in dec.parameters() if p.grad is not None)))
    terms = {k: round(float(v), 3) for k, v in res.items() if k in ("total", "reconstruction", "pair_distance", "bond_length", "bond_angle", "clash", "ramachandran", "sequence", "dihedral_total")}
    if not quiet or bad:
      print(i, terms, "grad norm %.3e" % gn, "coords max %.1f" % float(outs[1].abs().max()), "non-finite grads:", len(bad), bad if len(bad) < 40 else bad[:40], flush=True)
    for h in hooks:
        h.remove()
    if bad or not torch.isfinite(res["total"]):
        print("finite upstream gradients:", seen)
        break
    opt.step()
    opt.zero_grad(set_to_none=True)
