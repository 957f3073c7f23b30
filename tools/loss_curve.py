#!/usr/bin/env python3
"""Loss over a few hundred Adam steps on the synthetic batch of the benchmark (is the training step numerically healthy?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import bench
from protein_ensemble_vae_b200 import EGNNDecoder
from protein_ensemble_vae_b200 import losses as pl

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
C = bench.CFG
torch.manual_seed(0)
dec = EGNNDecoder(C["z_g"], C["z_l"], hidden_dim=256, num_layers=C["layers"], max_neighbors=40, dropout=0.1, precision=prec).cuda().train()
d = bench.synth_batch(B, C["L"], C["z_g"], C["z_l"], 0, device="cuda")
tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])
step = bench.make_train_step(type("Ctx", (), {"world": 1})(), dec, bench.LOSS_W)
for i in range(301):
    loss = step(d, tdih)
    if i % 25 == 0:
        gmax = max(float(p.abs().max()) for p in dec.parameters())
        print(f"step {i:4d}: loss {float(loss):.4f}  max |param| {gmax:.3f}", flush=True)
