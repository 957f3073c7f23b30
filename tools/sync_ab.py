#!/usr/bin/env python3
"""A/B in one process: the training step launched back to back vs with a host read of the loss after every step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import bench
from protein_ensemble_vae_b200 import EGNNDecoder
from protein_ensemble_vae_b200 import losses as pl

C = bench.CFG
torch.manual_seed(0)
dec = EGNNDecoder(C["z_g"], C["z_l"], hidden_dim=256, num_layers=C["layers"], max_neighbors=40, dropout=0.1, precision="bf16").cuda().train()
d = bench.synth_batch(256, C["L"], C["z_g"], C["z_l"], 0, device="cuda")
tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])
step = bench.make_train_step(type("Ctx", (), {"world": 1})(), dec, bench.LOSS_W)


def run(sync, n=15):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        loss = step(d, tdih)
        if sync:
            float(loss)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for _ in range(5):
    step(d, tdih)
for rnd in range(3):
    print(f"round {rnd}: back-to-back {run(False):.2f} ms/step, loss read every step {run(True):.2f} ms/step", flush=True)
