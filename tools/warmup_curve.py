#!/usr/bin/env python3
"""Per-step time of the first 45 training steps, with the caching allocator's segment count and the SM clock."""
import os, sys, subprocess
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
evs = []
for i in range(45):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step(d, tdih)
    b.record()
    evs.append((a, b, torch.cuda.memory_stats()["segment.all.current"], torch.cuda.memory_reserved() / 2**30))
torch.cuda.synchronize()
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print("after:", clk)
print(" ".join(f"{a.elapsed_time(b):.1f}" for a, b, _, _ in evs))
print("segments", [s for _, _, s, _ in evs][::5], "reserved GiB", [round(r, 1) for _, _, _, r in evs][::5])
