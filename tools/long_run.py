#!/usr/bin/env python3
"""Does the 50 ms step hold under sustained load?  600 training steps (~30 s), mean step time per block of 50 steps, SM clock,
power and temperature at the end of each block."""
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
for blk in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        step(d, tdih)
    b.record()
    torch.cuda.synchronize()
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"],
                       capture_output=True, text=True).stdout.strip()
    print(f"steps {50 * blk:4d}-{50 * blk + 49:4d}: {a.elapsed_time(b) / 50:.2f} ms/step   [{q}]", flush=True)
