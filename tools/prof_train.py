#!/usr/bin/env python3
"""torch-profiler kernel list of one training step at bench.py's config (non-pev kernels first)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from protein_ensemble_vae_b200 import EGNNDecoder
from protein_ensemble_vae_b200 import losses as pl

C = bench.CFG
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
dec = EGNNDecoder(C["z_g"], C["z_l"], hidden_dim=256, num_layers=C["layers"], max_neighbors=40, dropout=0.1, precision="bf16").cuda().train()
d = bench.synth_batch(B, C["L"], C["z_g"], C["z_l"], 0, device="cuda")
tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])
ctx = type("Ctx", (), {"world": 1})()
step = bench.make_train_step(ctx, dec, bench.LOSS_W)
for _ in range(3):
    step(d, tdih)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as pr:
    step(d, tdih)
    torch.cuda.synchronize()
ev = pr.key_averages()
tot = sum(e.device_time_total for e in ev)
own = sum(e.device_time_total for e in ev if "pev::" in e.key)
print(f"kernel time {tot / 1e3:.2f} ms, own kernels {100 * own / tot:.1f} %, launches {sum(e.count for e in ev)}")
for e in sorted([e for e in ev if "pev::" not in e.key], key=lambda e: -e.device_time_total)[:22]:
    print(f"  {e.device_time_total / 1e3:7.3f} ms {100 * e.device_time_total / tot:5.2f}%  {e.count:4d} x  {e.key[:120]}")
