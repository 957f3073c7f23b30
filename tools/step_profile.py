#!/usr/bin/env python3
"""Per-kernel CUDA time of one training step at bench.py's config (torch.profiler, no serialisation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import bench
from torch.profiler import ProfilerActivity, profile
from protein_ensemble_vae_b200 import EGNNDecoder, compute_total_loss
from protein_ensemble_vae_b200 import losses as pl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = bench.CFG
dev = "cuda"
torch.manual_seed(0)
dec = EGNNDecoder(C["z_g"], C["z_l"], hidden_dim=256, num_layers=C["layers"], max_neighbors=40, dropout=0.1,
                  precision="bf16").to(dev).train()
opt = torch.optim.Adam(dec.parameters(), lr=1e-4, fused=True)
d = bench.synth_batch(B, C["L"], C["z_g"], C["z_l"], 0, device=dev)
tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])


def step():
    o = dec(d["z_g"], d["z_l"], d["mask"])
    r = compute_total_loss(o[0], o[1], o[2], o[3], d["target_N"], d["target_CA"], d["target_C"], d["labels"], d["mask"],
                           d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **bench.LOSS_W)
    r["total"].backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows)
print(f"total device ms {tot:.2f}")
for k, ms, n in sorted(rows, key=lambda r: -r[1])[:40]:
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}% n={n:4d} {k[:110]}")
