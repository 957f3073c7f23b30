#!/usr/bin/env python3
"""Kernel breakdown and GPU idle time of bench.py's decode step (L=100, 8 layers, 2048 samples + Kabsch)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import bench
from torch.profiler import ProfilerActivity, profile
from protein_ensemble_vae_b200 import EGNNDecoder, kabsch_rmsd_batch

dd = bench.DECODE
dev = "cuda"
torch.manual_seed(0)
dec8 = EGNNDecoder(dd["z_g"], dd["z_l"], hidden_dim=256, num_layers=dd["layers"], max_neighbors=40, dropout=0.1,
                   precision="bf16").to(dev).eval()
S = dd["chunk"]
zg = torch.randn(S, dd["z_g"], device=dev)
zl = torch.randn(S, dd["L"], dd["z_l"], device=dev)
mask1 = torch.ones(S, dd["L"], device=dev)
ref_ca = torch.randn(dd["L"], 3, device=dev)


def step():
    with torch.no_grad():
        n, ca, c, lg = dec8(zg, zl, mask1)
        return kabsch_rmsd_batch(ca, ref_ca)


for _ in range(3):
    step()
torch.cuda.synchronize()
NS = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
span = iv[-1][1] - iv[0][0]
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e, n in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"{NS} decode steps: span {span / 1e3 / NS:.2f} ms/step, GPU busy {busy / 1e3 / NS:.2f} ms/step")
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in iv:
    agg[n][0] += 1
    agg[n][1] += (e - s) / 1e3
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{ms / NS:8.3f} ms/step n={c / NS:6.1f} {n[:120]}")
