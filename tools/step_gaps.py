#!/usr/bin/env python3
"""GPU busy time vs wall span of training steps at bench.py's config: union of the kernel intervals that torch.profiler
records (no serialisation), the largest idle gaps and the CPU op each gap follows."""
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
NS = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
span = iv[-1][1] - iv[0][0]
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
last_name = iv[0][2]
for s, e, n in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, last_name, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
    last_name = n
busy += cur_e - cur_s
print(f"{NS} steps: span {span / 1e3 / NS:.2f} ms/step, GPU busy {busy / 1e3 / NS:.2f} ms/step, idle {(span - busy) / 1e3 / NS:.2f} ms/step "
      f"in {len(gaps) / NS:.0f} gaps/step")
gaps.sort(reverse=True)
for g, a, b in gaps[:12]:
    print(f"  {g:8.1f} us after {a[:60]:60s} before {b[:60]}")
import collections
hist = collections.Counter()
for g, a, b in gaps:
    hist[min(int(g // 2) * 2, 20)] += g
print("idle by gap size (us bucket -> total us/step):", {k: round(v / NS, 1) for k, v in sorted(hist.items())})
