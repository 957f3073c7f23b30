#!/usr/bin/env python3
"""Kernel breakdown of the decode path (configs[3]: 8 layers, L=100, chunks of 2048 samples, + Kabsch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from protein_ensemble_vae_b200 import EGNNDecoder
from protein_ensemble_vae_b200 import distributed as pdist

S, L = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 100
torch.manual_seed(1)
dec = EGNNDecoder(512, 256, hidden_dim=256, num_layers=8, max_neighbors=40, dropout=0.1, precision="bf16").cuda().eval()
g = torch.Generator(device="cuda").manual_seed(1)
zg = torch.randn(S, 512, device="cuda", generator=g)
zl = torch.randn(S, L, 256, device="cuda", generator=g)
mask1 = torch.ones(L, device="cuda")
ref = torch.cumsum(torch.randn(L, 3, device="cuda", generator=g) * 2.2, 0)
run = lambda: pdist.decode_ensemble(dec, zg, zl, mask1, ref, chunk=2048)  # noqa: E731
run(); run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
print(f"{S} samples: {a.elapsed_time(b):.2f} ms -> {S / a.elapsed_time(b) * 1e3:.0f} conformers/s")
with profile(activities=[ProfilerActivity.CUDA]) as pr:
    run()
    torch.cuda.synchronize()
tot = sum(e.device_time_total for e in pr.key_averages())
print(f"kernel time {tot / 1e3:.2f} ms")
for e in sorted(pr.key_averages(), key=lambda e: -e.device_time_total)[:16]:
    print(f"  {e.device_time_total / 1e3:7.3f} ms {100 * e.device_time_total / tot:5.1f}%  {e.count:4d} x  {e.key[:100]}")
