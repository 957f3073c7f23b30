#!/usr/bin/env python3
"""Encoder forward + backward at configs[1] shapes (B conformers of L=256, ESM-2 width 1280): ms per step and the kernel list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from protein_ensemble_vae_b200.encoder import ProteinEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
torch.manual_seed(0)
enc = ProteinEncoder(seqemb_dim=1280, precision=prec).cuda().train()
g = torch.Generator(device="cuda").manual_seed(1)
emb = torch.randn(B, L, 1280, device="cuda", generator=g)
ca = torch.cumsum(torch.randn(B, L, 3, device="cuda", generator=g) * 2.2, 1)
n, c = ca + 0.8 * torch.randn(B, L, 3, device="cuda", generator=g), ca + 0.8 * torch.randn(B, L, 3, device="cuda", generator=g)
dih = torch.randn(B, L, 6, device="cuda", generator=g).clamp(-1, 1)
mask = torch.ones(B, L, device="cuda")


def step():
    enc.zero_grad(set_to_none=True)
    z_g, z_l, mu_g, lv_g, mu_l, lv_l = enc(emb, n, ca, c, dih, mask)
    (z_g.square().mean() + z_l.square().mean() + lv_g.mean() + lv_l.mean()).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
K = 5
for _ in range(K):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
print(f"encoder {prec} B={B} L={L}: {ms:.2f} ms per fwd+bwd, {B / ms * 1e3:.0f} conformers/s, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as pr:
    step()
    torch.cuda.synchronize()
rows = sorted(pr.key_averages(), key=lambda e: -e.device_time_total)[:18]
tot = sum(e.device_time_total for e in pr.key_averages())
print(f"kernel time {tot / 1e3:.2f} ms")
for e in rows:
    print(f"  {e.device_time_total / 1e3:7.3f} ms  {e.count:4d} x  {e.key[:110]}")
