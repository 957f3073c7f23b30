#!/usr/bin/env python3
"""Where do the encoder step's aten::copy_ / contiguous calls come from (python stacks, sizes)?"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from protein_ensemble_vae_b200.encoder import ProteinEncoder

B, L = 64, 256
torch.manual_seed(0)
enc = ProteinEncoder(seqemb_dim=1280).cuda().train()
g = torch.Generator(device="cuda").manual_seed(1)
emb = torch.randn(B, L, 1280, device="cuda", generator=g)
ca = torch.cumsum(torch.randn(B, L, 3, device="cuda", generator=g) * 2.2, 1)
dih = torch.randn(B, L, 6, device="cuda", generator=g).clamp(-1, 1)
mask = torch.ones(B, L, device="cuda")


def step():
    enc.zero_grad(set_to_none=True)
    out = enc(emb, ca, ca, ca, dih, mask)
    (out[0].square().mean() + out[1].square().mean() + out[3].mean() + out[5].mean()).backward()


step(); step()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as pr:
    step()
    torch.cuda.synchronize()
cnt = collections.Counter()
for e in pr.events():
    if e.name in ("aten::copy_", "aten::contiguous", "aten::clone") and e.input_shapes and e.input_shapes[0] and len(e.input_shapes[0]) == 2 \
            and e.input_shapes[0][0] >= B * L:
        st = [s for s in (e.stack or []) if "protein_ensemble_vae_b200" in s or "tools/" in s][:2]
        cnt[(e.name, str(e.input_shapes[:2]), " <- ".join(s.split("/")[-1] for s in st))] += 1
for k, v in cnt.most_common(25):
    print(v, k)
