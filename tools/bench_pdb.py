#!/usr/bin/env python3
"""Streaming PDB writer: S models of L residues from device tensors to a file (bytes, seconds, MB/s), and the same text
produced by a Python loop over the first models for scale."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_ensemble_vae_b200 import write_ensemble_pdb

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L = 100
g = torch.Generator(device="cuda").manual_seed(0)
ca = torch.cumsum(torch.randn(S, L, 3, device="cuda", generator=g) * 2.2, 1)
n, c = ca + 0.8 * torch.randn(S, L, 3, device="cuda", generator=g), ca + 0.8 * torch.randn(S, L, 3, device="cuda", generator=g)
mask = torch.ones(L, device="cuda")
path = "/tmp/pev_bench.pdb"
write_ensemble_pdb(path, n[:64], ca[:64], c[:64], mask)
torch.cuda.synchronize()
t0 = time.perf_counter()
nb = write_ensemble_pdb(path, n, ca, c, mask, sequence="ACDEFGHIKL" * 10)
dt = time.perf_counter() - t0
print(f"{S} models x L={L}: {nb / 1e6:.0f} MB in {dt:.2f} s = {nb / dt / 1e6:.0f} MB/s, {S / dt:.0f} models/s")
os.remove(path)
