#!/usr/bin/env python3
"""Eager vs CUDA-graph-replayed training step at bench.py's config (ms per step, CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import bench
from protein_ensemble_vae_b200 import EGNNDecoder, GraphedStep, compute_total_loss
from protein_ensemble_vae_b200 import losses as pl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = bench.CFG
dev = "cuda"
torch.manual_seed(0)
dec = EGNNDecoder(C["z_g"], C["z_l"], hidden_dim=256, num_layers=C["layers"], max_neighbors=40, dropout=0.1,
                  precision="bf16").to(dev).train()
params = list(dec.parameters())
opt = torch.optim.Adam(params, lr=1e-4, fused=True)
d = bench.synth_batch(B, C["L"], C["z_g"], C["z_l"], 0, device=dev)
tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])


def fb(d):
    o = dec(d["z_g"], d["z_l"], d["mask"])
    r = compute_total_loss(o[0], o[1], o[2], o[3], d["target_N"], d["target_CA"], d["target_C"], d["labels"], d["mask"],
                           d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **bench.LOSS_W)
    r["total"].backward()
    return r["total"].detach()


def timeit(fn, n=20):
    for _ in range(40):           # past the post-idle power transient (profiles/r02_sustained_step.md)
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def eager():
    fb(d)
    opt.step()
    opt.zero_grad(set_to_none=True)


t_e = timeit(eager)
torch.cuda.empty_cache()
g = GraphedStep(fb, d, params)


def graphed():
    g(d)
    opt.step()


t_g = timeit(graphed)
print(f"eager {t_e:.2f} ms/step, graphed {t_g:.2f} ms/step, loss {float(g.loss):.4f}, "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
