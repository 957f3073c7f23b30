#!/usr/bin/env python3
"""GPU check + timing of the v2 edge kernels (csrc/edge_tc2_kernels.cu) against plain torch fp32 on the same inputs.

usage: check_edge2.py [B] [reps] [L] [W]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from protein_ensemble_vae_b200 import _lib, egnn_tc2 as T2
from protein_ensemble_vae_b200._lib import ptr, stream
from protein_ensemble_vae_b200.graph import band_graph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Lr = int(sys.argv[3]) if len(sys.argv) > 3 else 256
Wn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
H = 256
dev = "cuda"
g = band_graph((Lr,) * B, Wn, dev)
N, E = g.num_nodes, g.num_edges
torch.manual_seed(0)
h = torch.randn(N, H, device=dev)
x = torch.randn(N, 3, device=dev) * 3
W1 = torch.randn(H, 2 * H + 1, device=dev) / 22
b1 = torch.randn(H, device=dev) * 0.1
W2, W5 = torch.randn(H, H, device=dev) / 16, torch.randn(H, H, device=dev) / 16
b2, b5, w6 = (torch.randn(H, device=dev) * 0.1 for _ in range(3))
b6 = torch.randn(1, device=dev)
wd = W1[:, 2 * H].contiguous()
Wcat = torch.cat([W1[:, :H], W1[:, H:2 * H]], 0)
ABh = (0.5 * (h @ Wcat.t() + torch.cat([b1, torch.zeros_like(b1)]))).to(torch.float16).contiguous()
row, col = g.row.long(), g.col.long()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


# ---------------------------------------------------------------- reference (fp32, bf16-rounded operands like the kernel)
def bf(t):
    return t.to(torch.bfloat16).float()


d2 = ((x[row] - x[col]) ** 2).sum(-1, keepdim=True)
hu = (ABh[row, :H] + ABh[col, H:]).float() + 0.5 * wd * d2
a = hu + hu * torch.tanh(hu)
hv = bf(a) @ bf(0.5 * W2).t() + 0.5 * b2
m = hv + hv * torch.tanh(hv)
agg_ref = torch.zeros(N, H, device=dev).index_add_(0, row, m)
hs = bf(m) @ bf(0.5 * W5).t() + 0.5 * b5
t = hs + hs * torch.tanh(hs)
w_ref = t @ w6 + b6

L = _lib.lib()
st = stream(x)
W2hp, W5hp = T2.packed_weight_scaled(W2, 0.5), T2.packed_weight_scaled(W5, 0.5)
hvT, mT = T2.alloc_tile_image(E, dev), T2.alloc_tile_image(E, dev)
agg = torch.empty(N, H, device=dev)
w = torch.empty(E, device=dev)
hs_out = torch.empty(E, H, dtype=torch.bfloat16, device=dev)
d2k = torch.empty(E, device=dev)
L.call("pev_edge_d2", ptr(x), ptr(g.row), ptr(g.col), E, ptr(d2k), st)
print(f"d2  rel err {rel(d2k, d2.squeeze(-1)):.3e}")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(reps + 1):
    ev[0].record()
    L.call("pev_edge2_fwd1", ptr(ABh), ptr(d2k), ptr(wd), ptr(W2hp), ptr(b2), ptr(g.row), ptr(g.col), N, E, ptr(hvT),
           ptr(mT), ptr(agg), st)
    ev[1].record()
    L.call("pev_edge2_fwd2", ptr(mT), ptr(W5hp), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), ptr(hs_out), st)
    ev[2].record()
    L.call("pev_edge2_fwd2", ptr(mT), ptr(W5hp), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), None, st)
    ev[3].record()
    torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
hv_k = T2.tile_image_to_rows(hvT, E).float()
print(f"B={B} L={Lr} W={Wn} N={N} E={E} ms fwd1={ms[0]:.3f} fwd2(train)={ms[1]:.3f} fwd2(infer)={ms[2]:.3f}")
print(f"hv  rel err {rel(hv_k, hv):.3e}")
print(f"m   rel err {rel(T2.tile_image_to_rows(mT, E).float(), m):.3e}")
print(f"agg rel err {rel(agg, agg_ref):.3e}")
print(f"hs  rel err {rel(hs_out.float(), hs):.3e}")
print(f"w   rel err {rel(w, w_ref):.3e}")
