#!/usr/bin/env python3
"""GPU check + timing of the v2 backward edge kernels against plain torch fp32 on the same (bf16-rounded) inputs.

usage: check_edge2_bwd.py [B] [reps] [L] [W]
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
bf16 = torch.bfloat16
x = torch.randn(N, 3, device=dev) * 3
ABh = (torch.randn(N, 2 * H, device=dev) * 0.5).to(torch.float16)
wd = torch.randn(H, device=dev) * 0.02
W2, W5 = torch.randn(H, H, device=dev) / 16, torch.randn(H, H, device=dev) / 16
w6 = torch.randn(H, device=dev) * 0.1
hs = (torch.randn(E, H, device=dev) * 0.7).to(bf16)
hv = (torch.randn(E, H, device=dev) * 0.7).to(bf16)
gw = torch.randn(E, device=dev)
gagg = torch.randn(N, H, device=dev) * 0.3
row, col = g.row.long(), g.col.long()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def bf(t):
    return t.to(bf16).float()


def one_plus_r(h):
    t = torch.tanh(h)
    return 1 + t + h * (1 - t * t)


# ---------------------------------------------------------------- reference
d2 = ((x[row] - x[col]) ** 2).sum(-1)
ghs = gw[:, None] * w6 * one_plus_r(hs.float())
gm = bf(ghs) @ bf(0.5 * W5)                       # [E,256]: gm[e,f] = sum_k ghs[e,k] W5h[k,f]
ghv = (gm + gagg[row]) * one_plus_r(hv.float())
db2h = ghv.sum(0)
ga = bf(ghv) @ bf(0.5 * W2)
hu = (ABh[row, :H] + ABh[col, H:]).float() + 0.5 * wd * d2[:, None]
ghu = ga * one_plus_r(hu)
gd2 = ghu @ (0.5 * wd)

L = _lib.lib()
st = stream(x)
W5thp = T2.packed_weight_scaled(W5, 0.5, transpose=True)
W2thp = T2.packed_weight_scaled(W2, 0.5, transpose=True)
hvT = T2.rows_to_tile_image(hv)
mT = T2.rows_to_tile_image((hv.float() + hv.float() * torch.tanh(hv.float())).to(bf16))
ghvT = T2.alloc_tile_image(E, dev)
db2h_k = torch.empty(H, device=dev)
ghu_k = torch.empty(E, H, dtype=bf16, device=dev)
gd2_k = torch.empty(E, device=dev)
d2k = torch.empty(E, device=dev)
L.call("pev_edge_d2", ptr(x), ptr(g.row), ptr(g.col), E, ptr(d2k), st)
ws = torch.empty(L.cdll.pev_edge2_wgrad_workspace_bytes() // 4, device=dev)
dW5_k, dW2_k = torch.empty(H, H, device=dev), torch.empty(H, H, device=dev)
db5h_k, dw6_k = torch.empty(H, device=dev), torch.empty(H, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for rep in range(reps + 1):
    ev[0].record()
    L.call("pev_edge2_bwd2", ptr(hs), ptr(gw), ptr(w6), ptr(W5thp), ptr(gagg), ptr(g.row), ptr(hvT), E, ptr(ws), ptr(ghvT),
           ptr(db2h_k), st)
    ev[1].record()
    L.call("pev_edge2_bwd1", ptr(ghvT), ptr(W2thp), ptr(ABh), ptr(d2k), ptr(g.row), ptr(g.col), ptr(wd), E, ptr(ghu_k),
           ptr(gd2_k), ptr(torch.empty(4 * E, device=dev)), st)
    ev[2].record()
    L.call("pev_edge2_wgrad5", ptr(hs), ptr(gw), ptr(w6), ptr(mT), E, ptr(ws), ptr(dW5_k), ptr(db5h_k), ptr(dw6_k), st)
    ev[3].record()
    L.call("pev_edge2_wgrad2", ptr(ghvT), ptr(ABh), ptr(d2k), ptr(g.row), ptr(g.col), ptr(wd), E, ptr(ws), ptr(dW2_k), st)
    ev[4].record()
    torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
print(f"B={B} L={Lr} W={Wn} N={N} E={E} ms bwd2={ms[0]:.3f} bwd1={ms[1]:.3f} wgrad5={ms[2]:.3f} wgrad2={ms[3]:.3f}")
ghv_k = T2.tile_image_to_rows(ghvT, E).float()
print(f"ghv  rel err {rel(ghv_k, ghv):.3e}")
print(f"db2h rel err {rel(db2h_k, db2h):.3e}")
# bwd1 consumed the kernel's own (bf16) ghv: rebuild the reference from it for a tight comparison
ga2 = ghv_k @ bf(0.5 * W2)
ghu2 = ga2 * one_plus_r(hu)
print(f"ghu  rel err {rel(ghu_k.float(), ghu2):.3e}   (vs end-to-end reference {rel(ghu_k.float(), ghu):.3e})")
print(f"gd2  rel err {rel(gd2_k, ghu2 @ (0.5 * wd)):.3e}")
# weight gradients
hsf, hvf = hs.float(), hv.float()
m = hvf + hvf * torch.tanh(hvf)
tt = hsf + hsf * torch.tanh(hsf)
dW5 = 0.5 * bf(ghs).t().double() @ bf(m).double()
print(f"dW5  rel err {rel(dW5_k, dW5):.3e}")
print(f"db5h rel err {rel(db5h_k, ghs.double().sum(0)):.3e}")
print(f"dw6  rel err {rel(dw6_k, (gw[:, None].double() * tt.double()).sum(0)):.3e}")
a = hu + hu * torch.tanh(hu)
dW2 = 0.5 * ghv_k.t().double() @ bf(a).double()
print(f"dW2  rel err {rel(dW2_k, dW2):.3e}")
