#!/usr/bin/env python3
"""Micro-driver: one launch of each tcgen05 edge kernel (stages 1-4) on a B x L=256 batch, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_ensemble_vae_b200 import _lib
from protein_ensemble_vae_b200._lib import ptr, stream
from protein_ensemble_vae_b200.egnn_tc import packed_weight
from protein_ensemble_vae_b200.graph import band_graph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H = 256
g = band_graph((256,) * B, 40, "cuda")
N, E = g.num_nodes, g.num_edges
torch.manual_seed(0)
AB = torch.randn(N, 2 * H, device="cuda")
x = torch.randn(N, 3, device="cuda")
wd = torch.randn(H, device="cuda") * 0.05
W2, W5 = torch.randn(H, H, device="cuda") / 16, torch.randn(H, H, device="cuda") / 16
b2, b5, w6 = (torch.randn(H, device="cuda") * 0.1 for _ in range(3))
b6 = torch.zeros(1, device="cuda")
bf = torch.bfloat16
v, a, da, m, dm, s, gs, gv, gu = (torch.empty(E, H, dtype=bf, device="cuda") for _ in range(9))
agg = torch.empty(N, H, device="cuda")
w = torch.empty(E, device="cuda")
gw = torch.randn(E, device="cuda")
gagg = torch.randn(N, H, device="cuda")
db5, dw6, db2 = (torch.empty(H, device="cuda") for _ in range(3))
gd2 = torch.empty(E, device="cuda")
L = _lib.lib()
st = stream(x)
W2p, W5p, W2t, W5t = packed_weight(W2), packed_weight(W5), packed_weight(W2, True), packed_weight(W5, True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for rep in range(reps + 1):
    ev[0].record()
    L.call("pev_edge_mlp1_fwd_bf16", ptr(AB), ptr(x), ptr(wd), ptr(W2p), ptr(b2), ptr(g.row), ptr(g.col), N, E, ptr(v), ptr(a), ptr(da), ptr(agg), st)
    ev[1].record()
    L.call("pev_edge_mlp2_fwd_bf16", ptr(v), ptr(W5p), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), ptr(s), ptr(m), ptr(dm), st)
    ev[2].record()
    L.call("pev_edge_mlp2_bwd_bf16", ptr(s), ptr(dm), ptr(gw), ptr(w6), ptr(W5t), ptr(gagg), ptr(g.row), E, ptr(gs), ptr(gv), ptr(db5), ptr(dw6), st)
    ev[3].record()
    L.call("pev_edge_mlp1_bwd_bf16", ptr(gv), ptr(da), ptr(W2t), ptr(wd), E, ptr(gu), ptr(gd2), ptr(db2), st)
    ev[4].record()
    torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
fl = 2.0 * E * H * H / 1e9
print(f"B={B} E={E} tiles={E // 128} ms/stage={['%.3f' % t for t in ms]} TFLOP/s={['%.0f' % (fl / t) for t in ms]}")
