#!/usr/bin/env python3
"""Micro-driver for ncu: one training-mode launch of every v2 edge kernel on a B x L=256 batch (W=40)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_ensemble_vae_b200 import _lib, egnn_tc2 as T2
from protein_ensemble_vae_b200._lib import ptr, stream
from protein_ensemble_vae_b200.graph import band_graph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H, dev = 256, "cuda"
g = band_graph((256,) * B, 40, dev)
N, E = g.num_nodes, g.num_edges
torch.manual_seed(0)
x = torch.randn(N, 3, device=dev) * 3
ABh = (torch.randn(N, 2 * H, device=dev) * 0.5).to(torch.float16)
wd = torch.randn(H, device=dev) * 0.02
W2, W5 = torch.randn(H, H, device=dev) / 16, torch.randn(H, H, device=dev) / 16
b2, b5, w6 = (torch.randn(H, device=dev) * 0.1 for _ in range(3))
b6 = torch.zeros(1, device=dev)
gw = torch.randn(E, device=dev)
gagg = torch.randn(N, H, device=dev) * 0.3
L = _lib.lib()
st = stream(x)
P = lambda W, t=False: T2.packed_weight_scaled(W, 0.5, transpose=t)
W2hp, W5hp, W2thp, W5thp = P(W2), P(W5), P(W2, True), P(W5, True)
hvT, mT, ghvT = (T2.alloc_tile_image(E, dev) for _ in range(3))
hs, ghu = (torch.empty(E, H, dtype=torch.bfloat16, device=dev) for _ in range(2))
agg, w, d2, gd2 = torch.empty(N, H, device=dev), torch.empty(E, device=dev), torch.empty(E, device=dev), torch.empty(E, device=dev)
db2h, db5h, dw6 = (torch.empty(H, device=dev) for _ in range(3))
dW5, dW2 = torch.empty(H, H, device=dev), torch.empty(H, H, device=dev)
ws = torch.empty(L.cdll.pev_edge2_wgrad_workspace_bytes() // 4, device=dev)
gd2p = torch.empty(4 * E, device=dev)
gAB, part, gx = torch.empty(N, 2 * H, device=dev), torch.empty(N, H, device=dev), torch.zeros(N, 3, device=dev)
names = ["fwd1", "fwd2", "bwd2", "wgrad5", "bwd1", "wgrad2", "sums", "fwd1_infer", "fwd2_infer"]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
L.call("pev_edge_d2", ptr(x), ptr(g.row), ptr(g.col), E, ptr(d2), st)
for rep in range(reps + 1):
    ev[0].record()
    L.call("pev_edge2_fwd1", ptr(ABh), ptr(d2), ptr(wd), ptr(W2hp), ptr(b2), ptr(g.row), ptr(g.col), N, E, ptr(hvT), ptr(mT), ptr(agg), st)
    ev[1].record()
    L.call("pev_edge2_fwd2", ptr(mT), ptr(W5hp), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), ptr(hs), st)
    ev[2].record()
    L.call("pev_edge2_bwd2", ptr(hs), ptr(gw), ptr(w6), ptr(W5thp), ptr(gagg), ptr(g.row), ptr(hvT), E, ptr(ws), ptr(ghvT), ptr(db2h), st)
    ev[3].record()
    L.call("pev_edge2_wgrad5", ptr(hs), ptr(gw), ptr(w6), ptr(mT), E, ptr(ws), ptr(dW5), ptr(db5h), ptr(dw6), st)
    ev[4].record()
    L.call("pev_edge2_bwd1", ptr(ghvT), ptr(W2thp), ptr(ABh), ptr(d2), ptr(g.row), ptr(g.col), ptr(wd), E, ptr(ghu), ptr(gd2), ptr(gd2p), st)
    ev[5].record()
    L.call("pev_edge2_wgrad2", ptr(ghvT), ptr(ABh), ptr(d2), ptr(g.row), ptr(g.col), ptr(wd), E, ptr(ws), ptr(dW2), st)
    ev[6].record()
    L.call("pev_edge2_sums", ptr(ghu), ptr(d2), ptr(g.row_ptr), ptr(g.col_ptr), ptr(g.csc_perm), N, E, ptr(ws), ptr(gAB), ptr(db2h), st)
    ev[7].record()
    L.call("pev_edge2_fwd1", ptr(ABh), ptr(d2), ptr(wd), ptr(W2hp), ptr(b2), ptr(g.row), ptr(g.col), N, E, None, ptr(mT), ptr(agg), st)
    ev[8].record()
    L.call("pev_edge2_fwd2", ptr(mT), ptr(W5hp), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), None, st)
    ev[9].record()
    torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]
print(f"B={B} E={E} tiles={(E + 127) // 128} " + " ".join(f"{n}={t:.3f}" for n, t in zip(names, ms)) + f" total={sum(ms):.3f} ms")
