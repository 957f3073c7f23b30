#!/usr/bin/env python3
"""Node-level tcgen05 kernels (pev_node_gemm / pev_node_wgrad) against the library TF32 GEMMs they replace, N = 65 536."""
import sys
import torch
sys.path.insert(0, ".")
from protein_ensemble_vae_b200 import egnn_tc as T  # noqa: E402

N = 65536
torch.manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda")  # noqa: E731
h, agg, gr, gp, gAB = r(N, 256), r(N, 256), r(N, 256), r(N, 256), r(N, 512)
Wcat, b1, W3, b3, W4, b4 = r(512, 256) / 16, r(512), r(256, 512) / 22, r(256), r(256, 256) / 16, r(256)
gamma, beta = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
W4t, W3t, Wct = W4.t().contiguous(), W3.t().contiguous(), Wcat.t().contiguous()
torch.backends.cuda.matmul.allow_tf32 = True


def t_ms(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


q = T.node_gemm(T.EPI_SILU, h, W3, b3, A2=agg)[0]
rows = [
    ("ABH  gemm+fp16", lambda: T.node_gemm(T.EPI_ABH, h, Wcat, b1, scale=0.5), lambda: torch.addmm(b1, h, Wcat.t()).half(), 201),
    ("SILU two-operand", lambda: T.node_gemm(T.EPI_SILU, h, W3, b3, A2=agg, out2=True),
     lambda: torch.nn.functional.silu(torch.addmm(b3, h, W3[:, :256].t()).addmm_(agg, W3[:, 256:].t())), 268),
    ("RES_LN", lambda: T.node_gemm(T.EPI_RES_LN, q, W4, b4, aux=h, gamma=gamma, beta=beta, out2=True, stats=True),
     lambda: torch.nn.functional.layer_norm(h + torch.addmm(b4, q, W4.t()), (256,), gamma, beta), 268),
    ("DSILU dgrad", lambda: T.node_gemm(T.EPI_DSILU, gr, W4t, aux=h), lambda: (gr @ W4) * torch.sigmoid(h), 201),
    ("PLAIN dgrad N=512", lambda: T.node_gemm(T.EPI_PLAIN, gp, W3t), lambda: gp @ W3, 201),
    ("PLAIN dgrad K=512", lambda: T.node_gemm(T.EPI_PLAIN, gAB, Wct), lambda: gAB @ Wcat, 201),
    ("wgrad 256x256", lambda: T.node_wgrad(gr, q), lambda: gr.t() @ q, 134),
    ("wgrad 512x256", lambda: T.node_wgrad(gAB, h, 0.5), lambda: 0.5 * (gAB.t() @ h), 201),
]
for name, mine, lib, mb in rows:
    a, b = t_ms(mine), t_ms(lib)
    print(f"{name:20s} tcgen05 {a * 1e3:7.1f} us ({mb / a / 1e3:5.2f} TB/s of {mb} MB)   library {b * 1e3:7.1f} us", flush=True)
