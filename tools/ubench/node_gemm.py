#!/usr/bin/env python3
"""Timing of the node-level weight-gradient GEMM forms (TF32 / bf16) at config 2 (N = 65536)."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True
N = 65536
dev = "cuda"
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for Do, Di in ((256, 256), (512, 256), (256, 512)):
    g = torch.randn(N, Do, device=dev); x = torch.randn(N, Di, device=dev)
    W = torch.randn(Do, Di, device=dev)
    gb, xb = g.bfloat16(), x.bfloat16()
    print(f"out={Do} in={Di}:",
          f"g.t()@x {t(lambda: g.t() @ x):.3f}",
          f"(x.t()@g).t() {t(lambda: (x.t() @ g).t()):.3f}",
          f"gt_contig@x {t(lambda: g.t().contiguous() @ x):.3f}",
          f"bf16 incl casts {t(lambda: torch.mm(g.bfloat16().t(), x.bfloat16(), out_dtype=torch.float32)):.3f}",
          f"bf16 gemm only {t(lambda: torch.mm(gb.t(), xb, out_dtype=torch.float32)):.3f}",
          f"dgrad g@W {t(lambda: g @ W):.3f}",
          f"fwd x@W.t() {t(lambda: x @ W.t()):.3f} ms")
