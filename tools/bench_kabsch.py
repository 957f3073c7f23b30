"""K4 micro-benchmark: batched Kabsch RMSD of S conformers (L residues) against one reference structure, and the all-pairs
matrix of a 512-member ensemble.  PEV_KABSCH_GROUP=g forces g conformers per warp (default: chosen from the problem size)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protein_ensemble_vae_b200 import kabsch_rmsd_batch, kabsch_rmsd_pairs  # noqa: E402


def t_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    S, L = 100000, 100
    g = torch.Generator(device="cuda").manual_seed(0)
    ref = torch.cumsum(torch.randn(L, 3, device="cuda", generator=g) * 2.2, 0)
    x = ref[None] + 0.5 * torch.randn(S, L, 3, device="cuda", generator=g)
    mask = torch.ones(L, device="cuda")
    r = kabsch_rmsd_batch(x, ref, mask)
    ms = t_ms(lambda: kabsch_rmsd_batch(x, ref, mask))
    gb = S * L * 12 / 1e9
    ens = x[:512].contiguous()
    p = kabsch_rmsd_pairs(ens, mask)
    ms_p = t_ms(lambda: kabsch_rmsd_pairs(ens, mask))
    print({"group": os.environ.get("PEV_KABSCH_GROUP", "auto"), "batch_ms": round(ms, 4), "batch_GBps": round(gb / ms * 1e3, 1),
           "pairs512_ms": round(ms_p, 4), "checksum": float(r.double().sum()), "pairs_checksum": float(p.double().sum())})


if __name__ == "__main__":
    main()
