import sys, torch
sys.path.insert(0, ".")
from protein_ensemble_vae_b200.egnn_tc import node_wgrad
torch.manual_seed(0)
for N in (1, 32, 64, 1000):
    X = torch.randn(N, 256, device="cuda"); G = torch.randn(N, 256, device="cuda")
    ref = (G.double().t() @ X.double()).float()
    got = node_wgrad(G, X)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print("N", N, "err", err, "got absmax", got.abs().max().item(), "nonzero frac", (got != 0).float().mean().item())
    if N == 32:
        # structure: identity-like probes
        G = torch.zeros(N, 256, device="cuda"); X = torch.zeros(N, 256, device="cuda")
        G[3, 5] = 1.0; X[3, 7] = 2.0
        got = node_wgrad(G, X); nz = got.nonzero().tolist(); print("probe (k=3,m=5,n=7) ->", nz[:8], [got[i, j].item() for i, j in nz[:8]])
        G.zero_(); X.zero_(); G[9, 40] = 1.0; X[9, 200] = 3.0
        got = node_wgrad(G, X); nz = got.nonzero().tolist(); print("probe (k=9,m=40,n=200) ->", nz[:8], [got[i, j].item() for i, j in nz[:8]])
        G.zero_(); X.zero_(); G[20, 130] = 1.0; X[20, 33] = 3.0
        got = node_wgrad(G, X); nz = got.nonzero().tolist(); print("probe (k=20,m=130,n=33) ->", nz[:8], [got[i, j].item() for i, j in nz[:8]])
