import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import check_edge3 as ce
H = 256
def run(W2, b2, tag):
    c = ce.setup((256,)*3, 40, 5)
    c["W2"] = W2; c["b2"] = b2
    g, N, E = c["g"], c["N"], c["E"]
    row, col = g.row.long(), g.col.long()
    d2, agg, w, hv, hs = ce.run3(c, True)
    k = hv.float()
    print(tag, "k[0,:8]", [round(v,4) for v in k[0,:8].tolist()], "absmax", k.abs().max().item(), "row130[:4]", [round(v,4) for v in k[130,:4].tolist()])
    return c, k
z = torch.zeros(H, device="cuda")
run(torch.zeros(H, H, device="cuda"), z, "W2=0   ")
run(2*torch.eye(H, device="cuda"), z, "W2=2I  ")
run(4*torch.eye(H, device="cuda"), z, "W2=4I  ")
W = torch.zeros(H, H, device="cuda"); W[0, 0] = 2.0
c, k = run(W, z, "W2=e00 ")
g = c["g"]; row, col = g.row.long(), g.col.long(); E = c["E"]
d2 = ((c["x"][row]-c["x"][col])**2).sum(-1)
hu = (c["ABh"][row, :H] + c["ABh"][col, H:]).float() + 0.5 * c["wd"] * d2[:, None]
a = ce.bf(ce.silu2(hu))
print("a[0,0]", a[0,0].item(), "nonzero cols in k row0:", (k[0].abs()>1e-6).nonzero().flatten().tolist()[:20])
print("k[:, 0] vs a[:,0] err", ce.rel(k[:,0], a[:,0]))
W = torch.zeros(H, H, device="cuda"); W[200, 5] = 2.0
c, k = run(W, z, "W2=e200,5")
print("nonzero cols in k row0:", (k[0].abs()>1e-6).nonzero().flatten().tolist()[:20], "k[:,200] vs a[:,5]", ce.rel(k[:,200], a[:,5]))
