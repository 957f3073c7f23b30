"""Node-level tcgen05 TF32 GEMMs with fused epilogues (csrc/node_gemm_kernels.cu, ``pev_node_gemm``) against torch float64
on the same operands; tolerance 2e-3 of max|ref| = TF32's 10-bit mantissa over K = 256 / 512 (the same arithmetic the
cuBLAS allow_tf32 path had), through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
ABH, SILU, RES_LN, PLAIN, DSILU = range(5)


def gemm(epi, A1, W, bias=None, A2=None, Nout=None, scale=1.0, aux=None, gamma=None, beta=None, eps=1e-5, train=True):
    from protein_ensemble_vae_b200 import _lib
    from protein_ensemble_vae_b200._lib import ptr, stream
    M, K1 = A1.shape
    K2 = 0 if A2 is None else A2.shape[1]
    Nout = Nout or W.shape[0]
    dev = A1.device
    out = torch.full((M, Nout), 7.0, dtype=torch.float16 if epi == ABH else torch.float32, device=dev)
    out2 = torch.full((M, Nout), 7.0, device=dev) if (train and epi in (SILU, RES_LN)) else None
    mean = torch.empty(M, device=dev) if (train and epi == RES_LN) else None
    rstd = torch.empty(M, device=dev) if (train and epi == RES_LN) else None
    _lib.lib().call("pev_node_gemm", epi, ptr(A1), K1, ptr(A2), K2, ptr(W), ptr(bias), M, Nout, float(scale), ptr(aux),
                    ptr(gamma), ptr(beta), float(eps), ptr(out), ptr(out2), ptr(mean), ptr(rstd), stream(A1))
    return out, out2, mean, rstd


@pytest.mark.parametrize("M", [1, 127, 128, 1000, 20000])
def test_node_gemm_epilogues(M):
    torch.manual_seed(M)
    r = lambda *s: torch.randn(*s, device="cuda")  # noqa: E731
    h, agg = r(M, 256), r(M, 256) * 3
    Wcat, b1 = r(512, 256) / 16, r(512) * 0.1
    W3, b3 = r(256, 512) / 22, r(256) * 0.1
    W4, b4 = r(256, 256) / 16, r(256) * 0.1
    gamma, beta = 1 + 0.1 * r(256), 0.1 * r(256)
    d = lambda t: t.double()  # noqa: E731
    # ABH: fp16(0.5 (h Wcat^T + b1))
    out, *_ = gemm(ABH, h, Wcat, b1, scale=0.5)
    ref = 0.5 * (d(h) @ d(Wcat).t() + d(b1))
    assert out.dtype == torch.float16 and rel_err(out.float(), ref) < 2e-3
    # SILU on two operands
    q, p, *_ = gemm(SILU, h, W3, b3, A2=agg)
    pref = torch.cat([d(h), d(agg)], 1) @ d(W3).t() + d(b3)
    assert rel_err(p, pref) < 2e-3 and rel_err(q, F.silu(pref)) < 2e-3
    q_inf, p_inf, *_ = gemm(SILU, h, W3, b3, A2=agg, train=False)
    assert p_inf is None and torch.equal(q_inf, q)
    # RES_LN
    y, rr, mean, rstd = gemm(RES_LN, q, W4, b4, aux=h, gamma=gamma, beta=beta)
    rref = d(h) + d(q) @ d(W4).t() + d(b4)
    assert rel_err(rr, rref) < 2e-3
    assert rel_err(y, F.layer_norm(rref, (256,), d(gamma), d(beta), 1e-5)) < 3e-3
    assert rel_err(mean, rref.mean(1)) < 2e-3 and rel_err(rstd, 1 / torch.sqrt(rref.var(1, unbiased=False) + 1e-5)) < 3e-3
    y_inf, r_inf, m_inf, _ = gemm(RES_LN, q, W4, b4, aux=h, gamma=gamma, beta=beta, train=False)
    assert r_inf is None and m_inf is None and torch.equal(y_inf, y)
    # PLAIN (Nout = 512, no bias), PLAIN + residual, DSILU
    gp = r(M, 256)
    W3t = W3.t().contiguous()                       # [512, 256]
    c, *_ = gemm(PLAIN, gp, W3t)
    assert rel_err(c, d(gp) @ d(W3)) < 2e-3
    c2, *_ = gemm(PLAIN, gp, W4.t().contiguous(), aux=h)
    assert rel_err(c2, d(gp) @ d(W4) + d(h)) < 2e-3
    gAB = r(M, 512)
    c3, *_ = gemm(PLAIN, gAB, Wcat.t().contiguous())                 # K = 512 from one operand
    assert rel_err(c3, d(gAB) @ d(Wcat)) < 2e-3
    g, *_ = gemm(DSILU, gp, W4.t().contiguous(), aux=p)
    sg = torch.sigmoid(pref)
    assert rel_err(g, (d(gp) @ d(W4)) * sg * (1 + pref * (1 - sg))) < 3e-3


@pytest.mark.parametrize("N", [1, 31, 32, 1000, 65536])
def test_node_wgrad_matches_float64(N):
    """out[Mo,256] = scale G^T X over N rows (split-K, MN-major TF32 operands); also into a column block of a wider matrix."""
    from protein_ensemble_vae_b200.egnn_tc import node_wgrad
    torch.manual_seed(N)
    X = torch.randn(N, 256, device="cuda")
    for Mo in (256, 512):
        G = torch.randn(N, Mo, device="cuda")
        ref = 0.5 * (G.double().t() @ X.double())
        got = node_wgrad(G, X, 0.5)
        assert got.shape == (Mo, 256) and rel_err(got, ref) < 2e-3
    wide = torch.full((256, 512), 7.0, device="cuda")
    G = torch.randn(N, 256, device="cuda")
    node_wgrad(G, X, out=wide[:, 256:])
    assert rel_err(wide[:, 256:], G.double().t() @ X.double()) < 2e-3 and float((wide[:, :256] - 7.0).abs().max()) == 0.0
    a, b = node_wgrad(G, X), node_wgrad(G, X)
    assert torch.equal(a, b)                                  # fixed-order reduction: bit-reproducible


# ------------------------------------------------------------------------------------------------- 3xTF32 forms
@pytest.mark.parametrize("M,K,Nout", [(1000, 256, 256), (77, 256, 512), (4096 + 5, 512, 256), (300, 768, 512)])
def test_gemm3_has_fp32_accuracy(M, K, Nout):
    """pev_node_gemm3 against float64: error at the level of an fp32 library GEMM (tools/check_gemm3.py prints the table),
    more than two orders of magnitude below plain TF32 on the same inputs."""
    from protein_ensemble_vae_b200 import egnn_tc
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g) * torch.exp(torch.randn(M, 1, device="cuda", generator=g))
    W = torch.randn(Nout, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nout, device="cuda", generator=g)
    res = torch.randn(M, Nout, device="cuda", generator=g)
    ref = A.double() @ W.double().t() + b.double() + res.double()
    out = egnn_tc.node_gemm3(A, egnn_tc.split_weight(W), b, res)
    e3 = float((out.double() - ref).abs().max() / ref.abs().max())
    e1 = float((egnn_tc.node_gemm(egnn_tc.EPI_PLAIN, A, W, b, aux=res)[0].double() - ref).abs().max() / ref.abs().max())
    ef = float(((A @ W.t() + b + res).double() - ref).abs().max() / ref.abs().max())
    print(f"gemm3 M={M} K={K} Nout={Nout}: 3xTF32 {e3:.2e}, TF32 {e1:.2e}, library fp32 {ef:.2e}")
    assert e3 < 1.5e-6 and e3 < 0.01 * e1, (e3, e1, ef)           # measured 0.25 - 1.1e-6 (library fp32: 0.4 - 0.9e-6)
    W3t = egnn_tc.split_weight(W, transpose=True)                    # [2K, Nout]: the image of W^T
    assert float((W3t[:K] + W3t[K:] - W.t()).abs().max() / W.abs().max()) < 2e-7


@pytest.mark.parametrize("N,Mo", [(5000, 256), (33, 512), (65536 + 17, 512), (1205760, 256)])
def test_wgrad3_has_fp32_accuracy(N, Mo):
    from protein_ensemble_vae_b200 import egnn_tc
    g = torch.Generator(device="cuda").manual_seed(N)
    G = torch.randn(N, Mo, device="cuda", generator=g)
    X = torch.randn(N, 256, device="cuda", generator=g) + 0.5
    ref = 0.5 * (G.double().t() @ X.double())
    out = egnn_tc.node_wgrad3(G, X, 0.5)
    e3 = float((out.double() - ref).abs().max() / ref.abs().max())
    e1 = float((egnn_tc.node_wgrad(G, X, 0.5).double() - ref).abs().max() / ref.abs().max())
    ef = float(((0.5 * (G.t() @ X)).double() - ref).abs().max() / ref.abs().max())
    print(f"wgrad3 N={N} Mo={Mo}: 3xTF32 {e3:.2e}, TF32 {e1:.2e}, library fp32 {ef:.2e}")
    assert e3 < 6e-6 and e3 < 0.02 * e1, (e3, e1, ef)             # measured 0.3 - 4.3e-6 (library fp32: 0.2 - 6.2e-6)
    wide = torch.zeros(Mo, 512, device="cuda")
    egnn_tc.node_wgrad3(G, X, 0.5, out=wide[:, 256:])
    assert torch.equal(wide[:, 256:], out) and float(wide[:, :256].abs().max()) == 0.0
    assert torch.equal(egnn_tc.node_wgrad3(G, X, 0.5), out)          # fixed-order reduction: bit-reproducible


def test_linear3x_autograd_matches_float64():
    from protein_ensemble_vae_b200 import egnn_tc
    torch.manual_seed(5)
    N = 3000
    x1 = torch.randn(N, 256, device="cuda", requires_grad=True)
    x2 = torch.randn(N, 256, device="cuda", requires_grad=True)
    W = (torch.randn(256, 512, device="cuda") / 16).requires_grad_()
    b = torch.randn(256, device="cuda", requires_grad=True)
    coef = torch.randn(N, 256, device="cuda")
    y = egnn_tc.Linear3x.apply(W, b, x1, x2)
    grads = torch.autograd.grad((y * coef).sum(), [x1, x2, W, b])
    d = [t.detach().double().requires_grad_() for t in (x1, x2, W, b)]
    yr = torch.cat([d[0], d[1]], 1) @ d[2].t() + d[3]
    gr = torch.autograd.grad((yr * coef.double()).sum(), d)
    assert float((y.detach().double() - yr.detach()).abs().max() / yr.detach().abs().max()) < 3e-6
    for a, r in zip(grads, gr):
        assert float((a.double() - r).abs().max() / r.abs().max()) < 6e-6
