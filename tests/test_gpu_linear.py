"""``pev_linear`` / ``pev_linear_wgrad`` through the Python binding (tc_linear.py): column-block operands (leading dimensions),
bias / ReLU / residual epilogues, wide outputs, the single-launch multi-block weight gradient, TF32 and 3xTF32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precise,tol", [(False, 2e-3), (True, 2e-6)])
@pytest.mark.parametrize("M,K,Nout", [(300, 512, 1536), (1000, 1280, 256), (129, 256, 1024), (5, 1024, 512)])
def test_linear_forward_epilogues_and_strides(M, K, Nout, precise, tol):
    from protein_ensemble_vae_b200 import tc_linear
    from protein_ensemble_vae_b200.egnn_tc import split_weight
    g = torch.Generator(device="cuda").manual_seed(M + K + Nout)
    wide = torch.randn(M, K + 256, device="cuda", generator=g)
    x = wide[:, 128:128 + K]                                       # a column block: lda = K + 256
    W = torch.randn(Nout, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nout, device="cuda", generator=g)
    resw = torch.randn(M, Nout + 64, device="cuda", generator=g)
    res = resw[:, 64:]
    outw = torch.full((M, Nout + 128), 7.0, device="cuda")
    Wk = split_weight(W) if precise else W
    ref = x.double() @ W.double().t() + b.double()
    y = tc_linear.linear_fwd(x, Wk, b, relu=True, precise=precise)
    assert float((y.double() - ref.clamp_min(0)).abs().max() / ref.abs().max()) < tol
    tc_linear.linear_fwd(x, Wk, b, res=res, precise=precise, out=outw[:, 128:])
    assert float((outw[:, 128:].double() - (ref + res.double())).abs().max() / ref.abs().max()) < tol
    assert float((outw[:, :128] - 7.0).abs().max()) == 0.0          # nothing written outside the column block


@pytest.mark.parametrize("precise,tol", [(False, 3e-3), (True, 6e-6)])
@pytest.mark.parametrize("N,Nout,K", [(4000, 1536, 512), (70000, 256, 1280), (33, 512, 256)])
def test_linear_wgrad_all_blocks_in_one_launch(N, Nout, K, precise, tol):
    from protein_ensemble_vae_b200 import tc_linear
    g = torch.Generator(device="cuda").manual_seed(N + K)
    G = torch.randn(N, Nout, device="cuda", generator=g)
    X = torch.randn(N, K + 256, device="cuda", generator=g)[:, 256:] + 0.25
    ref = G.double().t() @ X.double()
    out = tc_linear.linear_wgrad(G, X, precise)
    assert out.shape == (Nout, K)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < tol
    assert torch.equal(tc_linear.linear_wgrad(G, X, precise), out)  # fixed-order reduction


def test_tclinear_autograd_relu_and_residual():
    from protein_ensemble_vae_b200 import tc_linear
    torch.manual_seed(2)
    x = torch.randn(700, 512, device="cuda", requires_grad=True)
    W = (torch.randn(1024, 512, device="cuda") / 22).requires_grad_()
    b = torch.randn(1024, device="cuda", requires_grad=True)
    W2 = (torch.randn(512, 1024, device="cuda") / 32).requires_grad_()
    b2 = torch.randn(512, device="cuda", requires_grad=True)
    coef = torch.randn(700, 512, device="cuda")
    y = tc_linear.linear(tc_linear.linear(x, W, b, relu=True, precise=True), W2, b2, precise=True, res=x)
    grads = torch.autograd.grad((y * coef).sum(), [x, W, b, W2, b2])
    d = [t.detach().double().requires_grad_() for t in (x, W, b, W2, b2)]
    yr = torch.relu(d[0] @ d[1].t() + d[2]) @ d[3].t() + d[4] + d[0]
    gr = torch.autograd.grad((yr * coef.double()).sum(), d)
    assert float((y.detach().double() - yr.detach()).abs().max() / yr.detach().abs().max()) < 3e-6
    for a, r in zip(grads, gr):
        assert float((a.double() - r).abs().max() / r.abs().max()) < 1e-5


@pytest.mark.parametrize("precise", [False, True])
def test_fused_dropout_and_relu_backward_match_the_forward_mask(precise):
    """y = dropout(relu(x W^T + b)) is positively homogeneous of degree 1 in (W, b): <coef, y> = <dW, W> + <db, b> holds only
    if the backward pass re-derives exactly the dropout mask (counter hash) and ReLU pattern of the forward pass."""
    from protein_ensemble_vae_b200 import tc_linear
    torch.manual_seed(4)
    x = torch.randn(900, 512, device="cuda")
    W = (torch.randn(1024, 512, device="cuda") / 22).requires_grad_()
    b = torch.randn(1024, device="cuda", requires_grad=True)
    coef = torch.randn(900, 1024, device="cuda")
    y = tc_linear.linear(x, W, b, relu=True, precise=precise, p_drop=0.25)
    frac = float((y == 0).float().mean())
    assert 0.55 < frac < 0.70                                          # ~50 % ReLU zeros, then 25 % of the rest dropped
    gW, gb = torch.autograd.grad((y * coef).sum(), [W, b])
    lhs = float((y.detach().double() * coef.double()).sum())
    rhs = float((gW.double() * W.detach().double()).sum() + (gb.double() * b.detach().double()).sum())
    assert abs(lhs - rhs) < (2e-3 if not precise else 2e-5) * max(abs(lhs), 1.0), (lhs, rhs)
    res = torch.randn(900, 1024, device="cuda")
    y2 = tc_linear.linear(x, W, b, precise=precise, res=res, p_drop=0.5)
    kept = float(((y2 - res).abs() > 0).float().mean())
    assert 0.45 < kept < 0.55                                          # dropout acts before the residual is added


def test_linear_with_128_wide_output_runs_on_the_3xtf32_kernels():
    """256 -> 128 heads (n / c offset heads, latent_to_coords[4]): forward, data gradient and the half-empty weight-gradient
    row block."""
    from protein_ensemble_vae_b200 import tc_linear
    torch.manual_seed(6)
    x = torch.randn(3000, 256, device="cuda", requires_grad=True)
    W = (torch.randn(128, 256, device="cuda") / 16).requires_grad_()
    b = torch.randn(128, device="cuda", requires_grad=True)
    coef = torch.randn(3000, 128, device="cuda")
    assert tc_linear.supported(x, W)
    y = tc_linear.linear(x, W, b, relu=True)
    grads = torch.autograd.grad((y * coef).sum(), [x, W, b])
    d = [t.detach().double().requires_grad_() for t in (x, W, b)]
    yr = torch.relu(d[0] @ d[1].t() + d[2])
    gr = torch.autograd.grad((yr * coef.double()).sum(), d)
    assert y.shape == (3000, 128) and grads[1].shape == (128, 256)
    assert float((y.detach().double() - yr.detach()).abs().max() / yr.detach().abs().max()) < 3e-6
    for a, r in zip(grads, gr):
        assert float((a.double() - r).abs().max() / r.abs().max()) < 1e-5
