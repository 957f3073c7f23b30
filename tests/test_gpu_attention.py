"""Encoder attention kernels (csrc/attn_kernels.cu) against a float64 per-conformer restatement: forward, all three
gradients, ragged lengths that are not multiples of any tile size, both head shapes of the reference (8 x 64, 4 x 128)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, lengths, H):
    d = qkv.shape[1] // 3
    hd = d // H
    outs, o0 = [], 0
    for n in lengths:
        x = qkv[o0:o0 + n]
        q, k, v = (x[:, i * d:(i + 1) * d].reshape(n, H, hd).transpose(0, 1) for i in range(3))
        p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd), -1)
        outs.append((p @ v).transpose(0, 1).reshape(n, d))
        o0 += n
    return torch.cat(outs)


@pytest.mark.parametrize("lengths,H", [((48, 30, 17), 8), ((256, 256), 8), ((300, 1, 129, 512), 4), ((64,), 4)])
def test_packed_self_attention_matches_float64(lengths, H):
    from protein_ensemble_vae_b200 import attention as pa
    from protein_ensemble_vae_b200.encoder import _Packing
    torch.manual_seed(sum(lengths))
    B, L = len(lengths), max(lengths)
    mask = torch.zeros(B, L, device="cuda")
    for b, n in enumerate(lengths):
        mask[b, :n] = 1
    pk = _Packing(mask, B, L, "cuda")
    N = sum(lengths)
    qkv = torch.randn(N, 3 * 512, device="cuda", requires_grad=True)
    coef = torch.randn(N, 512, device="cuda")
    out = pa.self_attention(qkv, pk, H, 0.0, False)
    (g,) = torch.autograd.grad((out * coef).sum(), qkv)
    q64 = qkv.detach().double().requires_grad_()
    ref = _ref(q64, lengths, H)
    (gr,) = torch.autograd.grad((ref * coef.double()).sum(), q64)
    eo = float((out.double() - ref).abs().max() / ref.abs().max())
    eg = [float((g[:, i * 512:(i + 1) * 512].double() - gr[:, i * 512:(i + 1) * 512]).abs().max()
                / gr[:, i * 512:(i + 1) * 512].abs().max()) for i in range(3)]
    print(lengths, H, "out %.1e  dq %.1e dk %.1e dv %.1e" % (eo, *eg))
    assert eo < 3e-3 and max(eg) < 5e-3, (eo, eg)          # TF32 operands (10-bit mantissa), fp32 accumulation


def test_attention_dropout_is_consistent_between_forward_and_backward():
    """With dropout the backward pass re-derives the mask from (seed, index): the gradient must equal the derivative of the
    forward that was actually computed -- checked by linearity: out is linear in V, so <dOut, out> = <dV, V>."""
    from protein_ensemble_vae_b200 import attention as pa
    from protein_ensemble_vae_b200.encoder import _Packing
    torch.manual_seed(3)
    lengths = (70, 33)
    mask = torch.zeros(2, 70, device="cuda")
    mask[0, :70] = 1
    mask[1, :33] = 1
    pk = _Packing(mask, 2, 70, "cuda")
    qkv = torch.randn(103, 1536, device="cuda", requires_grad=True)
    coef = torch.randn(103, 512, device="cuda")
    out = pa.self_attention(qkv, pk, 8, 0.3, False)
    (g,) = torch.autograd.grad((out * coef).sum(), qkv)
    lhs = float((out * coef).sum())
    rhs = float((g[:, 1024:] * qkv[:, 1024:].detach()).sum())
    assert abs(lhs - rhs) < 5e-3 * max(abs(lhs), 1.0), (lhs, rhs)
    plain = pa.self_attention(qkv, pk, 8, 0.0, False)
    assert float((plain - out).abs().max()) > 1e-2          # dropout did something
