"""GPU parity: fused attention core (atq_attention_fwd / atq_attention_bwd) against an fp64
evaluation of the reference's explicit sequence (models/text_encoder.py:117-163): matmul, key-padding
masked_fill, softmax, dropout (same keep mask, regenerated on the host from the kernels' counter hash), matmul.
Tolerance (BASELINE north_star): rtol 1e-2 / atol 1e-3 on outputs and gradients."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import atq
from atq import attention as A

DEV = "cuda:0"


def _reference(q, k, v, heads, pad, scale, keep, p):
    b, l, e = q.shape
    d = e // heads
    qh, kh, vh = (t.double().view(b, l, heads, d).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-2, -1) * scale
    if pad is not None:
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    pr = torch.softmax(s, dim=-1)
    if p > 0:
        pr = pr * keep.double() / (1.0 - p)
    return (pr @ vh).transpose(1, 2).reshape(b, l, e)


def _run_case(b, heads, l, p, with_pad, mode, tol, hd=64):
    atq.set_gemm_mode(mode)
    try:
        g = torch.Generator().manual_seed(b * 1000 + heads * 100 + l)
        e = heads * hd
        q, k, v = (torch.randn(b, l, e, generator=g) for _ in range(3))
        dout = torch.randn(b, l, e, generator=g)
        pad = None
        if with_pad:
            lens = torch.randint(max(1, l // 3), l + 1, (b,), generator=g)
            pad = torch.arange(l)[None, :] >= lens[:, None]
        seed_val = 1234567891011 + l
        seed = torch.tensor([seed_val], dtype=torch.int64, device=DEV)
        scale = 1.0 / math.sqrt(hd)
        qg, kg, vg = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
        out = A.attention_core(qg, kg, vg, heads, None if pad is None else pad.to(DEV), scale, p, True, seed=seed)
        out.backward(dout.to(DEV))
        keep, p_eff = A.dropout_keep_mask(seed_val, b, heads, l, p)
        keep = torch.from_numpy(keep)
        if p > 0:
            frac = keep.float().mean().item()
            assert abs(p_eff - p) < 1e-4 and abs(frac - (1 - p)) < 0.02, (p_eff, frac)
            p = p_eff
        qr, kr, vr = (t.double().requires_grad_(True) for t in (q, k, v))
        ref = _reference(qr, kr, vr, heads, pad, scale, keep, p)
        ref.backward(dout.double())
        assert torch.allclose(out.detach().cpu().double(), ref.detach(), **tol), (out.detach().cpu().double() - ref.detach()).abs().max()
        for name, got, want in (("dq", qg.grad, qr.grad), ("dk", kg.grad, kr.grad), ("dv", vg.grad, vr.grad)):
            err = (got.cpu().double() - want).abs().max()
            assert torch.allclose(got.cpu().double(), want, **tol), (name, err)
    finally:
        atq.set_gemm_mode("parity")


@pytest.mark.parametrize("b,heads,l,p,with_pad", [
    (2, 2, 50, 0.0, False), (2, 3, 197, 0.0, False), (3, 2, 50, 0.1, True), (1, 12, 197, 0.1, False),
    (2, 1, 256, 0.0, True), (2, 2, 1, 0.0, False), (1, 2, 33, 0.25, True), (2, 2, 128, 0.0, False), (1, 1, 129, 0.1, True)])
def test_attention_core_parity_mode(b, heads, l, p, with_pad):
    _run_case(b, heads, l, p, with_pad, "parity", dict(rtol=1e-2, atol=1e-3))


@pytest.mark.parametrize("hd,b,heads,l,p,with_pad", [(24, 16, 8, 50, 0.1, True), (24, 2, 8, 50, 0.0, False), (32, 2, 4, 197, 0.1, True),
                                                      (8, 1, 3, 17, 0.0, False), (48, 2, 2, 256, 0.0, True), (56, 1, 2, 130, 0.1, False)])
def test_attention_core_narrow_heads(hd, b, heads, l, p, with_pad):
    """head_dim < 64 (BASELINE config 2: embed 192 / 8 heads = 24): zero-padded 64-wide tiles, same tolerance."""
    _run_case(b, heads, l, p, with_pad, "parity", dict(rtol=1e-2, atol=1e-3), hd=hd)


@pytest.mark.parametrize("b,heads,l,p,with_pad", [(2, 2, 50, 0.0, True), (1, 3, 197, 0.1, False)])
def test_attention_core_fast_mode(b, heads, l, p, with_pad):
    # single bf16 operands: ~2^-9 relative per product
    _run_case(b, heads, l, p, with_pad, "fast", dict(rtol=3e-2, atol=3e-2))


def test_attention_core_strided_inputs_and_determinism():
    """q/k/v as column slices of one fused [B, L, 3E] projection output (pitch 3E); same seed -> same bits."""
    torch.manual_seed(0)
    b, heads, l = 2, 2, 70
    e = heads * 64
    qkv = torch.randn(b, l, 3 * e, device=DEV)
    q, k, v = qkv[..., :e], qkv[..., e:2 * e], qkv[..., 2 * e:]
    seed = torch.tensor([77], dtype=torch.int64, device=DEV)
    o1 = A.attention_core(q, k, v, heads, None, None, 0.1, True, seed=seed)
    o2 = A.attention_core(q.contiguous(), k.contiguous(), v.contiguous(), heads, None, None, 0.1, True, seed=seed)
    assert torch.equal(o1, o2)
    o3 = A.attention_core(q, k, v, heads, None, None, 0.1, False)  # eval: no dropout
    ref = _reference(q.cpu(), k.cpu(), v.cpu(), heads, None, 1 / 8, None, 0.0)
    assert torch.allclose(o3.cpu().double(), ref, rtol=1e-2, atol=1e-3)


def test_attention_core_rejects_unsupported_shapes():
    x = torch.randn(1, 300, 64, device=DEV)
    with pytest.raises(RuntimeError):
        A.attention_core(x, x, x, 1)
    y = torch.randn(1, 8, 40, device=DEV)
    with pytest.raises(RuntimeError):
        A.attention_core(y, y, y, 2)   # head_dim 20: not a multiple of 8
    z = torch.randn(1, 8, 144, device=DEV)
    with pytest.raises(RuntimeError):
        A.attention_core(z, z, z, 2)   # head_dim 72 > 64
    with pytest.raises(RuntimeError):
        A.attention_core(torch.randn(1, 8, 64), torch.randn(1, 8, 64), torch.randn(1, 8, 64), 1)  # CPU tensors
