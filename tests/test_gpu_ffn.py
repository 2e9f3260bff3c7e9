"""GPU parity: atq.fused_ffn (linear1 -> gelu -> dropout -> linear2 with the activation / dropout / operand split
fused into one streaming kernel per direction) against the unfused layer sequence with the SAME dropout mask."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import atq
from atq import attention as A

DEV = "cuda:0"
TOL = dict(rtol=1e-2, atol=1e-3)


def _layers(k, h, m, seed):
    torch.manual_seed(seed)
    l1 = atq.ResidualPrecisionBoostLinear(k, h, precision_ratio=0.2, sparsity_target=0.1).to(DEV)
    l2 = atq.ResidualPrecisionBoostLinear(h, m, precision_ratio=0.4, sparsity_target=0.1).to(DEV)
    with torch.no_grad():
        l1.alpha.fill_(0.9)
        l2.alpha.fill_(1.1)
    return l1, l2


@pytest.mark.parametrize("tokens,k,h,m,p", [(800, 192, 384, 192, 0.0), (800, 192, 384, 192, 0.1), (50, 64, 72, 40, 0.25),
                                            (3000, 768, 3072, 768, 0.1), (7, 96, 8, 16, 0.5)])
@pytest.mark.parametrize("mode", ["parity", "fast"])
def test_fused_ffn_matches_layer_sequence(tokens, k, h, m, p, mode):
    atq.set_gemm_mode(mode)
    try:
        l1, l2 = _layers(k, h, m, tokens + h)
        g = torch.Generator(device=DEV).manual_seed(1)
        shape = (4, tokens // 4, k) if tokens % 4 == 0 else (tokens, k)
        x = torch.randn(*shape, device=DEV, generator=g)
        n_tok = x.numel() // k
        gy = torch.randn(*x.shape[:-1], m, device=DEV, generator=g)
        seed_val = 424242 + tokens
        seed = torch.tensor([seed_val], dtype=torch.int64, device=DEV)
        keep, p_eff = A.dropout_keep_mask_flat(seed_val, n_tok * h, p)
        keep = torch.from_numpy(keep).to(DEV).view(*x.shape[:-1], h)
        if p > 0 and keep.numel() > 10000:
            assert abs(keep.float().mean().item() - (1 - p)) < 0.02

        xa = x.clone().requires_grad_(True)
        ya = atq.fused_ffn(l1, l2, xa, p, True, seed=seed)
        ya.backward(gy)
        got = [ya.detach(), xa.grad] + [t.grad.clone() for t in (l1.weight, l1.alpha, l1.bias, l2.weight, l2.alpha, l2.bias)]
        for t in (l1, l2):
            t.zero_grad(set_to_none=True)

        xb = x.clone().requires_grad_(True)
        hmid = F.gelu(l1(xb))
        if p > 0:
            hmid = hmid * keep / (1.0 - p_eff)
        yb = l2(hmid)
        yb.backward(gy)
        want = [yb.detach(), xb.grad] + [t.grad.clone() for t in (l1.weight, l1.alpha, l1.bias, l2.weight, l2.alpha, l2.bias)]
        names = ["y", "dx", "dW1", "dalpha1", "db1", "dW2", "dalpha2", "db2"]
        tol = TOL if mode == "parity" else dict(rtol=3e-2, atol=3e-2)
        for n, a, b in zip(names, got, want):
            if "alpha" in n:  # scalar sums over the whole layer: relative
                assert abs(float(a) - float(b)) <= 2e-2 * abs(float(b)) + 5e-2, (n, float(a), float(b))
            else:
                assert torch.allclose(a, b, **tol), (n, (a - b).abs().max().item())
        # dW stays under the mask
        assert torch.count_nonzero(got[2] * (1 - l1.precision_mask)) == 0
    finally:
        atq.set_gemm_mode("parity")


def test_fused_ffn_eval_mode_has_no_dropout_and_rejects_other_layers():
    l1, l2 = _layers(64, 128, 32, 0)
    x = torch.randn(10, 64, device=DEV)
    y = atq.fused_ffn(l1, l2, x, 0.5, training=False)
    assert torch.allclose(y, l2(F.gelu(l1(x))), **TOL)
    t = atq.TernaryLinear(64, 128).to(DEV)
    assert not atq.fused_ffn_supported(t, l2, x)
    with pytest.raises(RuntimeError):
        atq.fused_ffn(t, l2, x)


@pytest.mark.parametrize("shape,p", [((4, 50, 192), 0.0), ((4, 50, 192), 0.1), ((3, 197, 768), 0.1), ((7, 4), 0.5)])
def test_gated_residual_matches_torch_sequence(shape, p):
    g = torch.Generator(device=DEV).manual_seed(3)
    src = torch.randn(*shape, device=DEV, generator=g)
    h = torch.randn(*shape, device=DEV, generator=g)
    gy = torch.randn(*shape, device=DEV, generator=g)
    gate_param = torch.tensor([0.8], device=DEV)
    seed_val = 99 + shape[-1]
    seed = torch.tensor([seed_val], dtype=torch.int64, device=DEV)
    keep, p_eff = A.dropout_keep_mask_flat(seed_val, src.numel(), p, stream_id=0x6A7ED)
    keep = torch.from_numpy(keep).to(DEV).view(*shape).float()
    outs = []
    for fused in (True, False):
        s_, h_, gp = src.clone().requires_grad_(True), h.clone().requires_grad_(True), gate_param.clone().requires_grad_(True)
        gate = torch.sigmoid(gp)
        if fused:
            out = atq.gated_residual(s_, h_, gate, p, True, seed=seed)
        else:
            out = s_ + (h_ * keep / (1.0 - p_eff)) * gate
        out.backward(gy)
        outs.append((out.detach(), s_.grad, h_.grad, gp.grad))
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), (a - b).abs().max()
    # eval mode: no dropout (the kernel contracts the multiply-add into one FMA, so compare with a tolerance)
    assert torch.allclose(atq.gated_residual(src, h, torch.sigmoid(gate_param), 0.3, training=False),
                          src + h * torch.sigmoid(gate_param), rtol=1e-6, atol=1e-6)
