"""GPU: the reference's --use_amp path (train_multimodal.py:411-416, 487-488: forward under CUDA autocast, GradScaler)
runs on the B200 layers, and FlatAdamW survives a checkpoint / resume (train_multimodal.py:652-662 saves the optimizer
state) with the same trajectory as torch.optim.AdamW."""
import io

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

import atq
from atq.optim import FlatAdamW

DEV = "cuda:0"


def test_layers_under_autocast_and_grad_scaler():
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(64, 192), atq.ResidualPrecisionBoostLinear(192, 96, 0.2, True, 0.15), nn.GELU(),
                        atq.TernaryLinear(96, 32)).to(DEV)
    x = torch.randn(40, 64, device=DEV)
    want = net(x).mean()
    want.backward()
    g_want = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    net.zero_grad(set_to_none=True)
    scaler = torch.amp.GradScaler("cuda", init_scale=256.0)  # (the default 2^16 overflows the fp16 nn.Linear grads by design)
    with torch.autocast("cuda", dtype=torch.float16):
        y = net(x)   # nn.Linear hands fp16 activations to the ternary layers, exactly as the reference's attention does
        assert y.dtype == torch.float32, "ternary layers compute and return fp32 under autocast"
        loss = y.mean()
    scaler.scale(loss).backward()
    inv = 1.0 / scaler.get_scale()
    g_got = [p.grad * inv for p in net.parameters() if p.grad is not None]
    assert abs(float(loss) - float(want)) <= 2e-2 * abs(float(want)) + 1e-2   # fp16 nn.Linear in front: loose
    assert len(g_got) == len(g_want)
    for a, b in zip(g_got, g_want):
        assert torch.isfinite(a).all()
        assert torch.allclose(a, b, rtol=5e-2, atol=5e-2 * float(b.abs().max()))
    assert net[3].weight.grad is None  # TernaryLinear contract unchanged


def test_block_under_autocast_runs():
    from workloads import models as M
    torch.manual_seed(1)
    blk = M.TernaryBlock(atq, 128, 2, 256, 0.1, True, 0.1).to(DEV).train()
    x = torch.randn(4, 70, 128, device=DEV, requires_grad=True)
    with torch.autocast("cuda", dtype=torch.float16):
        y = blk(x)
    y.float().square().mean().backward()
    assert torch.isfinite(x.grad).all() and all(torch.isfinite(p.grad).all() for p in blk.parameters() if p.grad is not None)


def _toy(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(20, 33), nn.Tanh(), nn.Linear(33, 7)).to(DEV)


def _step(m, opt, i):
    g = torch.Generator(device=DEV).manual_seed(50 + i)
    x = torch.randn(16, 20, device=DEV, generator=g)
    opt.zero_grad(set_to_none=True)
    m(x).square().mean().backward()
    opt.step()


def test_flat_adamw_checkpoint_resume_matches_torch():
    hp = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    a, b = _toy(0), _toy(0)
    oa, ob = FlatAdamW(a.parameters(), **hp), torch.optim.AdamW(b.parameters(), **hp)
    for i in range(3):
        _step(a, oa, i); _step(b, ob, i)
    # checkpoint through a CPU round trip (torch.load(map_location="cpu") is what a resume script does)
    buf = io.BytesIO()
    torch.save({"model": a.state_dict(), "opt": oa.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, map_location="cpu")
    a2 = _toy(123)
    a2.load_state_dict(ck["model"])
    oa2 = FlatAdamW(a2.parameters(), **hp)
    oa2.load_state_dict(ck["opt"])
    assert oa2.param_groups[0]["step"].is_cuda and float(oa2.param_groups[0]["step"]) == 3.0
    # an optimizer that already stepped and is then re-loaded must drop its cached pointer table too
    _step(a, oa, 99)
    oa.load_state_dict(ck["opt"]); a.load_state_dict(ck["model"])
    for i in range(3, 6):
        _step(a2, oa2, i); _step(b, ob, i); _step(a, oa, i)
    for pa, pa2, pb in zip(a.parameters(), a2.parameters(), b.parameters()):
        assert torch.allclose(pa2, pb, rtol=1e-5, atol=1e-6)
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6)
