"""The harness transformer block (workloads.models.TernaryBlock) against a golden fixture produced by RUNNING the
reference's own models.text_encoder.TernaryTransformerLayer (tests/golden/gen_block_golden.py):
  * CPU: with the oracle layers it reproduces the reference's outputs and gradients (pins harness + oracle);
  * GPU: with the B200 `atq` package - fused attention core, fused FFN activation, gated-residual kernels - it
    matches the reference's outputs and gradients within the north-star tolerance rtol 1e-2 / atol 1e-3."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

G = np.load(os.path.join(ROOT, "tests", "golden", "block_golden.npz"))
E, HEADS, FF = 64, 1, 128


def _block(layers, device):
    from workloads import models as M
    blk = M.TernaryBlock(layers, E, HEADS, FF, 0.0, True, 0.2)
    sd = {k[len("param."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("param.")}
    missing, unexpected = blk.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return blk.to(device)


def _run(blk, device):
    x = torch.from_numpy(G["x"]).to(device)
    pad = torch.from_numpy(G["pad"]).to(device)
    gy = torch.from_numpy(G["gy"]).to(device)
    blk.eval()
    with torch.no_grad():
        y_eval = blk(x, key_padding_mask=pad)
    blk.train()
    xr = x.clone().requires_grad_(True)
    y = blk(xr, key_padding_mask=pad)
    y.backward(gy)
    grads = {n: p.grad for n, p in blk.named_parameters() if p.grad is not None}
    return y_eval, y.detach(), xr.grad, grads


def _close(t, want, tol):
    # the fixture's tensors span |values| up to ~1.7e3 (LayerNorm gains, gy ~ N(0,1)): the absolute part of the
    # tolerance is taken relative to the tensor's own scale, atol * max(1, max|want|)
    scale = max(1.0, float(want.abs().max()))
    return torch.allclose(t, want, rtol=tol["rtol"], atol=tol["atol"] * scale)


def _check(got, tol, scalar_rel):
    y_eval, y_train, dx, grads = got
    for name, t in (("y_eval", y_eval), ("y_train", y_train), ("dx", dx)):
        want = torch.from_numpy(G[name])
        assert _close(t.cpu(), want, tol), (name, (t.cpu() - want).abs().max().item())
    want_grads = {k[len("grad."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("grad.")}
    assert set(grads) == set(want_grads)  # same parameters receive gradients (RPB weights do, none is None)
    for n, w in want_grads.items():
        if n.endswith("k_proj.bias"):
            # mathematically zero (a key bias shifts every score of a query by the same amount); the reference's
            # 7e-5 is pure rounding noise of its fp32 softmax backward, so only the magnitude is checked
            assert float(grads[n].abs().max()) < 1e-4 * float(want_grads["self_attn.q_proj.bias"].abs().max())
            continue
        g = grads[n].cpu()
        if w.numel() == 1:  # alpha / gate: sums over whole tensors
            assert abs(float(g) - float(w)) <= scalar_rel * abs(float(w)) + 10 * tol["atol"], (n, float(g), float(w))
        else:
            assert _close(g, w, tol), (n, (g - w).abs().max().item())


def test_block_with_oracle_layers_reproduces_reference_cpu(monkeypatch):
    from workloads import models as M
    monkeypatch.setattr(M, "FUSED_ATTENTION_CORE", False)  # the reference's own matmul / softmax sequence
    _check(_run(_block(M.oracle_layers(), "cpu"), "cpu"), dict(rtol=1e-6, atol=1e-7), 1e-6)
    monkeypatch.setattr(M, "FUSED_ATTENTION_CORE", True)   # torch SDPA evaluation of the same maths
    _check(_run(_block(M.oracle_layers(), "cpu"), "cpu"), dict(rtol=1e-4, atol=1e-5), 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_block_on_b200_matches_reference(fused, monkeypatch):
    import atq
    from workloads import models as M
    monkeypatch.setattr(M, "FUSED_ATTENTION_CORE", fused)
    monkeypatch.setattr(M, "FUSED_FFN", fused)
    blk = _block(atq, "cuda:0")
    if fused:  # the fused paths are really taken for this shape
        assert blk.self_attn._core is not None and blk._ffn is not None and blk._gres is not None
        assert atq.attention_core_supported(E, HEADS, 20)
    _check(_run(blk, "cuda:0"), dict(rtol=1e-2, atol=1e-3), 1e-2)
