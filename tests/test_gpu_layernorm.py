"""GPU parity of the fused LayerNorm (csrc/streaming.cu layernorm_{fwd,bwd}_kernel, atq.layer_norm): the op in front of
every ternary GEMM of the reference's transformer block (models/text_encoder.py:77,232,244 use nn.LayerNorm).
Checker: torch's F.layer_norm in float64 on the same device; tolerance 2e-5 * max|ref| + 1e-6 on values and
gradients (fp32 two-pass statistics; the gamma / beta gradients are fixed-order sums, so they are also run-to-run identical).
Also checks that the max|y| the forward kernel leaves is the exact maximum, and that the operand split of the GEMM that
follows picks it up instead of reducing again."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import atq
import atq._engine as eng

DEV = "cuda:0"


def _case(rows_shape, cols, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(*rows_shape, cols, device=DEV, generator=g) * 3.0 + 0.5
    w = 1.0 + 0.2 * torch.randn(cols, device=DEV, generator=g)
    b = 0.1 * torch.randn(cols, device=DEV, generator=g)
    dy = torch.randn(*rows_shape, cols, device=DEV, generator=g)
    return x, w, b, dy


@pytest.mark.parametrize("rows_shape,cols", [((1,), 4), ((7,), 12), ((3, 5), 128), ((16, 50), 192), ((2, 197), 768),
                                             ((1031,), 1024), ((4099,), 260), ((64, 197), 768)])
def test_layer_norm_matches_float64(rows_shape, cols):
    x, w, b, dy = _case(rows_shape, cols, 11 + cols)
    xi, wi, bi = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = atq.layer_norm(xi, wi, bi, 1e-5)
    y.backward(dy)
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    yr = F.layer_norm(xr, (cols,), wr, br, 1e-5)
    yr.backward(dy.double())
    for name, got, want in (("y", y, yr), ("dx", xi.grad, xr.grad), ("dgamma", wi.grad, wr.grad), ("dbeta", bi.grad, br.grad)):
        tol = 2e-5 * float(want.abs().max()) + 1e-6
        err = float((got.double() - want.detach()).abs().max())
        assert err <= tol, (name, rows_shape, cols, err, tol)


def test_layer_norm_parameter_gradients_are_deterministic():
    x, w, b, dy = _case((64, 197), 768, 3)
    grads = []
    for _ in range(2):
        xi, wi, bi = (t.clone().requires_grad_(True) for t in (x, w, b))
        atq.layer_norm(xi, wi, bi).backward(dy)
        grads.append((xi.grad.clone(), wi.grad.clone(), bi.grad.clone()))
    for a, c in zip(*grads):
        assert torch.equal(a, c)


def test_layer_norm_leaves_exact_absmax_for_the_following_gemm():
    atq.set_gemm_mode("parity")
    x, w, b, _ = _case((32, 197), 768, 5)
    lin = atq.TernaryLinear(768, 256).to(DEV)
    before = dict(eng._ABSMAX_HINT_STATS)
    y = atq.layer_norm(x, w, b)
    assert eng._ABSMAX_HINT_STATS["set"] == before["set"] + 1
    slot = eng._known_absmax(y.reshape(-1, 768))
    assert slot is not None
    sc, inv = float(slot[1]), float(slot[2])
    amax = float(y.abs().max())
    assert 2.0 ** 14 <= amax * sc < 2.0 ** 15 and sc * inv == 1.0
    hits = eng._ABSMAX_HINT_STATS["hit"]
    out = lin(y)
    assert eng._ABSMAX_HINT_STATS["hit"] == hits + 1          # the split of y reused the slot
    # and the result is the one the stand-alone reduction gives
    ref = lin(y.clone())
    assert torch.equal(out, ref)
    # an in-place edit invalidates the hint
    y.mul_(2.0)
    assert eng._known_absmax(y.reshape(-1, 768)) is None


def test_layer_norm_rejects_unsupported_shapes():
    x = torch.randn(4, 1028, device=DEV)
    w = torch.ones(1028, device=DEV)
    assert not atq.layer_norm_supported(x, w, w)
    with pytest.raises(RuntimeError):
        atq.layer_norm(x, w, w)
    x = torch.randn(4, 10, device=DEV)
    w = torch.ones(10, device=DEV)
    assert not atq.layer_norm_supported(x, w, w)
