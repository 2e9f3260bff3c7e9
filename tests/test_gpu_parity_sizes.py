"""GPU parity at BASELINE config-3 / config-4 sizes.  The checker is the REFERENCE'S OWN layer code (staged copy of
atq/layers.py, atq/precision_boost.py, atq/quantizers.py imported under the alias `ref_atq`) run on the same device in
float32 and in float64 -- at 8192 x 8192 weights and 16 384 tokens a CPU run of the reference takes minutes, its CUDA
run (torch.sort + cuBLAS/fp64) seconds, and it is the same Python.

Criterion (as tests/test_gpu_dropin.py): |ours - f64| <= 1e-2 |f64| + 1e-3 scale + 2 max|f32 - f64| elementwise;
integer outputs (T) bit-exact.  At 1.3e8 outputs per tensor and a contraction length of 8192 the fp32 accumulation
error of EITHER implementation (cuBLAS sgemm is 1.1e-3 from float64 there) reaches the absolute tolerance, so the
elementwise bound is allowed to be missed by at most 1e-7 of the elements, and by those by less than the bound itself.
"""
import pytest
import torch

from conftest import have_staged_reference, load_reference_atq

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not have_staged_reference(), reason="oracle/_ref not staged (python oracle/install_ref.py)")]

import atq

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _ieee_fp32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()


def _check(name, ours, f32, f64, relative):
    ours, f32, f64 = ours.detach().double(), f32.detach().double(), f64.detach().double()
    ref_err = float((f32 - f64).abs().max())
    scale = float(f64.abs().max()) if relative else 1.0
    bound = 1e-2 * f64.abs() + 1e-3 * scale + 2.0 * ref_err
    miss = (ours - f64).abs() - bound
    worst = float(miss.max())
    n_over = int((miss > 0).sum())
    if n_over and n_over <= 1e-7 * miss.numel() and float((miss - bound).max()) <= 0:
        return
    assert worst <= 0, (f"{name}: {int((miss > 0).sum())}/{miss.numel()} over by up to {worst:.3e}; |f64| max "
                        f"{float(f64.abs().max()):.3e}, reference fp32 err {ref_err:.3e}, ours {float((ours - f64).abs().max()):.3e}")


def _run_layer(mod, x, gy):
    mod.zero_grad(set_to_none=True)
    x = x.clone().requires_grad_(True)
    y = mod(x)
    y.backward(gy)
    out = {"y": y.detach(), "dx": x.grad, "dalpha": mod.alpha.grad, "dbias": mod.bias.grad}
    if mod.weight.grad is not None:
        out["dw"] = mod.weight.grad
    return out


def _layer_case(kind, m_out, k_in, tokens, ratio=None, sparsity=0.3, seed=0):
    ref = load_reference_atq()
    torch.manual_seed(seed)
    if kind == "ternary":
        r = ref.layers.TernaryLinear(k_in, m_out)
        o = atq.TernaryLinear(k_in, m_out)
    else:
        r = ref.precision_boost.ResidualPrecisionBoostLinear(k_in, m_out, precision_ratio=ratio, sparsity_target=sparsity)
        o = atq.ResidualPrecisionBoostLinear(k_in, m_out, precision_ratio=ratio, sparsity_target=sparsity)
    with torch.no_grad():
        r.bias.uniform_(-0.05, 0.05)
        r.alpha.fill_(0.731)
    o.load_state_dict(r.state_dict())
    r.to(DEV)
    o.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(seed + 1)
    x = torch.randn(tokens, k_in, device=DEV, generator=g)
    gy = torch.randn(tokens, m_out, device=DEV, generator=g)
    ours = _run_layer(o, x, gy)
    # integer part: the ternary pattern must be the reference's, bit for bit
    s = sparsity if kind == "rpb" else 0.3
    t_ref, _ = ref.quantizers.adaptive_ternary_quantization(r.weight.detach(), r.alpha, 0.05, s)
    t_our, _ = atq.adaptive_ternary_quantization(o.weight.detach(), o.alpha, 0.05, s)
    assert torch.equal(t_ref, t_our), "ternary pattern differs from the reference"
    del t_ref, t_our
    f32 = _run_layer(r, x, gy)
    r.double()
    f64 = _run_layer(r, x.double(), gy.double())
    if kind == "ternary":
        assert o.weight.grad is None and "dw" not in f32  # no gradient reaches TernaryLinear.weight
    else:
        assert float((ours["dw"] * (1 - o.precision_mask)).abs().max()) == 0.0
    for key in f64:
        _check(f"{kind} {m_out}x{k_in} @ {tokens} tokens: {key}", ours[key], f32[key], f64[key], relative=(key != "y"))


@pytest.mark.parametrize("kind,ratio", [("ternary", None), ("rpb", 0.05), ("rpb", 0.2)])
def test_config3_8192_square_16k_tokens(kind, ratio):
    """BASELINE config 3 upper size: 8192 x 8192 weights, 16 384 tokens, TernaryLinear and RPB 0.05 / 0.2."""
    _layer_case(kind, 8192, 8192, 16384, ratio)


@pytest.mark.parametrize("kind,ratio", [("ternary", None), ("rpb", 0.2)])
def test_config3_4096_square_64k_tokens(kind, ratio):
    """BASELINE config 3 token ceiling: 4096 x 4096 weights, 65 536 tokens."""
    _layer_case(kind, 4096, 4096, 65536, ratio)


@pytest.mark.parametrize("m_out,k_in,ratio", [(768, 768, 0.4), (3072, 768, 0.2), (768, 3072, 0.4)])
def test_config4_layer_shapes_full_token_count(m_out, k_in, ratio):
    """BASELINE config 4 per-GPU token count: 512 images x 197 tokens = 100 864; the [768,768] dW is the split-K case."""
    _layer_case("rpb", m_out, k_in, 100864, ratio, sparsity=0.13125)
