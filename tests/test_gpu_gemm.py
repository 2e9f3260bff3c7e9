"""GPU parity: the tcgen05 GEMM (K7/K8/K9) and the layers' forward/backward vs the oracle.
Tolerance (BASELINE north_star): rtol 1e-2 / atol 1e-3 against the reference's fp32 path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import atq
import atq._engine as eng
from oracle import atq_oracle as O

DEV = "cuda:0"
TOL = dict(rtol=1e-2, atol=1e-3)


def _gemm_ref(a, b):
    return (a.double() @ b.double().t())


@pytest.mark.parametrize("rows,cols,k", [(128, 128, 64), (128, 256, 128), (256, 512, 256), (1, 1, 8), (16, 1, 96),
                                         (800, 192, 192), (800, 384, 192), (16, 192, 512), (300, 200, 104),
                                         (129, 257, 72), (1024, 4096, 4096), (257, 10, 128), (256, 128, 3136)])
def test_tgemm_hi_lo_terms(rows, cols, k):
    g = torch.Generator().manual_seed(rows * 7 + cols * 3 + k)
    a = torch.randn(rows, k, generator=g)
    b = torch.randn(cols, k, generator=g) / k ** 0.5
    ag, bg = a.to(DEV), b.to(DEV)
    ref = _gemm_ref(a, b)
    # all three term structures: (hi,lo)x(hi,lo), (hi,lo)x(hi), (hi)x(hi)
    a2, b2 = eng.split_bf16(ag, True), eng.split_bf16(bg, True)
    y, _ = eng.tgemm(a2, b2, rows, cols, k)
    assert torch.allclose(y.cpu().double(), ref, rtol=1e-3, atol=1e-4), (y.cpu().double() - ref).abs().max()
    # exact-in-bf16 B (ternary): two A terms
    t = torch.randint(-1, 2, (cols, k), generator=g).float()
    tb = eng.split_bf16(t.to(DEV), False)
    y, _ = eng.tgemm(a2, tb, rows, cols, k)
    assert torch.allclose(y.cpu().double(), _gemm_ref(a, t), rtol=1e-3, atol=1e-3)
    # single term: compare against the same bf16-rounded inputs (exact products, fp32 accumulate)
    a1 = eng.split_bf16(ag, False)
    y, _ = eng.tgemm(a1, tb, rows, cols, k)
    a_r = a1[0][:, :k].float().cpu()
    assert torch.allclose(y.cpu().double(), _gemm_ref(a_r, t), rtol=1e-4, atol=1e-3)


def test_tgemm_epilogues():
    g = torch.Generator().manual_seed(0)
    rows, cols, k = 300, 200, 136
    a = torch.randn(rows, k, generator=g)
    t = torch.randint(-1, 2, (cols, k), generator=g).float()
    bias = torch.randn(cols, generator=g)
    scale = torch.tensor([0.37])
    ref_x = torch.randn(rows, cols, generator=g)
    a2 = eng.split_bf16(a.to(DEV), True)
    tb = eng.split_bf16(t.to(DEV), False)
    y, dot = eng.tgemm(a2, tb, rows, cols, k, scale=scale.to(DEV), bias=bias.to(DEV), dot_ref=ref_x.to(DEV))
    acc = _gemm_ref(a, t)
    assert torch.allclose(y.cpu().double(), acc * 0.37 + bias.double(), **TOL)
    want = float((acc * ref_x.double()).sum())
    assert abs(float(dot) - want) <= 1e-3 * abs(want) + 1e-2
    # masked dW epilogue with the ternary d(alpha) reduction
    m_out, k_in, n_tok = 96, 200, 333
    dy = torch.randn(n_tok, m_out, generator=g)
    x = torch.randn(n_tok, k_in, generator=g)
    mask = (torch.rand(m_out, k_in, generator=g) < 0.2).float()
    tq = torch.randint(-1, 2, (m_out, k_in), generator=g).float()
    packed, _ = eng.pack2_from_f32(tq.to(DEV).reshape(-1))
    dw, dalpha = eng.tgemm_dw_masked(eng.split_bf16_t(dy.to(DEV), True), eng.split_bf16_t(x.to(DEV), True), m_out, k_in,
                                     n_tok, mask=mask.to(DEV), packed=packed)
    G = dy.double().t() @ x.double()
    assert torch.allclose(dw.cpu().double(), G * mask.double(), **TOL)
    assert int((dw.cpu() != 0).sum()) <= int(mask.sum())
    want = float((G * tq.double() * (1 - mask.double())).sum())
    assert abs(float(dalpha) - want) <= 1e-3 * abs(want) + 1e-2


def _pair(kind, k_in, m_out, ratio, s, bias=True, seed=0):
    torch.manual_seed(seed)
    if kind == "ternary":
        ref, mod = O.OracleTernaryLinear(k_in, m_out, bias), atq.TernaryLinear(k_in, m_out, bias)
    else:
        ref = O.OracleRPBLinear(k_in, m_out, ratio, bias, s)
        mod = atq.ResidualPrecisionBoostLinear(k_in, m_out, ratio, bias, s)
    with torch.no_grad():
        ref.alpha.fill_(0.8)
    mod.load_state_dict(ref.state_dict())
    return ref, mod.to(DEV)


@pytest.mark.parametrize("kind,k_in,m_out,ratio,s,xshape", [
    ("ternary", 64, 32, 0, 0.3, (5, 7)), ("ternary", 192, 192, 0, 0.3, (16, 50)), ("ternary", 96, 1, 0, 0.3, (16, 50)),
    ("ternary", 4096, 4096, 0, 0.3, (1024,)), ("ternary", 100, 30, 0, 0.3, (9,)),
    ("rpb", 64, 32, 0.05, 0.3, (11,)), ("rpb", 192, 192, 0.2, 0.1333, (16, 50)), ("rpb", 192, 384, 0.2, 0.1, (16, 50)),
    ("rpb", 384, 192, 0.4, 0.15, (16, 50)), ("rpb", 96, 1, 0.2, 0.1, (16, 50)), ("rpb", 512, 192, 0.2, 0.2, (16,)),
    ("rpb", 3136, 128, 0.05, 0.3, (256,)), ("rpb", 128, 10, 0.1, 0.05, (256,)), ("rpb", 1024, 2048, 0.05, 0.3, (2048,)),
])
def test_layer_forward_backward_vs_oracle(kind, k_in, m_out, ratio, s, xshape):
    ref, mod = _pair(kind, k_in, m_out, ratio, s)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*xshape, k_in, generator=g)
    gy = torch.randn(*xshape, m_out, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    yg = mod(xg)
    yg.backward(gy.to(DEV))
    assert yg.shape == yr.shape
    assert torch.allclose(yg.detach().cpu(), yr.detach(), **TOL)
    assert torch.allclose(xg.grad.cpu(), xr.grad, **TOL)
    # reductions over N*M*K terms: same rtol, atol scaled by the magnitude of the summands
    n_tok = x.numel() // k_in
    assert torch.allclose(mod.alpha.grad.cpu(), ref.alpha.grad, rtol=1e-2, atol=1e-3 * max(1.0, (n_tok * m_out) ** 0.5))
    assert torch.allclose(mod.bias.grad.cpu(), ref.bias.grad, **TOL)
    if kind == "ternary":
        assert mod.weight.grad is None  # the reference's gradient contract
    else:
        assert torch.allclose(mod.weight.grad.cpu(), ref.weight.grad, **TOL)
        off_mask = mod.weight.grad.cpu()[ref.precision_mask == 0]
        assert float(off_mask.abs().max()) == 0.0 if off_mask.numel() else True
        assert int((mod.weight.grad != 0).sum()) <= int(ref.precision_mask.sum())


def test_layers_golden(golden, policy):
    tl = atq.TernaryLinear(64, 32)
    tl.load_state_dict({"weight": torch.from_numpy(golden["tl_weight"]), "bias": torch.from_numpy(golden["tl_bias"]),
                        "alpha": torch.from_numpy(golden["tl_alpha"])})
    tl.to(DEV)
    x = torch.from_numpy(golden["tl_x"]).to(DEV).requires_grad_(True)
    y = tl(x)
    y.backward(torch.from_numpy(golden["tl_gy"]).to(DEV))
    assert tl.weight.grad is None
    assert torch.allclose(y.detach().cpu(), torch.from_numpy(golden["tl_y"]), **TOL)
    assert torch.allclose(x.grad.cpu(), torch.from_numpy(golden["tl_dx"]), **TOL)
    assert torch.allclose(tl.alpha.grad.cpu(), torch.from_numpy(golden["tl_dalpha"]), rtol=1e-2, atol=1e-2)
    assert torch.allclose(tl.bias.grad.cpu(), torch.from_numpy(golden["tl_dbias"]), **TOL)

    for pre, (k_in, m_out, ratio, s, bias) in {"rpb": (64, 32, 0.05, 0.3, True), "rpb2": (96, 40, 0.4, 0.1333, False)}.items():
        rpb = atq.ResidualPrecisionBoostLinear(k_in, m_out, ratio, bias, s)
        sd = {"weight": torch.from_numpy(golden[f"{pre}_weight"]), "alpha": torch.from_numpy(golden[f"{pre}_alpha"]),
              "precision_mask": torch.from_numpy(golden[f"{pre}_mask"])}
        if bias:
            sd["bias"] = torch.from_numpy(golden[f"{pre}_bias"])
        rpb.load_state_dict(sd)
        rpb.to(DEV)
        x = torch.from_numpy(golden[f"{pre}_x"]).to(DEV).requires_grad_(True)
        y = rpb(x)
        y.backward(torch.from_numpy(golden[f"{pre}_gy"]).to(DEV))
        assert torch.allclose(y.detach().cpu(), torch.from_numpy(golden[f"{pre}_y"]), **TOL)
        assert torch.allclose(x.grad.cpu(), torch.from_numpy(golden[f"{pre}_dx"]), **TOL)
        assert torch.allclose(rpb.weight.grad.cpu(), torch.from_numpy(golden[f"{pre}_dw"]), **TOL)
        assert torch.allclose(rpb.alpha.grad.cpu(), torch.from_numpy(golden[f"{pre}_dalpha"]), rtol=1e-2, atol=1e-2)
        if pre == "rpb":
            assert int((rpb.weight.grad != 0).sum()) == policy["rpb_dw_nonzeros"]
            t, a = rpb.get_quantized_weights()
            assert a is rpb.alpha
            assert np.array_equal(t.cpu().numpy().astype(np.int8), golden["rpb_tq"])


def test_cache_invalidation_and_sparsity_attribute():
    ref, mod = _pair("rpb", 128, 64, 0.1, 0.1)
    x = torch.randn(33, 128)
    xg = x.to(DEV)
    with torch.no_grad():
        y1 = mod(xg)
        assert torch.allclose(y1.cpu(), ref(x), **TOL)
        # external scheduler assigns a new sparsity target: next forward must see it
        mod.sparsity_target = ref.sparsity_target = 0.45
        assert torch.allclose(mod(xg).cpu(), ref(x), **TOL)
        # in-place weight update (optimizer step / re-init) invalidates the cache
        ref.weight.mul_(-0.5).add_(0.01)
        mod.weight.copy_(ref.weight)
        assert torch.allclose(mod(xg).cpu(), ref(x), **TOL)
        ref.alpha.fill_(1.7)
        mod.alpha.fill_(1.7)
        assert torch.allclose(mod(xg).cpu(), ref(x), **TOL)
        # precision_ratio writes are inert
        mod.precision_ratio = 0.9
        assert torch.allclose(mod(xg).cpu(), ref(x), **TOL)


def test_training_steps_track_oracle():
    """A few Adam steps on both sides stay within tolerance (quantizer re-run every step)."""
    torch.manual_seed(3)
    ref = torch.nn.Sequential(O.OracleRPBLinear(96, 64, 0.1, True, 0.2), torch.nn.ReLU(), O.OracleTernaryLinear(64, 10))
    mod = torch.nn.Sequential(atq.ResidualPrecisionBoostLinear(96, 64, 0.1, True, 0.2), torch.nn.ReLU(), atq.TernaryLinear(64, 10))
    mod.load_state_dict(ref.state_dict())
    mod.to(DEV)
    o_r = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-4)
    o_g = torch.optim.Adam(mod.parameters(), lr=1e-3, weight_decay=1e-4)
    x = torch.randn(64, 96)
    y = torch.randint(0, 10, (64,))
    for _ in range(5):
        o_r.zero_grad(); o_g.zero_grad()
        lr = torch.nn.functional.cross_entropy(ref(x), y)
        lg = torch.nn.functional.cross_entropy(mod(x.to(DEV)), y.to(DEV))
        lr.backward(); lg.backward()
        o_r.step(); o_g.step()
        assert abs(float(lr) - float(lg)) < 1e-3
    assert mod[2].weight.grad is None and ref[2].weight.grad is None
    assert torch.allclose(mod[0].weight.detach().cpu(), ref[0].weight.detach(), rtol=1e-3, atol=1e-4)


def test_fast_mode_is_bf16_single_term():
    ref, mod = _pair("ternary", 256, 128, 0, 0.3)
    x = torch.randn(64, 256)
    atq.set_gemm_mode("fast")
    try:
        with torch.no_grad():
            y = mod(x.to(DEV)).cpu()
    finally:
        atq.set_gemm_mode("parity")
    xb = x.bfloat16().float()
    assert torch.allclose(y, ref(xb).detach(), rtol=1e-3, atol=1e-3)


def test_prepare_quantization_batched_equals_lazy():
    torch.manual_seed(5)
    ref = torch.nn.Sequential(O.OracleRPBLinear(192, 96, 0.2, True, 0.1333), torch.nn.Tanh(), O.OracleRPBLinear(96, 1, 0.2, True, 0.1),
                              torch.nn.Tanh(), O.OracleTernaryLinear(1, 8))
    mod = torch.nn.Sequential(atq.ResidualPrecisionBoostLinear(192, 96, 0.2, True, 0.1333), torch.nn.Tanh(),
                              atq.ResidualPrecisionBoostLinear(96, 1, 0.2, True, 0.1), torch.nn.Tanh(), atq.TernaryLinear(1, 8))
    mod.load_state_dict(ref.state_dict())
    mod.to(DEV)
    x = torch.randn(40, 192)
    assert atq.prepare_quantization(mod) == 3
    assert atq.prepare_quantization(mod) == 0  # nothing stale
    with torch.no_grad():
        assert torch.allclose(mod(x.to(DEV)).cpu(), ref(x), **TOL)
        mod[0].sparsity_target = ref[0].sparsity_target = 0.4
        ref[2].weight.mul_(1.5)
        mod[2].weight.copy_(ref[2].weight)
        assert atq.prepare_quantization(mod) == 2
        assert torch.allclose(mod(x.to(DEV)).cpu(), ref(x), **TOL)


@pytest.mark.parametrize("rows,cols,k", [(128, 128, 64), (300, 200, 128), (16, 1, 64), (800, 192, 192), (257, 10, 128),
                                         (1000, 384, 3136), (1024, 4096, 4096), (130, 520, 704)])
def test_tgemm_packed_b_unpacked_in_smem(rows, cols, k):
    """K7 with B given as 2-bit codec bytes: bit-for-bit the same accumulation as the bf16-B path
    (T is exact in bf16), and within tolerance of the fp64 answer."""
    g = torch.Generator().manual_seed(rows + cols + k)
    a = torch.randn(rows, k, generator=g)
    t = torch.randint(-1, 2, (cols, k), generator=g).float()
    packed, flag = eng.pack2_from_f32(t.to(DEV).reshape(-1))
    assert int(flag) == 0 and eng.packed_gemm_ok(k, packed)
    scale = torch.tensor([0.9], device=DEV)
    bias = torch.randn(cols, generator=g).to(DEV)
    ref_x = torch.randn(rows, cols, generator=g).to(DEV)
    tb = eng.split_bf16(t.to(DEV), False)
    for want_lo in (True, False):
        a_op = eng.split_bf16(a.to(DEV), want_lo)
        y_p, d_p = eng.tgemm_packed(a_op, packed, rows, cols, k, scale=scale, bias=bias, dot_ref=ref_x)
        y_b, d_b = eng.tgemm(a_op, tb, rows, cols, k, scale=scale, bias=bias, dot_ref=ref_x)
        assert torch.equal(y_p, y_b)
        assert torch.allclose(d_p, d_b, rtol=1e-5, atol=1e-4)
    want = _gemm_ref(a, t) * 0.9 + bias.cpu().double()
    a_bf = a.bfloat16().float()  # last loop iteration: single bf16 term -> exact products of the rounded A
    assert torch.allclose(y_p.cpu().double(), _gemm_ref(a_bf, t) * 0.9 + bias.cpu().double(), rtol=1e-4, atol=1e-3)
    a2 = eng.split_bf16(a.to(DEV), True)
    y, _ = eng.tgemm_packed(a2, packed, rows, cols, k, scale=scale, bias=bias)
    assert torch.allclose(y.cpu().double(), want, **TOL)


@pytest.mark.parametrize("policy", ["always", "never"])
def test_ternary_linear_packed_policy(policy):
    """TernaryLinear forward/backward through the packed-B kernels and through the bf16-TMA kernels agree
    with the oracle (and with each other bit for bit)."""
    ref, mod = _pair("ternary", 192, 128, 0, 0.3)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(16, 50, 192, generator=g)
    gy = torch.randn(16, 50, 128, generator=g)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(gy)
    atq.set_packed_gemm(policy)
    try:
        xg = x.to(DEV).requires_grad_(True)
        y = mod(xg)
        y.backward(gy.to(DEV))
    finally:
        atq.set_packed_gemm("auto")
    assert torch.allclose(y.detach().cpu(), ref(x).detach(), **TOL)
    assert torch.allclose(xg.grad.cpu(), xr.grad, **TOL)
    assert torch.allclose(mod.alpha.grad.cpu(), ref.alpha.grad, rtol=1e-2, atol=1e-1)
    assert mod.weight.grad is None


@pytest.mark.parametrize("rows,cols,k", [(128, 128, 64), (300, 200, 136), (16, 1, 96), (800, 192, 192), (192, 384, 800),
                                         (257, 10, 128), (96, 192, 16), (1000, 520, 1064), (4096, 4096, 1024),
                                         (192, 192, 4096), (768, 768, 8200), (1, 96, 2048)])
def test_tgemm_mn_major_operands(rows, cols, k):
    """dX reads B as [k, cols] row-major, dW reads A as [k, rows] and B as [k, cols] row-major (MN-major UMMA
    descriptors, no transposes): same result as the K-major path on explicitly transposed copies."""
    g = torch.Generator().manual_seed(rows * 3 + cols + k)
    a = torch.randn(rows, k, generator=g)
    b = torch.randn(cols, k, generator=g) / k ** 0.5
    ref = _gemm_ref(a, b)
    a_k = eng.split_bf16(a.to(DEV), True)                       # [rows, k]  K-major
    a_mn = eng.split_bf16(a.t().contiguous().to(DEV), True)     # [k, rows]  MN-major memory
    b_k = eng.split_bf16(b.to(DEV), True)
    b_mn = eng.split_bf16(b.t().contiguous().to(DEV), True)     # [k, cols]
    y_kk, _ = eng.tgemm(a_k, b_k, rows, cols, k)
    y_km, _ = eng.tgemm(a_k, b_mn + (1,), rows, cols, k)
    assert torch.allclose(y_km.cpu().double(), ref, rtol=1e-3, atol=1e-4), (y_km.cpu().double() - ref).abs().max()
    assert torch.allclose(y_km, y_kk, rtol=1e-5, atol=1e-5)
    mask = (torch.rand(rows, cols, generator=g) < 0.3).float().to(DEV)
    y_mm, _ = eng.tgemm_dw_masked(a_mn + (1,), b_mn + (1,), rows, cols, k, mask=mask)
    assert torch.allclose(y_mm.cpu().double(), ref * mask.cpu().double(), rtol=1e-3, atol=1e-4)
    # single-term variants
    y1, _ = eng.tgemm((a_k[0], None, a_k[2]), (b_mn[0], None, b_mn[2], 1), rows, cols, k)
    want1 = _gemm_ref(a_k[0][:, :k].float().cpu(), b_mn[0][:, :cols].float().cpu().t())
    assert torch.allclose(y1.cpu().double(), want1, rtol=1e-4, atol=1e-3)
    y2, _ = eng.tgemm_dw_masked((a_mn[0], None, a_mn[2], 1), (b_mn[0], None, b_mn[2], 1), rows, cols, k)
    want2 = _gemm_ref(a_mn[0][:, :rows].float().cpu().t(), b_mn[0][:, :cols].float().cpu().t())
    assert torch.allclose(y2.cpu().double(), want2, rtol=1e-4, atol=1e-3)


def test_fused_optimizer_step_invalidates_cache():
    """torch's fused AdamW updates weights without bumping Tensor._version; the optimizer post-step
    hook must still force re-quantization on the next forward."""
    torch.manual_seed(9)
    ref = O.OracleRPBLinear(128, 64, 0.1, True, 0.2)
    mod = atq.ResidualPrecisionBoostLinear(128, 64, 0.1, True, 0.2)
    mod.load_state_dict(ref.state_dict())
    mod.to(DEV)
    o_r = torch.optim.AdamW(ref.parameters(), lr=5e-2, weight_decay=1e-2)
    o_g = torch.optim.AdamW(mod.parameters(), lr=5e-2, weight_decay=1e-2, fused=True)
    x = torch.randn(32, 128)
    for _ in range(3):
        o_r.zero_grad(); o_g.zero_grad()
        ref(x).square().mean().backward()
        mod(x.to(DEV)).square().mean().backward()
        o_r.step(); o_g.step()
    with torch.no_grad():
        assert torch.allclose(mod(x.to(DEV)).cpu(), ref(x), rtol=1e-2, atol=2e-3)
        assert torch.allclose(mod.weight.cpu(), ref.weight, rtol=1e-3, atol=1e-4)


def test_dw_split_k_masked_epilogue_long_contraction():
    """dW with few output tiles and a long token dimension takes the split-K path (per-range slabs + fixed-order
    finalize): mask, d(alpha) reduction and determinism."""
    g = torch.Generator().manual_seed(4)
    m_out, k_in, n_tok = 192, 384, 6000
    dy = torch.randn(n_tok, m_out, generator=g)
    x = torch.randn(n_tok, k_in, generator=g)
    mask = (torch.rand(m_out, k_in, generator=g) < 0.2).float()
    tq = torch.randint(-1, 2, (m_out, k_in), generator=g).float()
    packed, _ = eng.pack2_from_f32(tq.to(DEV).reshape(-1))
    ga, xa = eng.split_bf16(dy.to(DEV), True), eng.split_bf16(x.to(DEV), True)
    dw, dalpha = eng.tgemm_dw_masked(ga + (1,), xa + (1,), m_out, k_in, n_tok, mask=mask.to(DEV), packed=packed)
    G = dy.double().t() @ x.double()
    assert torch.allclose(dw.cpu().double(), G * mask.double(), **TOL)
    want = float((G * tq.double() * (1 - mask.double())).sum())
    assert abs(float(dalpha) - want) <= 1e-3 * abs(want) + 5e-2
    dw2, dalpha2 = eng.tgemm_dw_masked(ga + (1,), xa + (1,), m_out, k_in, n_tok, mask=mask.to(DEV), packed=packed)
    assert torch.equal(dw, dw2) and torch.equal(dalpha, dalpha2)  # bitwise reproducible


def test_frozen_ternary_linear_is_quantized_once():
    """TernaryLinear.weight never receives a gradient, so no optimizer step can change it: its operands
    are built once, not once per step; a gradient reaching the weight (L1 regularisation, train.py:195-203) or a new
    Parameter re-arms the per-step invalidation."""
    import atq._native as nv
    torch.manual_seed(0)
    net = torch.nn.Sequential(atq.TernaryLinear(64, 48), torch.nn.ReLU(), atq.TernaryLinear(48, 16)).to(DEV)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-2)
    x = torch.randn(32, 64, device=DEV)
    builds = []
    orig = nv.call

    def counting(name, *a):
        if name == "atq_build_ternary_operands":
            builds.append(name)
        return orig(name, *a)
    nv.call = counting
    try:
        for _ in range(4):
            opt.zero_grad(set_to_none=True)
            net(x).square().mean().backward()
            opt.step()
        assert len(builds) == 2, builds            # one build per layer, ever
        assert net[0].weight.grad is None
        # a gradient reaches the weight: from now on every optimizer step invalidates the operands
        opt.zero_grad(set_to_none=True)
        (net(x).square().mean() + 1e-3 * net[0].weight.abs().sum()).backward()
        opt.step()
        n0 = len(builds)
        net(x)
        assert len(builds) == n0 + 1
    finally:
        nv.call = orig


@pytest.mark.parametrize("rows,cols,k", [(256, 512, 256), (800, 192, 192), (300, 200, 104), (257, 130, 72), (513, 129, 200)])
def test_cta_pair_kernels_on_ragged_shapes(rows, cols, k):
    """The cta_group::2 kernels are only selected from one wave of tiles on (smaller GEMMs are latency bound); bit 3 of
    atq_set_cta_pairs forces them, so their handling of ragged edges (rows % 256, cols % 128, k % 64) stays covered."""
    import atq._native as nv
    g = torch.Generator().manual_seed(rows + 3 * cols + 7 * k)
    a = torch.randn(rows, k, generator=g)
    b = torch.randn(cols, k, generator=g) / k ** 0.5
    t = torch.randint(-1, 2, (cols, k), generator=g).float()
    old = nv.lib.atq_set_cta_pairs(1 | 2 | 8)
    try:
        a2, b2 = eng.split_bf16(a.to(DEV), True), eng.split_bf16(b.to(DEV), True)
        y, _ = eng.tgemm(a2, b2, rows, cols, k)
        assert torch.allclose(y.cpu().double(), _gemm_ref(a, b), rtol=1e-3, atol=1e-4)
        y, _ = eng.tgemm(a2, eng.split_bf16(t.to(DEV), False), rows, cols, k)
        assert torch.allclose(y.cpu().double(), _gemm_ref(a, t), rtol=1e-3, atol=1e-3)
        # masked dW (MN-major operands read in place): out x in = rows x cols over k tokens
        dy = torch.randn(k, rows, generator=g)
        x = torch.randn(k, cols, generator=g)
        mask = (torch.rand(rows, cols, generator=g) < 0.2).float()
        dw, _ = eng.tgemm_dw_masked(eng.mn_view(eng.split_bf16(dy.to(DEV), True) + (0, None)),
                                    eng.mn_view(eng.split_bf16(x.to(DEV), True) + (0, None)), rows, cols, k, mask=mask.to(DEV))
        assert torch.allclose(dw.cpu().double(), (dy.double().t() @ x.double()) * mask.double(), **TOL)
    finally:
        nv.lib.atq_set_cta_pairs(old)
