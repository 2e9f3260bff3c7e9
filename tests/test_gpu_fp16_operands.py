"""GPU: the scaled-fp16 operand format of the default ("parity") GEMM mode -- per-tensor power-of-two scale from a
grid-level |x| max reduction, (hi, lo) fp16 pairs, 1/scale applied in the GEMM epilogue -- against float64."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import atq
import atq._engine as eng

DEV = "cuda:0"


def _ref(a, b):
    return a.double() @ b.double().t()


@pytest.mark.parametrize("shape,mag", [((800, 192), 1.0), ((17, 33), 3e-7), ((4096, 768), 250.0), ((5, 8), 6e4), ((64, 64), 1e-30)])
def test_absmax_scale_slot(shape, mag):
    g = torch.Generator().manual_seed(shape[0])
    x = (torch.randn(*shape, generator=g) * mag).to(DEV)
    slot = eng.absmax_slot(x)
    torch.cuda.synchronize()
    bits, s, inv, ticket = slot.view(torch.int32)[0].item(), float(slot[1]), float(slot[2]), slot.view(torch.int32)[3].item()
    assert bits == 0 and ticket == 0, "slot must be re-armed when the kernel has finished"
    m = float(x.abs().max())
    assert math.log2(s) == int(math.log2(s)) and s * inv == 1.0
    assert 2.0 ** 14 <= m * s < 2.0 ** 15
    # bound_mul and the extra scalar raise the bound
    extra = torch.tensor([m * 10.0], device=DEV)
    slot2 = eng.absmax_slot(x, 3.0, extra)
    assert 2.0 ** 14 <= 30.0 * m * float(slot2[1]) < 2.0 ** 15
    # strided rows (not contiguous): only the addressed elements count
    big = torch.zeros(shape[0], shape[1] + 5, device=DEV)
    big[:, shape[1]:] = 1e30
    big[:, :shape[1]] = x
    slot3 = eng.absmax_slot(big[:, :shape[1]])
    assert float(slot3[1]) == s


def test_absmax_degenerate_inputs():
    for fill in (0.0, float("inf"), float("nan")):
        x = torch.full((64, 32), fill, device=DEV)
        slot = eng.absmax_slot(x)
        assert float(slot[1]) == 1.0 and float(slot[2]) == 1.0
    x = torch.randn(64, 32, device=DEV)
    x[3, 3] = float("nan")  # NaNs are skipped by the reduction; the operand itself still carries them
    slot = eng.absmax_slot(x)
    assert 2.0 ** 14 <= float(x[~x.isnan()].abs().max()) * float(slot[1]) < 2.0 ** 15


def test_split_reconstructs_22_bits():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(300, 200, generator=g) * torch.logspace(-6, 0, 200)  # 6 decades of dynamic range per row
    op = eng.split_operand(x.to(DEV))
    hi, lo, pitch, mn, slot = op
    assert hi.dtype == torch.float16 and slot is not None
    rec = (hi.double() + lo.double()) * float(slot[2])
    err = (rec.cpu() - x.double()).abs()
    m = float(x.abs().max())
    assert float((err - (x.double().abs() * 2.0 ** -21 + m * 2.0 ** -38)).max()) <= 0.0


@pytest.mark.parametrize("rows,cols,k", [(128, 128, 64), (300, 200, 104), (800, 192, 192), (1024, 1024, 4096), (16, 1, 96)])
def test_tgemm_scaled_fp16_terms(rows, cols, k):
    g = torch.Generator().manual_seed(rows + cols + k)
    a = torch.randn(rows, k, generator=g) * 37.0
    b = torch.randn(cols, k, generator=g) / k ** 0.5 * 1e-3
    ref = _ref(a, b)
    a2, b2 = eng.split_operand(a.to(DEV)), eng.split_operand(b.to(DEV))
    y, _ = eng.tgemm(a2, b2, rows, cols, k)
    scale = float(ref.abs().max())
    err = float((y.cpu().double() - ref).abs().max())
    # operand representation ~2^-22; fp32 accumulation in TMEM truncates once per tcgen05.mma of the hi x hi pass
    # (k/16 roundings at full accumulator size, ~4e-8 relative each)
    tol = (2e-6 + 4e-8 * (k / 16)) * scale
    assert err <= tol, (err, tol)
    # MN-major consumption of the same memory (dX / dW layouts)
    at, bt = eng.split_operand(a.t().contiguous().to(DEV)), eng.split_operand(b.t().contiguous().to(DEV))
    y_km, _ = eng.tgemm(a2, eng.mn_view(bt), rows, cols, k)
    y_mm, _ = eng.tgemm_dw_masked(eng.mn_view(at), eng.mn_view(bt), rows, cols, k)
    assert float((y_km.cpu().double() - ref).abs().max()) <= tol
    assert float((y_mm.cpu().double() - ref).abs().max()) <= tol


def test_format_mismatch_is_rejected():
    a = torch.randn(64, 64, device=DEV)
    a16 = eng.split_operand(a)
    abf = eng.split_bf16(a, True)
    with pytest.raises(RuntimeError, match="same element format"):
        eng.tgemm(a16, abf, 64, 64, 64)


def test_packed_b_with_fp16_activations():
    """The 2-bit codec bytes expanded to fp16 +-1/0 tiles in shared memory == the fp16 copy of T through TMA."""
    g = torch.Generator().manual_seed(5)
    rows, cols, k = 200, 384, 256
    t = torch.randint(-1, 2, (cols, k), generator=g).float().to(DEV)
    a = torch.randn(rows, k, generator=g).to(DEV)
    packed, _ = eng.pack2_from_f32(t.reshape(-1))
    a2 = eng.split_operand(a)
    tb = eng.split_bf16(t, False, atq._native.unit_slot(t.device))
    y_p, _ = eng.tgemm_packed(a2, packed, rows, cols, k)
    y_b, _ = eng.tgemm(a2, tb, rows, cols, k)
    assert torch.equal(y_p, y_b)
    ref = _ref(a.cpu(), t.cpu())
    assert float((y_p.cpu().double() - ref).abs().max()) <= 4e-6 * float(ref.abs().max())


def test_scale_follows_the_data_under_graph_replay():
    """absmax + split + GEMM captured once; replays with inputs of very different magnitude must re-derive the scale."""
    x = torch.randn(256, 128, device=DEV)
    w = torch.randn(64, 128, device=DEV)
    static_x = x.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            eng.tgemm(eng.split_operand(static_x), eng.split_operand(w), 256, 64, 128)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y, _ = eng.tgemm(eng.split_operand(static_x), eng.split_operand(w), 256, 64, 128)
    for mag in (1.0, 1e-5, 3e4, 1.0):
        static_x.copy_(x * mag)
        graph.replay()
        ref = _ref((x * mag).cpu(), w.cpu())
        assert float((y.cpu().double() - ref).abs().max()) <= 4e-6 * float(ref.abs().max()), mag


@pytest.mark.parametrize("mode", ["parity", "parity_bf16", "fast"])
def test_layers_in_every_mode(mode):
    from oracle import atq_oracle as O
    prev = atq.get_gemm_mode()
    atq.set_gemm_mode(mode)
    try:
        torch.manual_seed(3)
        ref = O.OracleRPBLinear(192, 96, 0.2, True, 0.15)
        mod = atq.ResidualPrecisionBoostLinear(192, 96, 0.2, True, 0.15)
        mod.load_state_dict(ref.state_dict())
        mod.to(DEV)
        x, gy = torch.randn(300, 192), torch.randn(300, 96)
        xr = x.clone().requires_grad_(True)
        ref(xr).backward(gy)
        xg = x.to(DEV).requires_grad_(True)
        mod(xg).backward(gy.to(DEV))
        tol = dict(rtol=1e-2, atol=1e-3) if mode != "fast" else dict(rtol=5e-2, atol=5e-2)
        assert torch.allclose(xg.grad.cpu(), xr.grad, **tol)
        assert torch.allclose(mod.weight.grad.cpu(), ref.weight.grad, **(tol if mode != "fast" else dict(rtol=5e-2, atol=2e-1)))
        if mode == "parity":  # the scaled-fp16 path is two orders of magnitude tighter than the stated tolerance
            assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-5)
            assert torch.allclose(mod.weight.grad.cpu(), ref.weight.grad, rtol=1e-4, atol=2e-5)
    finally:
        atq.set_gemm_mode(prev)
