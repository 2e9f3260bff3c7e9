"""GPU: out-of-bounds canaries (compute-sanitizer is closed on this GPU pool, so the memcheck role is played by guard
regions around every output buffer of the C-ABI calls most exposed to ragged shapes: nothing outside the documented
extent may change)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

import atq._engine as eng
import atq._native as nv

DEV = "cuda:0"
GUARD = 64  # elements on each side


def _guarded(n, dtype, fill):
    buf = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device=DEV)
    return buf, buf[GUARD: GUARD + n]


def _intact(buf, n, fill):
    return bool((buf[:GUARD] == fill).all()) and bool((buf[GUARD + n:] == fill).all())


@pytest.mark.parametrize("n", [1, 3, 4, 2047, 2048, 2049, 4097, 100003])
def test_batched_codec_and_split_stay_inside_their_buffers(n):
    dev = 0
    st = nv.stream_ptr(dev)
    w = ((torch.rand(n + 16, device=DEV) * 2 - 1) / 4)[:n]   # 16-byte aligned start, ragged length
    thr = torch.tensor(0.1, device=DEV)
    pbuf, packed = _guarded((n + 3) // 4, torch.uint8, 0xAB)
    # slicing at GUARD = 64 bytes keeps 16-byte alignment
    arr = lambda ts: (ctypes.c_void_p * 1)(*[t.data_ptr() for t in ts])
    ns = (ctypes.c_int64 * 1)(n)
    nv.call("atq_ternarize_pack2_batched", dev, 1, arr([w]), ns, arr([thr]), arr([packed]), st)
    obuf, out = _guarded(n, torch.float32, 7.0)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    nv.call("atq_unpack2_to_f32_batched", dev, 1, arr([packed]), ns, arr([out]), flag.data_ptr(), st)
    p2buf, packed2 = _guarded((n + 3) // 4, torch.uint8, 0xCD)
    nv.call("atq_pack2_from_f32_batched", dev, 1, arr([out]), ns, arr([packed2]), flag.data_ptr(), st)
    torch.cuda.synchronize()
    assert _intact(pbuf, (n + 3) // 4, 0xAB) and _intact(obuf, n, 7.0) and _intact(p2buf, (n + 3) // 4, 0xCD)
    assert torch.equal(packed, packed2) and int(flag) == 0
    assert torch.equal(packed, eng.ternarize_pack2(w.contiguous(), thr))


@pytest.mark.parametrize("rows,cols", [(1, 8), (37, 24), (800, 192), (300, 200), (2047, 192)])
def test_operand_split_and_gemm_output_stay_inside_their_buffers(rows, cols):
    dev = 0
    st = nv.stream_ptr(dev)
    x = torch.randn(rows, cols, device=DEV)
    pitch = nv.round_up(cols, 8)
    hbuf, hi = _guarded(rows * pitch, torch.float16, 9.0)
    lbuf, lo = _guarded(rows * pitch, torch.float16, 9.0)
    slot = nv.new_slot(x.device)
    nv.call("atq_absmax_scale", dev, x.data_ptr(), rows, cols, cols, 1.0, None, slot.data_ptr(), st)
    nv.call("atq_split_bf16", dev, x.data_ptr(), rows, cols, cols, hi.data_ptr(), lo.data_ptr(), pitch, slot.data_ptr(), st)
    torch.cuda.synchronize()
    assert _intact(hbuf, rows * pitch, 9.0) and _intact(lbuf, rows * pitch, 9.0)
    if pitch != cols:  # padding columns are documented as untouched
        assert bool((hi.view(rows, pitch)[:, cols:] == 9.0).all())
    # GEMM: out pitch wider than cols, guards around the whole output
    w = torch.randn(72, cols, device=DEV)
    wop = eng.split_operand(w)
    out_pitch = 80
    ybuf, y = _guarded(rows * out_pitch, torch.float32, -3.0)
    oa = nv.operand(hi.view(rows, pitch), lo.view(rows, pitch), pitch, 0, slot)
    ob = nv.operand(*wop)
    nv.call("atq_tgemm", dev, rows, 72, cols, ctypes.byref(oa), ctypes.byref(ob), None, None, y.data_ptr(), out_pitch, None, 0, None,
            None, 0, st)
    torch.cuda.synchronize()
    assert _intact(ybuf, rows * out_pitch, -3.0)
    y2 = y.view(rows, out_pitch)
    assert bool((y2[:, 72:] == -3.0).all())
    ref = x.double() @ w.double().t()
    assert torch.allclose(y2[:, :72].double(), ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
