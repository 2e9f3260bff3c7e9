"""GPU parity: threshold (K1/K2), ternarize (K3/K4), codec (K5/K6), routing (K10) -- bit-exact vs the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import atq
import atq._engine as eng
from atq.bit_packing import TernaryBitPacking
from oracle import atq_oracle as O

DEV = "cuda:0"


def _unbits(bits, shape):
    n = int(np.prod(shape))
    return (np.unpackbits(bits[0])[:n].astype(np.int8) - np.unpackbits(bits[1])[:n].astype(np.int8)).reshape(shape)


def _weights(shape, seed, kind):
    g = torch.Generator().manual_seed(seed)
    n = int(np.prod(shape))
    if kind == "uniform":
        w = (torch.rand(shape, generator=g) * 2 - 1) / max(1, shape[-1]) ** 0.5
    elif kind == "normal":
        w = torch.randn(shape, generator=g)
    elif kind == "ties":  # heavy duplicates: values on a coarse grid
        w = torch.randint(-8, 9, shape, generator=g).float() * 0.125
    elif kind == "wide":  # exponents spread over many binades, some denormals / zeros
        w = torch.randn(shape, generator=g) * torch.pow(10.0, torch.randint(-30, 5, shape, generator=g).float())
        w.view(-1)[:: 13] = 0.0
        w.view(-1)[1:: 17] = 1e-42
    return w


@pytest.mark.parametrize("shape", [(1, 96), (10, 128), (128, 3136), (192, 192), (33, 65), (1000, 1003), (4096, 1024), (7,)])
@pytest.mark.parametrize("kind", ["uniform", "normal", "ties", "wide"])
def test_threshold_and_ternarize_bit_exact(shape, kind):
    w = _weights(shape, hash((shape, kind)) % 9973, kind)
    wg = w.to(DEV)
    for s in (0.3, 0.05, 0.1333, 0.5, 0.999):
        thr_o = O.adaptive_threshold(w.numpy(), s)
        thr_g = eng.adaptive_threshold(wg, s)
        assert np.float32(thr_g.item()).tobytes() == np.float32(thr_o).tobytes(), (shape, kind, s)
        t_o = O.ternarize(w.numpy(), thr_o)
        t_g, a = atq.adaptive_ternary_quantization(wg, None, 0.05, s)
        assert t_g.dtype == torch.float32 and t_g.shape == wg.shape and not t_g.requires_grad
        assert np.array_equal(t_g.cpu().numpy().astype(np.int8), t_o)
        a_o = O.optimal_alpha(w.numpy(), t_o)
        assert abs(float(a) - float(a_o)) <= 2e-6 * max(abs(float(a_o)), 1e-30) + 1e-38
        packed = eng.ternarize_pack2(wg, thr_g)
        assert np.array_equal(packed.cpu().numpy(), O.pack2(t_o))


def test_threshold_edge_branches():
    w = _weights((50, 20), 3, "normal")
    wg = w.to(DEV)
    # k >= n: max + 1 -> everything zero; alpha* = mean|W|
    t, a = atq.adaptive_ternary_quantization(wg, None, 0.05, 1.0)
    assert float(t.abs().sum()) == 0.0
    assert abs(float(a) - float(w.abs().double().mean())) < 1e-6
    thr = eng.adaptive_threshold(wg, 1.0)
    assert float(thr) == float(np.float32(w.abs().max().item()) + np.float32(1.0))
    # k == 0: threshold_factor * mean|W| (fp32 reduction: compare within 1 ulp, T given that thr)
    thr = eng.adaptive_threshold(wg, 0.0, 0.05)
    want = O.adaptive_threshold(w.numpy(), 0.0, 0.05)
    assert abs(float(thr) - float(want)) <= 2 * np.spacing(np.float32(want))
    t, _ = atq.adaptive_ternary_quantization(wg, None, 0.05, 0.0)
    assert np.array_equal(t.cpu().numpy().astype(np.int8), O.ternarize(w.numpy(), np.float32(thr.item())))
    # tiny n where int(s*n) == 0 although s > 0
    w3 = torch.tensor([0.3, -0.2, 0.1])
    t, _ = atq.adaptive_ternary_quantization(w3.to(DEV), None, 0.05, 0.3)
    t_o, _, _ = O.adaptive_ternary_quantization(w3.numpy(), None, 0.05, 0.3)
    assert np.array_equal(t.cpu().numpy().astype(np.int8), t_o)


def test_alpha_passthrough_and_nan():
    wg = torch.randn(64, 64, device=DEV)
    alpha = torch.nn.Parameter(torch.tensor([1.25], device=DEV))
    t, a = atq.adaptive_ternary_quantization(wg, alpha)
    assert a is alpha
    w = torch.randn(40, 40)
    w[3, 4] = float("nan")
    w[5, 6] = float("inf")
    w[7, 8] = -float("inf")
    t, _ = atq.adaptive_ternary_quantization(w.to(DEV), None, 0.05, 0.3)
    t_o, _, _ = O.adaptive_ternary_quantization(w.numpy(), None, 0.05, 0.3)
    assert np.array_equal(t.cpu().numpy().astype(np.int8), t_o)
    assert t[3, 4] == 0 and t[5, 6] == 1 and t[7, 8] == -1
    # non-contiguous view
    wt = torch.randn(48, 80)
    t, _ = atq.adaptive_ternary_quantization(wt.to(DEV).t(), None, 0.05, 0.3)
    t_o, _, _ = O.adaptive_ternary_quantization(wt.t().contiguous().numpy(), None, 0.05, 0.3)
    assert np.array_equal(t.cpu().numpy().astype(np.int8), t_o)


def test_quantizer_golden(golden, policy):
    w = torch.from_numpy(golden["q_ties_w"]).to(DEV)
    for i, s in enumerate(policy["q_ties_s"]):
        t, a = atq.adaptive_ternary_quantization(w, None, 0.05, s)
        assert np.array_equal(t.cpu().numpy().astype(np.int8), golden[f"q_ties_t{i}"]), s
        assert abs(float(a) - float(golden[f"q_ties_a{i}"])) < 1e-6
    w = torch.from_numpy(golden["q_randn_w"]).to(DEV)
    for i, s in enumerate(policy["q_randn_s"]):
        t, a = atq.adaptive_ternary_quantization(w, None, 0.05, s)
        ref = _unbits(golden[f"q_randn_t{i}"], w.shape)
        diff = int((t.cpu().numpy().astype(np.int8) != ref).sum())
        assert diff <= (2 if s == 0.0 else 0)
        assert abs(float(a) - float(golden[f"q_randn_a{i}"])) < 1e-5
        assert abs(float((t == 0).float().mean()) - policy[f"q_randn_zero_frac{i}"]) < 1e-4
    for name in ("k1", "k2", "k3", "k4", "k5"):
        t, _ = atq.adaptive_ternary_quantization(torch.from_numpy(golden[f"q_{name}_w"]).to(DEV),
                                                 torch.tensor([1.25], device=DEV), 0.05, policy[f"q_{name}_s"])
        assert np.array_equal(t.cpu().numpy().astype(np.int8), golden[f"q_{name}_t"]), name


def test_threshold_batched_matches_single():
    shapes = [(192, 192), (384, 192), (1, 96), (96, 192), (192, 512), (10, 128), (3, 5)]
    ws = [_weights(s, i, "uniform").to(DEV) for i, s in enumerate(shapes)]
    ss = [0.3, 0.2, 0.1333, 0.1, 0.15, 0.0, 1.0]
    thr = eng.adaptive_threshold_batched(ws, ss)
    for i, (w, s) in enumerate(zip(ws, ss)):
        assert float(thr[i]) == float(eng.adaptive_threshold(w, s))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 12, 1001, 4096, 65537, 1 << 20])
def test_codec_bit_exact(n):
    g = torch.Generator().manual_seed(n)
    t = torch.randint(-1, 2, (n,), generator=g).float()
    p = TernaryBitPacking.pack_ternary_weights(t.to(DEV))
    assert p["packed_weights"].dtype == torch.uint8 and p["packed_weights"].numel() == (n + 3) // 4
    assert np.array_equal(p["packed_weights"].cpu().numpy(), O.pack2(t.numpy()))
    assert p["metadata"] == {"num_values": n, "encoding": {0: -1, 1: 0, 2: 1}}
    u = TernaryBitPacking.unpack_ternary_weights(p)
    assert u.dtype == torch.float32 and torch.equal(u.cpu(), t)
    for dt in (torch.bfloat16, torch.int8):
        assert torch.equal(eng.unpack2(p["packed_weights"], n, dt).float().cpu(), t)


def test_codec_golden_and_errors(golden, policy):
    doc = torch.from_numpy(golden["pack_doc_in"]).to(DEV)
    p = TernaryBitPacking.pack_ternary_weights(doc)
    assert p["packed_weights"].cpu().tolist() == [0x24, 0x49, 0x92]
    assert tuple(p["original_shape"]) == (3, 4) and p["metadata"]["num_values"] == 12
    assert torch.equal(TernaryBitPacking.unpack_ternary_weights(p).cpu(), torch.from_numpy(golden["pack_doc_unpacked"]))
    p = TernaryBitPacking.pack_ternary_weights(torch.from_numpy(golden["pack_tail_in"]).to(DEV))
    assert p["packed_weights"].cpu().tolist() == [0x86, 0x06]
    for name in ("r1", "r2", "r3"):
        t = torch.from_numpy(golden[f"pack_{name}_in"]).float().to(DEV)
        p = TernaryBitPacking.pack_ternary_weights(t)
        assert np.array_equal(p["packed_weights"].cpu().numpy(), golden[f"pack_{name}_bytes"])
        assert torch.equal(TernaryBitPacking.unpack_ternary_weights(p), t)
    with pytest.raises(ValueError, match="Input must contain only ternary values"):
        TernaryBitPacking.pack_ternary_weights(torch.tensor([0.5, 1.0], device=DEV))
    with pytest.raises(ValueError):
        TernaryBitPacking.pack_ternary_weights(torch.tensor([0.0, float("nan"), 1.0, 1.0, 0.0], device=DEV))
    neg0 = TernaryBitPacking.pack_ternary_weights(torch.tensor([-0.0, 0.0, 1.0, -1.0], device=DEV))
    assert neg0["packed_weights"].cpu().tolist() == [0x01 | (0x01 << 2) | (0x02 << 4) | (0x00 << 6)]
    bad = {"packed_weights": torch.tensor([0xFF], dtype=torch.uint8, device=DEV), "original_shape": torch.Size([4]),
           "metadata": {"num_values": 4, "encoding": {0: -1, 1: 0, 2: 1}}}
    with pytest.raises(KeyError):
        TernaryBitPacking.unpack_ternary_weights(bad)
    t = torch.from_numpy(golden["ftm_t"]).float().to(DEV)
    y = TernaryBitPacking.fast_ternary_matmul(TernaryBitPacking.pack_ternary_weights(t),
                                              torch.from_numpy(golden["ftm_x"]).to(DEV), alpha=2.0)
    assert torch.allclose(y.cpu(), torch.from_numpy(golden["ftm_y"]), rtol=1e-2, atol=1e-3)


def test_codec_large_roundtrip_properties():
    """BASELINE config 5 scale (2^28 here per tensor; the bench runs 1B in layer chunks):
    quantize -> pack -> unpack is the identity on T, zero fraction matches the target, and a byte
    checksum of the packed stream equals the checksum of packing the unpacked tensor again."""
    n = 1 << 28
    g = torch.Generator(device=DEV).manual_seed(5)
    w = (torch.rand(n, device=DEV, generator=g) * 2 - 1) / 64
    thr = eng.adaptive_threshold(w, 0.3)
    t = eng.ternarize_f32(w, thr)
    k = int(0.3 * n)
    below = int((w.abs() < thr).sum())
    at = int((w.abs() == thr).sum())
    assert below <= k < below + at  # thr is exactly the k-th order statistic
    packed = eng.ternarize_pack2(w, thr)
    u = eng.unpack2(packed, n)
    assert torch.equal(u, t)
    p2, flag = eng.pack2_from_f32(u)
    assert int(flag) == 0 and torch.equal(p2, packed)
    assert int(packed.long().sum()) == int(p2.long().sum())
    assert abs(float((t == 0).float().mean()) - 0.3) < 1e-3


def test_routing(golden, policy):
    x = torch.from_numpy(golden["route_x"]).to(DEV).requires_grad_(True)
    gy = torch.from_numpy(golden["route_gy"]).to(DEV)
    for i, case in enumerate(policy["route_cases"]):
        x.grad = None
        y = atq.SelectiveGradientRouting.apply(x, 0.05, case["f"])
        y.backward(gy)
        assert np.array_equal(x.grad.cpu().numpy(), golden[f"route_g{i}"])
        assert int((x.grad != 0).sum()) == case["kept"]
    with pytest.raises(RuntimeError, match="kthvalue"):
        atq.SelectiveGradientRouting.apply(x, 0.05, 1.0).backward(gy)
    assert atq.apply_selective_routing(x) is x
    # larger randomised case vs the oracle
    xr = torch.randn(333, 77)
    gr = torch.randn(333, 77)
    xg = xr.to(DEV).requires_grad_(True)
    atq.SelectiveGradientRouting.apply(xg, 0.05, 0.3).backward(gr.to(DEV))
    assert np.array_equal(xg.grad.cpu().numpy(), O.routing_backward(xr.numpy(), gr.numpy(), 0.3))


@pytest.mark.parametrize("kind", ["uniform", "normal", "ties", "wide", "constant", "two_level"])
def test_sampled_select_large_layers_bit_exact(kind):
    """n >= 32M takes the sampling front-end (bracket + one filter pass + candidate select); massive ties
    and adversarial layouts must fall back to the full radix select and stay exact."""
    n = (1 << 25) + 3
    g = torch.Generator().manual_seed(11)
    if kind == "constant":
        w = torch.full((n,), 0.25)
    elif kind == "two_level":  # sample positions see only small values, the rest is large
        w = torch.rand(n, generator=g) + 1.0
        stride = n // 8192
        w[stride // 2:: stride] = 1e-3
    else:
        w = _weights((n,), 17, kind)
    wg = w.to(DEV)
    a = np.abs(w.numpy())
    for s in (0.3, 0.05, 0.97, 1e-6):
        k = int(s * n)
        if k == 0:
            continue
        want = np.float32(np.partition(a, k)[k])
        got = eng.adaptive_threshold(wg, s)
        assert np.float32(got.item()).tobytes() == want.tobytes(), (kind, s)
    # routing percentile goes through the same kernel (k-1 indexing)
    got = eng.select_kth_abs(wg, 0)
    assert np.float32(got.item()).tobytes() == np.float32(a.min()).tobytes()
    got = eng.select_kth_abs(wg, n - 1)
    assert np.float32(got.item()).tobytes() == np.float32(a.max()).tobytes()
    # batched entry point mixes small, medium (sampled only because the batch is large) and large layers
    small = _weights((192, 192), 3, "uniform").to(DEV)
    medium = wg[: (1 << 22) + 8]
    thr = eng.adaptive_threshold_batched([small, wg, medium, small], [0.3, 0.3, 0.2, 0.1])
    assert float(thr[1]) == float(eng.adaptive_threshold(wg, 0.3))
    assert float(thr[0]) == float(eng.adaptive_threshold(small, 0.3))
    assert float(thr[3]) == float(eng.adaptive_threshold(small, 0.1))
    km = int(0.2 * medium.numel())
    assert np.float32(thr[2].item()).tobytes() == np.float32(np.partition(a[: medium.numel()], km)[km]).tobytes()


def test_batched_codec_matches_per_layer_and_oracle():
    """Whole-model codec kernels (one launch, 128-bit packed accesses): bytes identical to the per-layer kernels and
    to the oracle for ragged layer sizes (tails shorter than a 2048-weight warp step, n % 4 != 0)."""
    g = torch.Generator().manual_seed(11)
    shapes = [(4096, 512), (300, 77), (1, 96), (2050, 1), (64, 64), (5, 5), (1, 1), (1000, 2049)]
    ws = [((torch.rand(*s, generator=g) * 2 - 1) / 8).to(DEV) for s in shapes]
    ss = [0.05, 0.3, 0.2, 0.5, 0.13125, 0.4, 0.3, 0.25]
    thr = eng.adaptive_threshold_batched(ws, ss)
    packed = eng.ternarize_pack2_batched(ws, thr)
    for w, s, p, t in zip(ws, ss, packed, thr):
        want = eng.ternarize_pack2(w, t)
        assert torch.equal(p, want)
        t_ref, _, _ = O.adaptive_ternary_quantization(w.cpu().numpy(), None, 0.05, s)
        assert np.array_equal(p.cpu().numpy(), O.pack2(t_ref))
    outs, flag = eng.unpack2_batched(packed, [w.numel() for w in ws])
    assert int(flag) == 0
    for w, p, o in zip(ws, packed, outs):
        assert torch.equal(o, eng.unpack2(p, w.numel()))
    repacked, flag = eng.pack2_from_f32_batched(outs)
    assert int(flag) == 0 and all(torch.equal(a, b) for a, b in zip(repacked, packed))
    # validation flags: a non-ternary value / a code 3 anywhere in the batch
    bad = [o.clone() for o in outs]
    bad[3][7] = 0.5
    _, flag = eng.pack2_from_f32_batched(bad)
    assert int(flag) == 1
    badp = [p.clone() for p in packed]
    badp[0][123] = 0xFF
    _, flag = eng.unpack2_batched(badp, [w.numel() for w in ws])
    assert int(flag) == 1


@pytest.mark.parametrize("fused", [1, 0])
def test_select_fused_and_per_pass_paths_agree_with_sort(fused):
    """The exact select runs its three digit passes in one cooperative launch (default) or as one launch per pass
    (atq_set_fused_select(0)); both must return sorted(|x|)[k] bit for bit -- single layers (ragged sizes, unaligned
    start, ties, constants, first and last rank) and a batch of layers of mixed sizes, twice in a row (state re-armed)."""
    import atq._native as nv
    nv.lib.atq_set_fused_select(fused)
    try:
        g = torch.Generator().manual_seed(17)
        cases = [torch.randn(n, generator=g) for n in (1, 2, 5, 1001, 65537, (1 << 20) + 3, 5_000_011)]
        cases.append(torch.randint(-3, 4, (300_001,), generator=g).float())     # heavy ties
        cases.append(torch.full((70_000,), -0.375))                              # constant
        big = torch.randn((1 << 20) + 9, generator=g)
        cases.append(big[1:])                                                    # 4-byte aligned, not 16
        for x in cases:
            xg = x.to(DEV)
            if x.data_ptr() != x.untyped_storage().data_ptr():                  # keep the misalignment on the device
                xg = big.to(DEV)[1:]
            srt = torch.sort(xg.abs()).values
            n = xg.numel()
            for k in sorted({0, n // 3, n // 2, n - 1}):
                for _ in range(2):
                    got = eng.select_kth_abs(xg, k)
                    assert got.item() == srt[k].item(), (fused, n, k)
        ws = [torch.randn(s, generator=g).to(DEV) for s in ((192, 192), (1000, 1003), (7,), (2048, 1024), (33, 65))]
        ss = [0.3, 0.05, 0.5, 0.1333, 0.999]
        for _ in range(2):
            thr = eng.adaptive_threshold_batched(ws, ss)
            for t, w, s in zip(thr, ws, ss):
                assert t.item() == np.float32(O.adaptive_threshold(w.cpu().numpy(), s)), (fused, tuple(w.shape), s)
    finally:
        nv.lib.atq_set_fused_select(1)
