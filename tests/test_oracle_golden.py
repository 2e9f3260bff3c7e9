"""CPU: the oracle (oracle/atq_oracle.py) against the golden fixtures produced by the reference,
and against the live reference when /root/reference is present (build container)."""
import numpy as np
import pytest
import torch

from conftest import have_reference, load_reference_atq
from oracle import atq_oracle as O


def _unbits(bits, shape):
    plus = np.unpackbits(bits[0])[: int(np.prod(shape))].astype(np.int8)
    minus = np.unpackbits(bits[1])[: int(np.prod(shape))].astype(np.int8)
    return (plus - minus).reshape(shape)


def test_pack_golden(golden, policy):
    assert O.pack2(golden["pack_doc_in"]).tolist() == golden["pack_doc_bytes"].tolist() == [0x24, 0x49, 0x92]
    assert O.pack2(golden["pack_tail_in"]).tolist() == golden["pack_tail_bytes"].tolist() == [0x86, 0x06]
    for name in ("r1", "r2", "r3"):
        t = golden[f"pack_{name}_in"]
        b = O.pack2(t)
        assert np.array_equal(b, golden[f"pack_{name}_bytes"])
        assert np.array_equal(O.unpack2(b, t.size).reshape(t.shape), golden[f"pack_{name}_unpacked"])
    with pytest.raises(ValueError, match="ternary values"):
        O.pack2(np.array([0.5, 1.0]))
    assert policy["pack_invalid_raises"] == "Input must contain only ternary values (-1, 0, 1)"
    with pytest.raises(KeyError):
        O.unpack2(np.array([0xFF], np.uint8), 4)
    assert O.compute_memory_savings(4096 * 4096) == policy["memory_savings_4096"]
    assert O.compute_memory_savings(13) == policy["memory_savings_13"]


def test_fast_matmul_golden(golden):
    t = golden["ftm_t"]
    y = O.fast_ternary_matmul(O.pack2(t), t.shape, golden["ftm_x"], 2.0)
    assert np.array_equal(y, golden["ftm_y"])


def test_quantizer_golden(golden, policy):
    w = golden["q_ties_w"]
    for i, s in enumerate(policy["q_ties_s"]):
        t, a, _ = O.adaptive_ternary_quantization(w, None, 0.05, s)
        assert np.array_equal(t, golden[f"q_ties_t{i}"]), s
        assert abs(float(a) - float(golden[f"q_ties_a{i}"])) < 1e-6
    w = golden["q_randn_w"]
    for i, s in enumerate(policy["q_randn_s"]):
        t, a, _ = O.adaptive_ternary_quantization(w, None, 0.05, s)
        ref = _unbits(golden[f"q_randn_t{i}"], w.shape)
        if s == 0.0:
            # mean branch: the threshold is an fp32 reduction, elements within 1 ulp of it may differ
            assert (t != ref).sum() <= 2
        else:
            assert np.array_equal(t, ref), s
        assert abs(float(a) - float(golden[f"q_randn_a{i}"])) < 1e-5 * max(1.0, abs(float(a)))
    for name in ("k1", "k2", "k3", "k4", "k5"):
        t, _, _ = O.adaptive_ternary_quantization(golden[f"q_{name}_w"], 1.25, 0.05, policy[f"q_{name}_s"])
        assert np.array_equal(t, golden[f"q_{name}_t"]), name


def test_layers_golden(golden, policy):
    tl = O.OracleTernaryLinear(64, 32)
    with torch.no_grad():
        tl.weight.copy_(torch.from_numpy(golden["tl_weight"]))
        tl.bias.copy_(torch.from_numpy(golden["tl_bias"]))
        tl.alpha.copy_(torch.from_numpy(golden["tl_alpha"]))
    x = torch.from_numpy(golden["tl_x"]).requires_grad_(True)
    y = tl(x)
    y.backward(torch.from_numpy(golden["tl_gy"]))
    assert tl.weight.grad is None and policy["tl_weight_grad_is_none"]
    assert torch.equal(y.detach(), torch.from_numpy(golden["tl_y"]))
    assert torch.equal(x.grad, torch.from_numpy(golden["tl_dx"]))
    assert torch.equal(tl.alpha.grad, torch.from_numpy(golden["tl_dalpha"]))
    assert torch.equal(tl.bias.grad, torch.from_numpy(golden["tl_dbias"]))

    rpb = O.OracleRPBLinear(64, 32, 0.05, True, 0.3)
    with torch.no_grad():
        rpb.weight.copy_(torch.from_numpy(golden["rpb_weight"]))
        rpb.bias.copy_(torch.from_numpy(golden["rpb_bias"]))
        rpb.alpha.copy_(torch.from_numpy(golden["rpb_alpha"]))
        rpb.precision_mask.copy_(torch.from_numpy(golden["rpb_mask"]))
    x = torch.from_numpy(golden["rpb_x"]).requires_grad_(True)
    y = rpb(x)
    y.backward(torch.from_numpy(golden["rpb_gy"]))
    assert torch.equal(y.detach(), torch.from_numpy(golden["rpb_y"]))
    assert torch.equal(x.grad, torch.from_numpy(golden["rpb_dx"]))
    assert torch.equal(rpb.weight.grad, torch.from_numpy(golden["rpb_dw"]))
    assert torch.equal(rpb.alpha.grad, torch.from_numpy(golden["rpb_dalpha"]))
    assert int((rpb.weight.grad != 0).sum()) == policy["rpb_dw_nonzeros"] == 102
    t, _ = rpb.get_quantized_weights()
    assert np.array_equal(t.numpy().astype(np.int8), golden["rpb_tq"])

    # closed form (fp64) agrees with the autograd graph
    cf = O.ternary_linear_reference(torch.from_numpy(golden["rpb_x"]), rpb.weight.detach(), rpb.alpha.detach(),
                                    rpb.bias.detach(), 0.3, rpb.precision_mask, torch.from_numpy(golden["rpb_gy"]))
    assert torch.allclose(cf["y"].float(), y.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(cf["dw"].float(), rpb.weight.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(cf["dalpha"].float(), rpb.alpha.grad, rtol=1e-4, atol=1e-4)


def test_mask_popcounts(policy):
    for key, want in policy["mask_popcounts"].items():
        shape, r = key.split("@")
        m, k = (int(v) for v in shape.split("x"))
        torch.manual_seed(0)
        w = torch.empty(m, k)
        torch.nn.init.kaiming_uniform_(w, a=5 ** 0.5)
        assert int(O.precision_mask_from_weight(w, float(r)).sum()) == want == int(float(r) * m * k)


def test_routing_golden(golden, policy):
    x, gy = golden["route_x"], golden["route_gy"]
    for i, case in enumerate(policy["route_cases"]):
        g = O.routing_backward(x, gy, case["f"])
        assert np.array_equal(g, golden[f"route_g{i}"])
        assert int((g != 0).sum()) == case["kept"]
    with pytest.raises(RuntimeError, match="kthvalue"):
        O.routing_backward(x, gy, 1.0)
    assert policy["route_f1_raises"].startswith("kthvalue()")


@pytest.mark.skipif(not have_reference(), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("shape,s", [((257, 33), 0.3), ((64, 64), 0.05), ((1, 96), 0.2), ((1000,), 0.5),
                                     ((31, 7), 0.999), ((12, 12), 0.0), ((5, 5), 1.0)])
def test_oracle_vs_live_reference(shape, s):
    ref = load_reference_atq()
    torch.manual_seed(hash((shape, s)) % 1000)
    w = torch.randn(*shape) * 0.1
    w.view(-1)[::7] = w.view(-1)[0]  # force ties
    t_ref, a_ref = ref.quantizers.adaptive_ternary_quantization(w, None, 0.05, s)
    t, a, _ = O.adaptive_ternary_quantization(w.numpy(), None, 0.05, s)
    if s == 0.0:
        assert (t != t_ref.numpy()).sum() <= 1
    else:
        assert np.array_equal(t, t_ref.numpy().astype(np.int8))
    assert abs(float(a) - float(a_ref)) <= 1e-5 * max(1.0, abs(float(a_ref)))
    p_ref = ref.bit_packing.TernaryBitPacking.pack_ternary_weights(t_ref)
    assert np.array_equal(O.pack2(t), p_ref["packed_weights"].numpy())
