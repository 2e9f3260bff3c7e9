"""Generate the golden fixtures by RUNNING THE REFERENCE (/root/reference) on seeded inputs.

Run in the build container (the reference does not travel to the GPU box):
    python tests/golden/gen_golden.py
Writes tests/golden/atq_golden.npz and tests/golden/policy_golden.json.
Every array is an input to, or an output of, the reference's own code; nothing here is
computed by this repo's oracle or CUDA path.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ATQ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from atq.quantizers import adaptive_ternary_quantization  # noqa: E402
from atq.layers import TernaryLinear  # noqa: E402
from atq.precision_boost import ResidualPrecisionBoostLinear  # noqa: E402
from atq.routing import SelectiveGradientRouting, apply_selective_routing  # noqa: E402
from atq.bit_packing import TernaryBitPacking  # noqa: E402
from atq.mixed_precision_atq import (MixedPrecisionATQ, GradualQuantizationScheduler,  # noqa: E402
                                     PrecisionControlledLinear)

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
policy = {}

# ---------------- codec (atq/bit_packing.py) ----------------
doc = torch.tensor([[-1, 0, 1, -1], [0, 1, -1, 0], [1, -1, 0, 1]], dtype=torch.float)
p = TernaryBitPacking.pack_ternary_weights(doc)
out["pack_doc_in"] = doc.numpy()
out["pack_doc_bytes"] = p["packed_weights"].numpy()
policy["pack_doc_meta"] = {"num_values": p["metadata"]["num_values"],
                           "encoding": {str(k): v for k, v in p["metadata"]["encoding"].items()},
                           "original_shape": list(p["original_shape"])}
out["pack_doc_unpacked"] = TernaryBitPacking.unpack_ternary_weights(p).numpy()

tail = torch.tensor([1, 0, -1, 1, 1, 0], dtype=torch.float)
p = TernaryBitPacking.pack_ternary_weights(tail)
out["pack_tail_in"] = tail.numpy()
out["pack_tail_bytes"] = p["packed_weights"].numpy()

g = torch.Generator().manual_seed(1234)
for name, shape in (("r1", (37, 29)), ("r2", (5, 3, 7)), ("r3", (1001,))):
    t = torch.randint(-1, 2, shape, generator=g).float()
    p = TernaryBitPacking.pack_ternary_weights(t)
    out[f"pack_{name}_in"] = t.numpy().astype(np.int8)
    out[f"pack_{name}_bytes"] = p["packed_weights"].numpy()
    out[f"pack_{name}_unpacked"] = TernaryBitPacking.unpack_ternary_weights(p).numpy().astype(np.int8)

try:
    TernaryBitPacking.pack_ternary_weights(torch.tensor([0.5, 1.0]))
    policy["pack_invalid_raises"] = None
except ValueError as e:
    policy["pack_invalid_raises"] = str(e)

policy["memory_savings_4096"] = TernaryBitPacking.compute_memory_savings(torch.zeros(4096, 4096))
policy["memory_savings_13"] = TernaryBitPacking.compute_memory_savings(torch.zeros(13))

t = torch.randint(-1, 2, (24, 40), generator=g).float()
x = torch.randn(9, 40, generator=g)
p = TernaryBitPacking.pack_ternary_weights(t)
out["ftm_t"] = t.numpy().astype(np.int8)
out["ftm_x"] = x.numpy()
out["ftm_y"] = TernaryBitPacking.fast_ternary_matmul(p, x, alpha=2.0).numpy()

# ---------------- quantizer (atq/quantizers.py) ----------------
ties = torch.tensor([0.1, -0.1, 0.2, -0.2, 0.3, -0.3, 0.4, -0.4, 0.5, -0.5])
out["q_ties_w"] = ties.numpy()
tie_s = [0.0, 0.1, 0.2, 0.3, 0.5, 0.95, 1.0]
policy["q_ties_s"] = tie_s
for i, s in enumerate(tie_s):
    tq, a = adaptive_ternary_quantization(ties, None, 0.05, s)
    out[f"q_ties_t{i}"] = tq.numpy().astype(np.int8)
    out[f"q_ties_a{i}"] = np.float32(a.item())

torch.manual_seed(0)
w = torch.randn(1000, 100)
out["q_randn_w"] = w.numpy()
rs = [0.0, 0.3, 0.999999]
policy["q_randn_s"] = rs
for i, s in enumerate(rs):
    tq, a = adaptive_ternary_quantization(w, None, 0.05, s)
    out[f"q_randn_t{i}"] = np.packbits((tq.numpy().astype(np.int8) + 1).astype(np.uint8) == 2), \
        np.packbits((tq.numpy().astype(np.int8) + 1).astype(np.uint8) == 0)
    out[f"q_randn_t{i}"] = np.stack(out[f"q_randn_t{i}"])
    out[f"q_randn_a{i}"] = np.float32(a.item())
    policy[f"q_randn_zero_frac{i}"] = float((tq == 0).float().mean())

# kaiming-uniform shaped layers with odd sizes, alpha passed through
torch.manual_seed(7)
for name, (m, k, s) in {"k1": (33, 65, 0.3), "k2": (10, 128, 0.05), "k3": (1, 96, 0.1333),
                        "k4": (96, 192, 0.2), "k5": (128, 131, 0.15)}.items():
    lin = torch.nn.Linear(k, m)
    w = lin.weight.detach().clone()
    alpha_in = torch.nn.Parameter(torch.tensor([1.25]))
    tq, a = adaptive_ternary_quantization(w, alpha_in, 0.05, s)
    assert a is alpha_in
    out[f"q_{name}_w"] = w.numpy()
    out[f"q_{name}_t"] = tq.numpy().astype(np.int8)
    policy[f"q_{name}_s"] = s

# ---------------- layers ----------------
torch.manual_seed(11)
lin = TernaryLinear(64, 32)
with torch.no_grad():
    lin.alpha.fill_(0.75)
x = torch.randn(5, 7, 64, requires_grad=True)
y = lin(x)
gy = torch.randn_like(y)
y.backward(gy)
assert lin.weight.grad is None
out["tl_weight"] = lin.weight.detach().numpy()
out["tl_bias"] = lin.bias.detach().numpy()
out["tl_alpha"] = lin.alpha.detach().numpy()
out["tl_x"] = x.detach().numpy()
out["tl_gy"] = gy.numpy()
out["tl_y"] = y.detach().numpy()
out["tl_dx"] = x.grad.numpy()
out["tl_dalpha"] = lin.alpha.grad.numpy()
out["tl_dbias"] = lin.bias.grad.numpy()
policy["tl_weight_grad_is_none"] = True
policy["tl_zero_frac"] = float((adaptive_ternary_quantization(lin.weight, lin.alpha)[0] == 0).float().mean())

torch.manual_seed(12)
rpb = ResidualPrecisionBoostLinear(64, 32, precision_ratio=0.05, sparsity_target=0.3)
with torch.no_grad():
    rpb.alpha.fill_(0.6)
x = torch.randn(11, 64, requires_grad=True)
y = rpb(x)
gy = torch.randn_like(y)
y.backward(gy)
out["rpb_weight"] = rpb.weight.detach().numpy()
out["rpb_bias"] = rpb.bias.detach().numpy()
out["rpb_alpha"] = rpb.alpha.detach().numpy()
out["rpb_mask"] = rpb.precision_mask.numpy()
out["rpb_x"] = x.detach().numpy()
out["rpb_gy"] = gy.numpy()
out["rpb_y"] = y.detach().numpy()
out["rpb_dx"] = x.grad.numpy()
out["rpb_dw"] = rpb.weight.grad.numpy()
out["rpb_dalpha"] = rpb.alpha.grad.numpy()
out["rpb_dbias"] = rpb.bias.grad.numpy()
policy["rpb_dw_nonzeros"] = int((rpb.weight.grad != 0).sum())
policy["rpb_mask_popcount"] = int(rpb.precision_mask.sum())
tq, a = rpb.get_quantized_weights()
assert a is rpb.alpha
out["rpb_tq"] = tq.numpy().astype(np.int8)

# second RPB case: other ratio/sparsity, no bias, s changed after construction
torch.manual_seed(13)
rpb = ResidualPrecisionBoostLinear(96, 40, precision_ratio=0.4, bias=False, sparsity_target=0.1)
rpb.sparsity_target = 0.1333
x = torch.randn(3, 4, 96, requires_grad=True)
y = rpb(x)
gy = torch.randn_like(y)
y.backward(gy)
for k_, v_ in dict(weight=rpb.weight.detach(), alpha=rpb.alpha.detach(), mask=rpb.precision_mask, x=x.detach(),
                   gy=gy, y=y.detach(), dx=x.grad, dw=rpb.weight.grad, dalpha=rpb.alpha.grad).items():
    out[f"rpb2_{k_}"] = v_.numpy()

# mask popcounts the survey lists
pc = {}
for (m, k, r) in ((192, 192, 0.2), (192, 192, 0.4), (192, 512, 0.2), (1, 96, 0.2), (128, 3136, 0.05), (10, 128, 0.1)):
    l_ = ResidualPrecisionBoostLinear(k, m, precision_ratio=r)
    pc[f"{m}x{k}@{r}"] = int(l_.precision_mask.sum())
policy["mask_popcounts"] = pc

# ---------------- routing ----------------
torch.manual_seed(21)
x = torch.randn(4, 5, requires_grad=True)
gy = torch.randn(4, 5)
policy["route_cases"] = []
for i, f in enumerate((0.3, 0.0, 0.7, 0.05)):
    x.grad = None
    SelectiveGradientRouting.apply(x, 0.05, f).backward(gy)
    out[f"route_g{i}"] = x.grad.numpy().copy()
    policy["route_cases"].append({"f": f, "kept": int((x.grad != 0).sum())})
out["route_x"] = x.detach().numpy()
out["route_gy"] = gy.numpy()
try:
    x.grad = None
    SelectiveGradientRouting.apply(x, 0.05, 1.0).backward(gy)
    policy["route_f1_raises"] = None
except RuntimeError as e:
    policy["route_f1_raises"] = str(e)[:80]
assert apply_selective_routing(x) is x

# ---------------- host policy (atq/mixed_precision_atq.py) ----------------
names = ["image_encoder.projector", "text_encoder.layers.0.self_attn.q_proj", "text_encoder.layers.1.linear1",
         "text_encoder.attention_pool.0", "text_projector", "fusion.cross_attention.q", "encoder.ffn.intermediate",
         "conv_stem", "image_projector", "final_head", "embed_tokens", "plain"]
tab = []
for nm in names:
    for ep in (0, 1, 5, 8, 9, 12):
        for thr in (0.3, 0.2, 0.05):
            pr, cs = MixedPrecisionATQ.calculate_quantization_params(None, nm, ep, 10, thr)
            tab.append([nm, ep, thr, pr, cs, MixedPrecisionATQ.get_layer_importance(None, nm)])
policy["quant_params"] = tab


class _Dummy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.image_proj = ResidualPrecisionBoostLinear(16, 8, precision_ratio=0.2, sparsity_target=0.1)
        self.text_attention = ResidualPrecisionBoostLinear(16, 8, precision_ratio=0.2, sparsity_target=0.1)
        self.ffn = ResidualPrecisionBoostLinear(16, 8)
        self.plain = TernaryLinear(16, 8)


sched_tab = []
for (E, wu, fe) in ((10, 2, None), (25, 5, None), (6, 1, 3)):
    d = _Dummy()
    sch = GradualQuantizationScheduler(d, E, 0.3, 0.2, warmup_epochs=wu, final_epochs=fe)
    rows = []
    for ep in range(E + 2):
        v, t_ = sch.step(ep)
        rows.append([ep, v, t_] + [[m.precision_ratio, m.sparsity_target] for m in (d.image_proj, d.text_attention, d.ffn)])
    sched_tab.append({"E": E, "warmup": wu, "final": fe, "vision": sch.vision_sparsity_schedule,
                      "text": sch.text_sparsity_schedule, "rows": rows,
                      "plain_has_sparsity": hasattr(d.plain, "sparsity_target")})
policy["scheduler"] = sched_tab

pcl = PrecisionControlledLinear(32, 16, importance=1.44)
policy["pcl"] = {"ratio": pcl.linear.precision_ratio, "sparsity": pcl.linear.sparsity_target,
                 "keys": sorted(pcl.state_dict().keys())}
policy["tl_state_keys"] = sorted(TernaryLinear(4, 4).state_dict().keys())
policy["rpb_state_keys"] = sorted(ResidualPrecisionBoostLinear(4, 4).state_dict().keys())

np.savez_compressed(os.path.join(HERE, "atq_golden.npz"), **out)
with open(os.path.join(HERE, "policy_golden.json"), "w") as f:
    json.dump(policy, f, indent=1, sort_keys=True)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "atq_golden.npz")), "bytes")
