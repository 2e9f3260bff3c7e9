"""Small end-to-end pass over every kernel family (compute-sanitizer target: tools/sanitize.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch

import atq
import atq._engine as eng
from atq import attention as A
from atq.bit_packing import TernaryBitPacking
from atq.contrastive import HardNegativeMiningInfoNCE

dev = torch.device("cuda:0")
torch.manual_seed(0)
# quantizer / codec (per-layer and whole-model forms)
ws = [((torch.rand(m, k) * 2 - 1) / 8).to(dev) for m, k in ((384, 512), (77, 33), (2050, 3))]
thr = eng.adaptive_threshold_batched(ws, [0.3, 0.1, 0.2])
packed = eng.ternarize_pack2_batched(ws, thr)
outs, flag = eng.unpack2_batched(packed, [w.numel() for w in ws])
eng.pack2_from_f32_batched(outs)
t, a = atq.adaptive_ternary_quantization(ws[0], None, 0.05, 0.3)
TernaryBitPacking.unpack_ternary_weights(TernaryBitPacking.pack_ternary_weights(t))
# layers: RPB on CTA pairs (rows >= 256, cols >= 128) and single-CTA shapes, TernaryLinear incl. the packed-B kernel
for mode in ("parity", "parity_bf16", "fast"):
    atq.set_gemm_mode(mode)
    for mod, n_tok in ((atq.ResidualPrecisionBoostLinear(384, 512, 0.2, True, 0.15), 300), (atq.ResidualPrecisionBoostLinear(96, 40, 0.2, True, 0.1), 50),
                       (atq.TernaryLinear(256, 192), 100), (atq.TernaryLinear(256, 192), 700)):
        mod = mod.to(dev)
        x = torch.randn(n_tok, mod.in_features, device=dev, requires_grad=True)
        mod(x).square().mean().backward()
atq.set_gemm_mode("parity")
# block-level kernels: attention (head dims 24 and 64), fused FFN, gated residual, fused loss, cluster split
for heads, hd, l in ((8, 24, 50), (2, 64, 197)):
    q, k, v = (torch.randn(2, l, heads * hd, device=dev, requires_grad=True) for _ in range(3))
    pad = torch.zeros(2, l, dtype=torch.bool, device=dev)
    pad[1, l // 2:] = True
    A.attention_core(q, k, v, heads, pad, None, 0.1, True).sum().backward()
l1 = atq.ResidualPrecisionBoostLinear(128, 256, 0.2, True, 0.1).to(dev)
l2 = atq.ResidualPrecisionBoostLinear(256, 128, 0.4, True, 0.1).to(dev)
h = torch.randn(300, 128, device=dev, requires_grad=True)
y = atq.fused_ffn(l1, l2, h, 0.1, True)
g = torch.sigmoid(torch.ones(1, device=dev, requires_grad=True) * 0.8)
atq.gated_residual(h, y, g, 0.1, True).sum().backward()
img = torch.randn(96, 64, device=dev, requires_grad=True)
txt = torch.randn(96, 64, device=dev, requires_grad=True)
HardNegativeMiningInfoNCE()(img, txt).backward()
torch.cuda.synchronize()
print("sanitize target ok")
