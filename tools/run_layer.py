"""A few fwd+bwd iterations of ONE layer (ncu target):  python tools/run_layer.py OUTxIN tokens [rpb0.2|ternary] [mode] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch

import atq

out_f, in_f = (int(a) for a in sys.argv[1].split("x"))
tokens = int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "rpb0.2"
atq.set_gemm_mode(sys.argv[4] if len(sys.argv) > 4 else "parity")
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
mod = (atq.TernaryLinear(in_f, out_f) if kind == "ternary" else atq.ResidualPrecisionBoostLinear(in_f, out_f, float(kind[3:]), True, 0.3)).to(dev)
x = torch.randn(tokens, in_f, device=dev, requires_grad=True)
gy = torch.randn(tokens, out_f, device=dev)
for _ in range(iters):
    mod.zero_grad(set_to_none=True)
    x.grad = None
    mod(x).backward(gy)
torch.cuda.synchronize()
print("ok", float(x.grad.abs().sum()))
