"""One forward+backward of the fused attention core at the config-4 image-tower shape (ncu target)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch

import atq
from atq import attention as A

atq.set_gemm_mode(sys.argv[1] if len(sys.argv) > 1 else "parity")
b = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
h, l = 12, 197
q, k, v = (torch.randn(b, l, h * 64, device=dev, requires_grad=True) for _ in range(3))
seed = torch.tensor([5], dtype=torch.int64, device=dev)
for _ in range(2):
    q.grad = k.grad = v.grad = None
    A.attention_core(q, k, v, h, None, None, 0.1, True, seed=seed).backward(torch.ones(b, l, h * 64, device=dev))
torch.cuda.synchronize()
print("ok", float(q.grad.abs().sum()))
