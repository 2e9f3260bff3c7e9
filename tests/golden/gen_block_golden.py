"""Golden fixture for the transformer block: RUN THE REFERENCE's own
`models.text_encoder.TernaryTransformerLayer` (CPU, eval mode and a training-mode backward with dropout 0) on
seeded inputs and store its parameters, inputs, outputs and gradients.

    python tests/golden/gen_block_golden.py      (build container only: needs /root/reference)

Writes tests/golden/block_golden.npz.  Nothing in it is computed by this repo's oracle or CUDA path.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ATQ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from models.text_encoder import TernaryTransformerLayer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(1234)
E, HEADS, FF, B, L = 64, 1, 128, 3, 20
layer = TernaryTransformerLayer(E, HEADS, dim_feedforward=FF, dropout=0.0, use_rpb=True, sparsity_target=0.2)
with torch.no_grad():  # make alpha / gate / norms non-trivial
    for n, p in layer.named_parameters():
        if n.endswith("alpha"):
            p.fill_(0.6 + 0.1 * (len(n) % 5))
    layer.gate.fill_(0.3)
    layer.norm1.weight.uniform_(0.5, 1.5)
    layer.norm2.bias.uniform_(-0.2, 0.2)
x = torch.randn(B, L, E)
lengths = torch.tensor([20, 13, 7])
pad = torch.arange(L)[None, :] >= lengths[:, None]
gy = torch.randn(B, L, E)

out = {f"param.{k}": v.detach().numpy().copy() for k, v in layer.state_dict().items()}
out["x"], out["pad"], out["gy"] = x.numpy(), pad.numpy(), gy.numpy()
layer.eval()
with torch.no_grad():
    out["y_eval"] = layer(x, src_key_padding_mask=pad).numpy()
layer.train()  # dropout p = 0: deterministic
xr = x.clone().requires_grad_(True)
y = layer(xr, src_key_padding_mask=pad)
y.backward(gy)
out["y_train"] = y.detach().numpy()
out["dx"] = xr.grad.numpy()
for n, p in layer.named_parameters():
    if p.grad is not None:
        out[f"grad.{n}"] = p.grad.numpy().copy()
np.savez_compressed(os.path.join(HERE, "block_golden.npz"), **out)
print("wrote", len(out), "arrays;", sorted(k for k in out if k.startswith("grad."))[:6], "...")
