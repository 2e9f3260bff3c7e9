"""Golden fixture for the contrastive loss: RUN THE REFERENCE's own
utils.enhanced_contrastive.HardNegativeMiningInfoNCE / ContrastiveLearningManager (CPU) on seeded embeddings and
store inputs, loss values and gradients.

    python tests/golden/gen_loss_golden.py      (build container only: needs /root/reference)

Writes tests/golden/loss_golden.npz.  Nothing in it is computed by this repo's code.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ATQ_REFERENCE", "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot"):  # utils/__init__ imports the plotting module
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
from utils.enhanced_contrastive import ContrastiveLearningManager, HardNegativeMiningInfoNCE  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for tag, (b, e, epoch, total) in {"b16": (16, 192, 5, 10), "b64": (64, 96, 0, 10), "b96_late": (96, 64, 9, 10)}.items():
    g = torch.Generator().manual_seed(b * 7 + e)
    img = torch.randn(b, e, generator=g).requires_grad_(True)
    txt = (0.5 * img.detach() + torch.randn(b, e, generator=g)).requires_grad_(True)  # correlated pairs
    crit = HardNegativeMiningInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, temperature_schedule=True)
    man = ContrastiveLearningManager(None, crit)  # (model, criterion): the model is only used for mining, not for the loss
    crit.set_epoch(epoch, total)
    man.set_epoch(epoch, total)
    loss = man.compute_loss(img, txt)
    loss.backward()
    out[f"{tag}.img"], out[f"{tag}.txt"] = img.detach().numpy(), txt.detach().numpy()
    out[f"{tag}.cfg"] = np.array([epoch, total], dtype=np.int64)
    out[f"{tag}.loss"] = loss.detach().numpy()
    out[f"{tag}.dimg"], out[f"{tag}.dtxt"] = img.grad.numpy(), txt.grad.numpy()
    out[f"{tag}.temperature"] = np.array(crit.get_current_temperature())
np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)
print("wrote", len(out), "arrays", {k: float(v) for k, v in out.items() if k.endswith(".loss")})
