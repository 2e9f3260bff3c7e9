"""CPU: the harness contrastive loss (workloads.models.HardNegativeInfoNCE + ContrastiveManager, the vectorised
restatement of utils/enhanced_contrastive.py:8-162, 269-417) against a golden fixture produced by RUNNING the
reference's own loss (tests/golden/gen_loss_golden.py): loss values, both gradients, temperature schedule."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from workloads import models as M

G = np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))


@pytest.mark.parametrize("tag", ["b16", "b64", "b96_late"])
def test_contrastive_loss_matches_reference_fixture(tag):
    img = torch.from_numpy(G[f"{tag}.img"]).requires_grad_(True)
    txt = torch.from_numpy(G[f"{tag}.txt"]).requires_grad_(True)
    epoch, total = (int(v) for v in G[f"{tag}.cfg"])
    crit = M.HardNegativeInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, temperature_schedule=True)
    man = M.ContrastiveManager(crit)
    crit.set_epoch(epoch, total)
    man.set_epoch(epoch, total)
    assert abs(crit.get_current_temperature() - float(G[f"{tag}.temperature"])) < 1e-12
    loss = man.compute_loss(img, txt)
    loss.backward()
    assert torch.allclose(loss.detach(), torch.from_numpy(G[f"{tag}.loss"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(img.grad, torch.from_numpy(G[f"{tag}.dimg"]), rtol=1e-5, atol=1e-7)
    assert torch.allclose(txt.grad, torch.from_numpy(G[f"{tag}.dtxt"]), rtol=1e-5, atol=1e-7)
