"""GPU parity of the fused hard-negative InfoNCE (csrc/loss_sm100.cu + atq/contrastive.py):
  * the golden fixture produced by running the reference's own loss on the CPU (tests/golden/gen_loss_golden.py),
  * the reference's own class (staged copy, utils/enhanced_contrastive.py) run on the same device in fp32 and fp64, up to
    the B = 4096 global batch of BASELINE config 4 (its mask loop :118-120 does 8 192 indexed writes there).
Criterion for gradients: |ours - f64| <= 1e-2 |f64| + 1e-3 max|f64| + 2 max|f32 - f64|."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, have_staged_reference

pytestmark = pytest.mark.gpu

import atq
from atq.contrastive import ContrastiveLearningManager, HardNegativeMiningInfoNCE

DEV = "cuda:0"
G = np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))


@pytest.mark.parametrize("tag", ["b16", "b64", "b96_late"])
def test_fused_loss_matches_reference_fixture(tag):
    img = torch.from_numpy(G[f"{tag}.img"]).to(DEV).requires_grad_(True)
    txt = torch.from_numpy(G[f"{tag}.txt"]).to(DEV).requires_grad_(True)
    epoch, total = (int(v) for v in G[f"{tag}.cfg"])
    crit = HardNegativeMiningInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, temperature_schedule=True)
    man = ContrastiveLearningManager(None, crit)
    crit.set_epoch(epoch, total)
    man.set_epoch(epoch, total)
    assert abs(crit.get_current_temperature() - float(G[f"{tag}.temperature"])) < 1e-12
    loss = man.compute_loss(img, txt)
    loss.backward()
    assert abs(float(loss) - float(G[f"{tag}.loss"])) <= 1e-5 * abs(float(G[f"{tag}.loss"])) + 1e-6
    for got, key in ((img.grad, "dimg"), (txt.grad, "dtxt")):
        want = torch.from_numpy(G[f"{tag}.{key}"])
        assert torch.allclose(got.cpu(), want, rtol=1e-2, atol=1e-3 * float(want.abs().max())), (tag, key, (got.cpu() - want).abs().max())


def _reference_loss_cls():
    from oracle import ref_env
    ref_env.activate("b200")  # `utils` = the reference's package, `atq` = this repo's (the loss module does not use atq)
    from utils.enhanced_contrastive import ContrastiveLearningManager as RefMan, HardNegativeMiningInfoNCE as RefLoss
    return RefLoss, RefMan


@pytest.mark.skipif(not have_staged_reference(), reason="oracle/_ref not staged (python oracle/install_ref.py)")
@pytest.mark.parametrize("b,e,epoch", [(16, 192, 5), (257, 96, 0), (512, 768, 9), (4096, 768, 5)])
def test_fused_loss_vs_reference_class_on_device(b, e, epoch):
    RefLoss, RefMan = _reference_loss_cls()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=DEV).manual_seed(b + e)
        img0 = torch.randn(b, e, device=DEV, generator=g)
        txt0 = 0.5 * img0 + torch.randn(b, e, device=DEV, generator=g)
        res = {}
        for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            img, txt = img0.detach().clone().to(dt).requires_grad_(True), txt0.detach().clone().to(dt).requires_grad_(True)
            crit = RefLoss(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, temperature_schedule=True)
            man = RefMan(None, crit)
            crit.set_epoch(epoch, 10)
            man.set_epoch(epoch, 10)
            loss = man.compute_loss(img, txt)
            loss.backward()
            res[name] = (loss.detach().double(), img.grad.double(), txt.grad.double())
        img, txt = img0.detach().clone().requires_grad_(True), txt0.detach().clone().requires_grad_(True)
        crit = HardNegativeMiningInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, temperature_schedule=True)
        man = ContrastiveLearningManager(None, crit)
        crit.set_epoch(epoch, 10)
        man.set_epoch(epoch, 10)
        loss = man.compute_loss(img, txt)
        loss.backward()
        l32, l64 = float(res["f32"][0]), float(res["f64"][0])
        assert abs(float(loss) - l64) <= 1e-5 * abs(l64) + 2 * abs(l32 - l64) + 1e-6, (float(loss), l32, l64)
        for got, i in ((img.grad, 1), (txt.grad, 2)):
            f32, f64 = res["f32"][i], res["f64"][i]
            bound = 1e-2 * f64.abs() + 1e-3 * float(f64.abs().max()) + 2 * float((f32 - f64).abs().max())
            miss = (got.double() - f64).abs() - bound
            assert float(miss.max()) <= 0, (b, e, i, int((miss > 0).sum()), float(miss.max()), float(f64.abs().max()))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_fused_loss_edge_cases():
    crit = HardNegativeMiningInfoNCE(temperature_schedule=False)
    # ratio 1.0: topk(k = B) keeps everything -> every negative is hard; B = 1: a single positive, loss = entropy terms only
    for b, ratio in ((8, 1.0), (1, 0.5), (2, 0.5)):
        crit.hardest_mining_ratio = ratio
        img = torch.randn(b, 32, device=DEV, requires_grad=True)
        txt = torch.randn(b, 32, device=DEV, requires_grad=True)
        loss = crit(img, txt)
        loss.backward()
        assert torch.isfinite(loss) and torch.isfinite(img.grad).all() and torch.isfinite(txt.grad).all()
    with pytest.raises(RuntimeError):
        crit(torch.randn(4, 8), torch.randn(4, 8))  # CPU tensors: no fallback
