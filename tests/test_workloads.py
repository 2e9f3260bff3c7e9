"""CPU: the bench harness models (workloads/) against the reference's own models and loss, run
with the oracle's CPU layers.  Needs /root/reference (build container); skipped on the GPU box."""
import sys
import types

import pytest
import torch

from conftest import REFERENCE, have_reference
from oracle import policy as P
from workloads import models as M
from workloads import train as T

needs_ref = pytest.mark.skipif(not have_reference(), reason="reference not mounted (GPU box)")


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "atq" or k.startswith("atq.")}
    sys.path.insert(0, REFERENCE)
    try:
        import importlib
        import torchvision.models as tvm
        mods = {}
        for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils")]:
            sys.modules.pop(k)
        orig = tvm.resnet18
        tvm.resnet18 = lambda weights=None, **kw: orig(weights=None, **kw)
        try:
            mods["mc"] = importlib.import_module("models.multimodal_classifier")
            mods["ic"] = importlib.import_module("models.image_classifier")
            mods["loss"] = importlib.import_module("utils.enhanced_contrastive")
            mods["mp"] = importlib.import_module("atq.mixed_precision_atq")
            mods["resnet_patch"] = (tvm, orig)
        except Exception:
            tvm.resnet18 = orig
            raise
        return mods, saved
    finally:
        sys.path.remove(REFERENCE)


def _restore(mods, saved):
    tvm, orig = mods["resnet_patch"]
    tvm.resnet18 = orig
    for k in [k for k in sys.modules if k == "atq" or k.startswith("atq.") or k.split(".")[0] in ("models", "utils")]:
        sys.modules.pop(k)
    sys.modules.update(saved)


@needs_ref
def test_retrieval_model_matches_reference(monkeypatch):
    # 1e-6 agreement with the reference model needs the reference's own matmul/softmax sequence in the
    # attention core (the fused SDPA evaluation differs in the last bits; checked separately below)
    monkeypatch.setattr(M, "FUSED_ATTENTION_CORE", False)
    mods, saved = _import_reference()
    try:
        torch.manual_seed(0)
        ref = mods["mc"].ATQMultimodalRetrieval(vocab_size=300, embed_dim=64, hidden_dim=128, vision_threshold=0.3,
                                                text_threshold=0.2, use_residual=True)
        mine = M.RetrievalModel(M.oracle_layers(), 300, 64, 128, 0.3, 0.2, True)
        missing, unexpected = mine.load_state_dict(ref.state_dict(), strict=False)
        assert missing == []
        assert all(k.startswith("fusion.") for k in unexpected)
        n_rpb = sum(1 for m in mine.modules() if hasattr(m, "precision_mask"))
        assert n_rpb == 29  # 28 executed per step + image_projector
        cfg = T.RetrievalCfg(name="t", vocab=300, embed_dim=64, hidden_dim=128, image_size=64, batch=4)
        images, captions, lengths = T.synthetic_batches(cfg, 1, seed=1)[0]
        # the intended gradual-quantization schedule, reference implementation vs oracle policy
        sch = mods["mp"].GradualQuantizationScheduler(ref, 10, 0.3, 0.2, warmup_epochs=2)
        for epoch in (0, 5, 9):
            v, t = sch.step(epoch)
            assert (v, t) == P.scheduler_step(mine, epoch, 10, 0.3, 0.2, warmup_epochs=2)
            want = {n: (m.precision_ratio, m.sparsity_target) for n, m in ref.named_modules()
                    if hasattr(m, "precision_mask") and not n.startswith("fusion.")}
            got = {n: (m.precision_ratio, m.sparsity_target) for n, m in mine.named_modules() if hasattr(m, "precision_mask")}
            assert got == want
            ref.eval(); mine.eval()
            with torch.no_grad():
                ri, rt = ref(images, captions, lengths, return_embeddings=True)
                mi, mt = mine(images, captions, lengths)
            assert torch.allclose(ri, mi, atol=1e-6) and torch.allclose(rt, mt, atol=1e-6)
        # training mode: same dropout stream, same loss and gradients
        ref.train(); mine.train()
        crit_r = mods["loss"].HardNegativeMiningInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5)
        man_r = mods["loss"].ContrastiveLearningManager(model=ref, criterion=crit_r, similarity_threshold=0.7)
        crit_m = M.HardNegativeInfoNCE()
        man_m = M.ContrastiveManager(crit_m)
        for ep in (0, 5, 9):
            for c in (crit_r, man_r, crit_m, man_m):
                c.set_epoch(ep, 10)
            torch.manual_seed(7)
            lr = man_r.compute_loss(*ref(images, captions, lengths, return_embeddings=True))
            torch.manual_seed(7)
            lm = man_m.compute_loss(*mine(images, captions, lengths))
            assert torch.allclose(lr, lm, atol=1e-6), (ep, float(lr), float(lm))
        ref.zero_grad(); mine.zero_grad()
        lr.backward(); lm.backward()
        gr = dict(ref.named_parameters())
        for n, p in mine.named_parameters():
            if p.grad is None:
                assert gr[n].grad is None, n
            else:
                assert torch.allclose(p.grad, gr[n].grad, atol=1e-5, rtol=1e-4), n
    finally:
        _restore(mods, saved)


@needs_ref
def test_classifier_and_loss_match_reference():
    mods, saved = _import_reference()
    try:
        torch.manual_seed(0)
        ref = mods["ic"].ATQImageClassifier(num_classes=10, input_channels=1, use_rpb=True, sparsity_target=0.3, hidden_size=128)
        mine = M.ImageClassifier(M.oracle_layers())
        mine.load_state_dict(ref.state_dict())
        ref.eval(); mine.eval()
        x = torch.randn(8, 1, 28, 28)
        with torch.no_grad():
            assert torch.allclose(ref(x), mine(x), atol=1e-6)
        # loss at a larger batch, all curriculum stages, with weights
        torch.manual_seed(3)
        a, b = torch.randn(33, 16), torch.randn(33, 16)
        crit_r = mods["loss"].HardNegativeMiningInfoNCE()
        man_r = mods["loss"].ContrastiveLearningManager(model=None, criterion=crit_r)
        man_m = M.ContrastiveManager(M.HardNegativeInfoNCE())
        for ep in (0, 4, 9, 12):
            crit_r.set_epoch(ep, 10); man_r.set_epoch(ep, 10)
            man_m.criterion.set_epoch(ep, 10); man_m.set_epoch(ep, 10)
            assert torch.allclose(man_r.compute_loss(a, b), man_m.compute_loss(a, b), atol=1e-6)
    finally:
        _restore(mods, saved)


def test_oracle_policy_against_golden(policy):
    for name, epoch, thr, ratio, sparsity, importance in policy["quant_params"]:
        assert P.layer_importance(name) == importance
        assert P.quant_params(name, epoch, 10, thr) == (ratio, sparsity)
    for tab in policy["scheduler"]:
        assert P.sparsity_schedule(tab["E"], tab["warmup"], tab["final"], 0.05, 0.3) == tab["vision"]
        assert P.sparsity_schedule(tab["E"], tab["warmup"], tab["final"], 0.05, 0.2) == tab["text"]


def test_cpu_port_step_runs():
    """The reference arm of bench.py: one optimisation step of the config-2 model on the CPU oracle."""
    cfg = T.RetrievalCfg(name="t", vocab=200, embed_dim=32, hidden_dim=64, image_size=32, batch=4)
    model, crit, man = T.build_retrieval(M.oracle_layers(), cfg)
    P.scheduler_step(model, cfg.epoch, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs)
    opt = T.make_optimizer(model, cfg)
    batch = T.synthetic_batches(cfg, 1, seed=0)[0]
    l0 = float(T.retrieval_step(model, man, opt, batch))
    l1 = float(T.retrieval_step(model, man, opt, batch))
    assert l0 == l0 and l1 == l1
    # gradient contract: unused modules keep grad None
    assert model.image_projector.weight.grad is None and model.temperature.grad is None


def test_fused_attention_core_matches_explicit_sequence():
    """workloads.models.FUSED_ATTENTION_CORE only changes HOW softmax(QK^T)V is evaluated."""
    import workloads.models as M
    from oracle import atq_oracle as O
    torch.manual_seed(0)
    layers = M.oracle_layers()
    att = M.TernaryAttention(layers, 48, 4, 0.1, True, 0.2).eval()
    x = torch.randn(3, 7, 48)
    pad = torch.zeros(3, 7, dtype=torch.bool)
    pad[0, 5:] = True
    pad[2, 3:] = True
    outs = []
    for fused in (True, False):
        M.FUSED_ATTENTION_CORE = fused
        try:
            outs.append(att(x, x, x, key_padding_mask=pad))
        finally:
            M.FUSED_ATTENTION_CORE = True
    assert torch.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-4)  # outputs reach ~1e2


def test_schedule_fingerprint_tracks_every_baked_scalar():
    """GraphedRetrievalStep re-captures when a host-side scalar baked into the capture moves (the reference changes
    them between epochs: train_multimodal.py:403-413 LR scheduler, set_epoch; mixed_precision_atq.py:323-401)."""
    cfg = T.RetrievalCfg(name="t", vocab=200, embed_dim=32, hidden_dim=64, image_size=32, batch=4)
    model, crit, man = T.build_retrieval(M.oracle_layers(), cfg)
    opt = T.make_optimizer(model, cfg)
    f0 = T.schedule_fingerprint(model, man, opt)
    assert f0 == T.schedule_fingerprint(model, man, opt)
    opt.param_groups[0]["lr"] *= 0.5
    f1 = T.schedule_fingerprint(model, man, opt)
    assert f1 != f0
    P.scheduler_step(model, cfg.epoch + 1, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs)
    f2 = T.schedule_fingerprint(model, man, opt)
    assert f2 != f1
    crit.set_epoch(cfg.epoch + 1, cfg.total_epochs)
    f3 = T.schedule_fingerprint(model, man, opt)
    assert f3 != f2
    man.set_epoch(cfg.total_epochs - 1, cfg.total_epochs)
    assert T.schedule_fingerprint(model, man, opt) != f3
    # a learning rate held in a device tensor is read at replay time: not part of the fingerprint
    opt.param_groups[0]["lr"] = torch.tensor(1e-4)
    f4 = T.schedule_fingerprint(model, man, opt)
    opt.param_groups[0]["lr"] = torch.tensor(2e-4)
    assert T.schedule_fingerprint(model, man, opt) == f4
