"""Drivers for the reference's OWN models / loss / schedulers (oracle/_ref).  TEST / BASELINE INFRASTRUCTURE ONLY.

Everything here calls the reference through its public names exactly as train.py / train_multimodal.py do
(file:line cited per function).  `oracle.ref_env.activate(...)` decides which `atq` the reference's models bind to:
the reference's own (CPU baseline / checker) or this repo's B200 package (drop-in arm).  The same functions run
on both sides, so a parity test compares like with like.
"""
from __future__ import annotations

import torch


def build_retrieval(vocab=3000, embed_dim=192, hidden_dim=384, seed=42):
    """train_multimodal.py:282-290 (use_residual=True, vision 0.3 / text 0.2: the config-2 flags)."""
    from models.multimodal_classifier import ATQMultimodalRetrieval
    torch.manual_seed(seed)
    return ATQMultimodalRetrieval(vocab_size=vocab, embed_dim=embed_dim, hidden_dim=hidden_dim, vision_threshold=0.3,
                                  text_threshold=0.2, use_residual=True)


def build_loss(model, epoch, total_epochs):
    """train_multimodal.py:334-346 + the per-epoch set_epoch calls at :437-438."""
    from utils.enhanced_contrastive import ContrastiveLearningManager, HardNegativeMiningInfoNCE
    criterion = HardNegativeMiningInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5,
                                          temperature_schedule=True)
    manager = ContrastiveLearningManager(model=model, criterion=criterion, similarity_threshold=0.7)
    criterion.set_epoch(epoch, total_epochs)
    manager.set_epoch(epoch, total_epochs)
    return criterion, manager


def step_schedule(model, epoch, total_epochs, warmup_epochs=2):
    """train_multimodal.py:350-357 + the INTENDED scheduler.step(epoch) of :442 (the script rebinds the name to the
    LR scheduler at :403; BASELINE config 2 says "gradual quant", so it is driven explicitly)."""
    from atq.mixed_precision_atq import GradualQuantizationScheduler
    GradualQuantizationScheduler(model, total_epochs, vision_sparsity=0.3, text_sparsity=0.2, warmup_epochs=warmup_epochs,
                                 verbose=False).step(epoch)


def make_optimizer(model, lr=5e-5):
    """train_multimodal.py:361-366."""
    return torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=1e-4, betas=(0.9, 0.98))


def retrieval_step(model, manager, optimizer, batch):
    """train_multimodal.py:467 (zero_grad), :540-585 (standard branch, no clip / distill / EMA: script defaults)."""
    images, captions, lengths = batch
    optimizer.zero_grad()
    img, txt = model(images, captions, lengths, return_embeddings=True)
    loss = manager.compute_loss(img, txt)
    loss.backward()
    optimizer.step()
    return loss


def synthetic_retrieval_batch(batch=16, image_size=160, vocab=3000, seq_len=50, min_len=5, max_len=20, seed=1):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, image_size, image_size, generator=g)
    captions = torch.randint(4, vocab, (batch, seq_len), generator=g)
    lengths = torch.randint(min_len, max_len, (batch,), generator=g)
    return images, captions, lengths


def retrieval_forward_backward(model, manager, batch):
    """Embeddings, loss and every parameter gradient of one evaluation-mode pass (dropout off, BatchNorm on running
    statistics: deterministic on every backend)."""
    model.eval()
    model.zero_grad(set_to_none=True)
    images, captions, lengths = batch
    feats = {}
    # the tensor the fp32 ResNet18 trunk hands to the ATQ part (models/multimodal_classifier.py:84-86): its gradient
    # is what the ternary projector's dX GEMM produces
    def _tap(mod, args):
        args[0].register_hook(lambda g: feats.__setitem__("dfeat", g.detach().cpu()))  # (returns None: input unchanged)

    hook = model.image_encoder.feature_norm.register_forward_pre_hook(_tap)
    img, txt = model(images, captions, lengths, return_embeddings=True)
    hook.remove()
    loss = manager.compute_loss(img, txt)
    loss.backward()
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    return {"img": img.detach().cpu(), "txt": txt.detach().cpu(), "loss": loss.detach().cpu(), "grads": grads,
            "dfeat": feats["dfeat"]}


def build_classifier(seed=0):
    """train.py:33-39 (config 1: --use-rpb, sparsity 0.3, hidden 128)."""
    from models.image_classifier import ATQImageClassifier
    torch.manual_seed(seed)
    return ATQImageClassifier(num_classes=10, input_channels=1, use_rpb=True, sparsity_target=0.3, hidden_size=128)


def classifier_forward_backward(model, x, y, sparsity):
    """train.py:146-149 (progressive sparsity through the attribute), :171-206 (CE -> backward)."""
    for m in model.modules():
        if hasattr(m, "sparsity_target"):
            m.sparsity_target = sparsity
    model.eval()
    model.zero_grad(set_to_none=True)
    logits = model(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    return {"logits": logits.detach().cpu(), "loss": loss.detach().cpu(), "grads": grads}


def build_block(embed_dim=768, num_heads=12, dim_feedforward=3072, seed=3):
    """models/text_encoder.py:166-249 -- the block BASELINE config 4's towers are made of."""
    from models.text_encoder import TernaryTransformerLayer
    torch.manual_seed(seed)
    return TernaryTransformerLayer(embed_dim, num_heads, dim_feedforward, dropout=0.1, use_rpb=True, sparsity_target=0.3)


def block_forward_backward(block, x, pad_mask, gy):
    block.eval()
    block.zero_grad(set_to_none=True)
    x = x.clone().requires_grad_(True)
    y = block(x, src_key_padding_mask=pad_mask)
    y.backward(gy)
    grads = {n: p.grad.detach().cpu() for n, p in block.named_parameters() if p.grad is not None}
    return {"y": y.detach().cpu(), "dx": x.grad.detach().cpu(), "grads": grads}
