"""Synthetic training steps for the BASELINE.json configs (bench / test harness).

Step bodies follow the reference's loops on synthetic tensors of the named shapes
(SURVEY.md 8d): train.py:170-212 (config 1) and train_multimodal.py:540-585 (configs 2 and 4),
with GradualQuantizationScheduler.step driven explicitly (the reference rebinds `scheduler` to the
LR scheduler at train_multimodal.py:403, so its quantization schedule never runs -- SURVEY 3.3;
BASELINE config 2 says "gradual quant", so the harness applies the intended behaviour)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import models as M


# True: configs built on the B200 `atq` use its fused contrastive loss; False: the torch-op restatement in models.py
FUSED_LOSS = True


@dataclass
class RetrievalCfg:
    name: str
    vocab: int = 3000
    embed_dim: int = 192
    hidden_dim: int = 384
    image_size: int = 160
    seq_len: int = 50
    batch: int = 16            # per-GPU batch
    image_tower: str = "resnet18"
    text_heads: int = 8
    text_layers: int = 4
    total_epochs: int = 10
    epoch: int = 5             # mid-schedule: per-layer sparsities differ
    warmup_epochs: int = 2
    lr: float = 5e-5
    min_len: int = 5
    max_len: int = 20


FLICKR8K_SHAPE = RetrievalCfg(name="train_multimodal.py synthetic Flickr8k shape: image 160, embed 192, hidden 384, "
                                   "batch 16, gradual quant + residual")
VITB16 = RetrievalCfg(name="ViT-B/16-sized ternary image encoder + 12-layer ternary text encoder, contrastive "
                           "batch 512 per GPU (4096 across 8xB200)",
                      embed_dim=768, hidden_dim=3072, image_size=224, batch=512, image_tower="vit",
                      text_heads=12, text_layers=12, max_len=50)


def build_retrieval(layers, cfg: RetrievalCfg, seed=42):
    torch.manual_seed(seed)  # identical init on every rank (SURVEY 8e determinism)
    model = M.RetrievalModel(layers, cfg.vocab, cfg.embed_dim, cfg.hidden_dim, 0.3, 0.2, True,
                             image_tower=cfg.image_tower, text_heads=cfg.text_heads, text_layers=cfg.text_layers,
                             max_seq_length=cfg.seq_len,
                             vit_cfg=dict(embed_dim=cfg.embed_dim, depth=12, num_heads=12, dim_feedforward=cfg.hidden_dim,
                                          image_size=cfg.image_size) if cfg.image_tower == "vit" else None)
    fused = getattr(layers, "HardNegativeMiningInfoNCE", None) if FUSED_LOSS else None
    if fused is not None:  # the B200 package's fused loss (atq/contrastive.py); the CPU oracle namespace has none
        criterion = fused(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5, temperature_schedule=True)
        manager = layers.ContrastiveLearningManager(None, criterion)
    else:
        criterion = M.HardNegativeInfoNCE(temperature=0.07, lambda_reg=0.02, hard_negative_weight=0.5,
                                          temperature_schedule=True)
        manager = M.ContrastiveManager(criterion)
    criterion.set_epoch(cfg.epoch, cfg.total_epochs)
    manager.set_epoch(cfg.epoch, cfg.total_epochs)
    return model, criterion, manager


def make_optimizer(model, cfg: RetrievalCfg, capturable=False, fused=False, adamw_cls=None):
    # train_multimodal.py:361-366 (AdamW, betas (0.9, 0.98)); `fused` only selects torch's single-kernel
    # CUDA implementation of the same update; `adamw_cls` (atq.optim.FlatAdamW) the one-launch own kernel
    if adamw_cls is not None:
        return adamw_cls(model.parameters(), lr=cfg.lr, weight_decay=1e-4, betas=(0.9, 0.98))
    kw = dict(fused=True) if fused else {}
    return torch.optim.AdamW(model.parameters(), lr=cfg.lr, weight_decay=1e-4, betas=(0.9, 0.98),
                             capturable=capturable, **kw)


def synthetic_batches(cfg: RetrievalCfg, count, seed, pin=False):
    """`count` host batches (images fp32, captions int64, lengths int64) of the config's shape."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(count):
        images = torch.randn(cfg.batch, 3, cfg.image_size, cfg.image_size, generator=g)
        captions = torch.randint(4, cfg.vocab, (cfg.batch, cfg.seq_len), generator=g)
        lengths = torch.randint(cfg.min_len, cfg.max_len, (cfg.batch,), generator=g)
        if pin:
            images, captions, lengths = images.pin_memory(), captions.pin_memory(), lengths.pin_memory()
        out.append((images, captions, lengths))
    return out


def retrieval_step(model, manager, optimizer, batch, gather=None, grad_sync=None, prepare=None):
    """One optimisation step; returns the loss tensor (still on the device)."""
    images, captions, lengths = batch
    if grad_sync is not None:
        grad_sync.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    if prepare is not None:
        if getattr(model, "parallel_towers", False) and images.is_cuda:
            model.prepare_fn = prepare  # each tower re-quantizes its own layers on its own stream (models.py)
        else:
            prepare(model)  # batched re-quantization of every layer the last optimizer step touched
    img, txt = model(images, captions, lengths, return_embeddings=True)
    if gather is not None:
        img, txt = gather(img), gather(txt)
    loss = manager.compute_loss(img, txt)
    loss.backward()
    if grad_sync is not None:
        grad_sync.reduce()
    optimizer.step()
    return loss


# ---- config 1 -------------------------------------------------------------------------

def build_classifier(layers, seed=0):
    torch.manual_seed(seed)
    return M.ImageClassifier(layers, num_classes=10, input_channels=1, use_rpb=True, sparsity_target=0.3, hidden_size=128)


def classifier_batches(count, seed=0, batch=256, pin=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(count):
        x = torch.randn(batch, 1, 28, 28, generator=g)
        y = torch.randint(0, 10, (batch,), generator=g)
        out.append((x.pin_memory(), y.pin_memory()) if pin else (x, y))
    return out


def classifier_step(model, optimizer, batch):
    x, y = batch
    optimizer.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x), y)
    loss.backward()
    optimizer.step()
    return loss


# ---- CUDA-graph capture of the whole optimisation step ----------------------------------

def schedule_fingerprint(model, manager, optimizer):
    """Every host-side scalar a captured step bakes in (see GraphedRetrievalStep): a tuple that changes exactly when
    a replay would train with a stale value.  ~30 us for config 2's 28 layers."""
    items = []
    for m in model.modules():
        st = getattr(m, "sparsity_target", None)
        if st is not None and not isinstance(st, torch.Tensor):
            items.append(float(st))
    for g in optimizer.param_groups:
        for key in ("lr", "betas", "eps", "weight_decay"):
            v = g.get(key)
            if not isinstance(v, torch.Tensor):  # a device tensor is read by the kernel at replay time: nothing baked
                items.append(v)
    for obj in (manager, getattr(manager, "criterion", None)):
        for key in ("epoch", "total_epochs", "curriculum_stage", "current_epoch", "temperature", "base_temperature",
                    "lambda_reg", "hard_negative_weight", "hardest_mining_ratio", "temperature_schedule"):
            if obj is not None and hasattr(obj, key):
                items.append(getattr(obj, key))
    try:
        import atq
        items.append(atq.get_gemm_mode())
    except (ImportError, AttributeError):
        pass
    return tuple(items)


class GraphedRetrievalStep:
    """Captures zero_grad + forward + loss + backward (+ gradient all-reduce) + AdamW into ONE CUDA
    graph.  The small-shape configs are launch-bound (~1000 kernels per step, most a few
    microseconds); the hot path was written without host synchronisation (thresholds, alpha and the
    validation flags stay on the device) precisely so that it can be captured.  Replays re-run the
    per-layer quantization kernels on the live weights, exactly like the eager step.

    Host-side scalars are baked into a capture: every layer's k = int(sparsity_target * numel) and mode, the
    optimizer's lr / betas / eps / weight decay, the loss temperature and curriculum stage, the GEMM mode.  The
    reference changes them between epochs (GradualQuantizationScheduler.step, MixedPrecisionATQ, LR schedulers,
    set_epoch), so `__call__` compares a fingerprint of those scalars with the one taken at capture and RE-CAPTURES when
    it moved (`recaptures` counts them) - a replay never trains with a stale schedule."""

    def __init__(self, model, manager, optimizer, example_batch, gather=None, grad_sync=None, warmup=3, prepare=None):
        self.model, self.manager, self.optimizer = model, manager, optimizer
        self.gather, self.grad_sync, self.prepare = gather, grad_sync, prepare
        self.static_batch = tuple(t.clone() for t in example_batch)
        self.recaptures = 0
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                retrieval_step(model, manager, optimizer, self.static_batch, gather, grad_sync, prepare)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._capture()

    def _capture(self):
        self.graph = torch.cuda.CUDAGraph()
        if self.grad_sync is None:
            self.optimizer.zero_grad(set_to_none=True)
        try:
            import atq._native as nv
            k0 = nv.kernel_launch_count()
        except ImportError:
            nv, k0 = None, 0
        self.fingerprint = schedule_fingerprint(self.model, self.manager, self.optimizer)
        with torch.cuda.graph(self.graph):
            self.static_loss = retrieval_step(self.model, self.manager, self.optimizer, self.static_batch, self.gather,
                                              self.grad_sync, self.prepare)
        # kernels of libatq_sm100 recorded in the graph (each replay launches them again)
        self.own_kernels_per_replay = (nv.kernel_launch_count() - k0) if nv is not None else 0

    def __call__(self, batch):
        if schedule_fingerprint(self.model, self.manager, self.optimizer) != self.fingerprint:
            torch.cuda.synchronize()   # the old graph's replays must be finished before its memory pool is dropped
            self.graph = None
            self.recaptures += 1
            self._capture()
        for dst, src in zip(self.static_batch, batch):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss

    def release(self):
        """Back to eager execution: graph replays update the weights without bumping tensor versions,
        so the layers' operand caches must be dropped."""
        for m in self.model.modules():
            ops = getattr(m, "_ops", None)
            if ops is not None:
                ops.key = None
