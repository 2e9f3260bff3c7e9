"""Data-parallel plumbing for ATQ training on one 8xB200 box (new work defined by BASELINE.json;
the reference is single-process).  One process per GPU over torch.distributed (NCCL on
NVLink 5 / NVSwitch; gloo in the CPU tests).  Exactly two exchange steps exist on this path:

  C2  all-gather of the [B_local, E] image and text embeddings before the contrastive loss
      (utils/enhanced_contrastive.py mines negatives over the whole batch);
  C1  one all-reduce (SUM) of the parameter gradients after backward.

Wiring (b) of SURVEY H9: the gather carries no gradient for remote rows, every rank evaluates the
identical global loss and back-propagates only through its own rows, so the SUM (not the mean)
of the per-rank parameter gradients is the gradient of the global-batch loss.
Weights are replicated and quantized redundantly (deterministic kernels => identical T on every
rank, no broadcast).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Initialise the default process group from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_embeddings(local: torch.Tensor, group=None) -> torch.Tensor:
    """[B_local, E] -> [B_local * world, E] in rank order; only this rank's rows carry autograd."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    src = local.detach().contiguous()
    out = torch.empty((world * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    dist.all_gather_into_tensor(out, src, group=group)
    parts: List[torch.Tensor] = list(out.split(src.shape[0], dim=0))
    parts[rank] = local
    return torch.cat(parts, dim=0)


class FlatGradAllReduce:
    """One flat fp32 buffer for every parameter gradient that the step produces, reduced with one NCCL
    all-reduce per bucket.  After backward the fresh gradients are packed into the buffer with a single
    multi-tensor copy, reduced, and `p.grad` is re-pointed at views of the buffer (no copy back).
    Parameters that never receive a gradient (TernaryLinear.weight, modules off the training path,
    SURVEY H8) keep grad=None so the optimizer skips them exactly as in the single-process reference.

    `sparse_masks`: {parameter: mask} for parameters whose gradient is exactly zero
    outside `mask != 0` on every rank - ResidualPrecisionBoostLinear.weight with its `precision_mask`
    (dW = G .* mask, atq/precision_boost.py:72; the mask is identical on all ranks).  Only the masked
    entries travel: they are gathered into the flat buffer, reduced, and scattered back into the dense
    `weight.grad` (whose other entries stay zero).  `rpb_masks(model)` builds the dictionary."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 256 << 20, group=None,
                 sparse_masks: Optional[dict] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.flat: Optional[torch.Tensor] = None
        self.active: List[torch.nn.Parameter] = []
        self.views: List[torch.Tensor] = []
        self.sparse_masks = {id(p): m for p, m in (sparse_masks or {}).items()}
        self.sparse_idx: List[Optional[torch.Tensor]] = []
        self._mask_versions = None

    def _versions(self):
        return tuple((id(m), m._version) for m in self.sparse_masks.values())

    def _bind(self):
        self.active = [p for p in self.params if p.grad is not None]
        self.sparse_idx = []
        for p in self.active:
            mask = self.sparse_masks.get(id(p))
            ok = mask is not None and p.is_contiguous() and mask.shape == p.shape
            self.sparse_idx.append(torch.nonzero(mask.reshape(-1) != 0).reshape(-1) if ok else None)
        self._mask_versions = self._versions()
        sizes = [p.numel() if idx is None else idx.numel() for p, idx in zip(self.active, self.sparse_idx)]
        # every slice starts on a 16-byte boundary (vectorised consumers such as atq.optim.FlatAdamW); the padding
        # elements stay zero and simply travel with the all-reduce
        self.flat = torch.zeros(sum((n + 3) // 4 * 4 for n in sizes), dtype=torch.float32, device=self.active[0].device)
        self.views = []
        off = 0
        for p, idx, n in zip(self.active, self.sparse_idx, sizes):
            piece = self.flat[off: off + n]
            if idx is not None:
                self.views.append(piece)  # compact: the masked entries only
            else:
                # same sizes AND strides as the parameter (channels-last conv weights stay channels-last:
                # fused optimizers require param/grad layouts to match); dense tensors only
                dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
                self.views.append(piece.as_strided(p.size(), p.stride()) if dense else piece.view_as(p))
            off += (n + 3) // 4 * 4

    def zero_grad(self):
        """Drop every gradient so autograd writes fresh tensors (no read-modify-write accumulation)."""
        for p in self.params:
            p.grad = None

    def reduce(self):
        if self.flat is None or (self.sparse_masks and self._mask_versions != self._versions()):
            self._bind()  # first step (or a mask was re-initialised): discover the gradients this graph produces
        dst, grads = [], []
        for p, view, idx in zip(self.active, self.views, self.sparse_idx):
            if p.grad is None:  # parameter unused this step: contributes zeros
                view.zero_()
            elif idx is not None:
                torch.index_select(p.grad.view(-1), 0, idx, out=view)  # gather the masked entries
            else:
                dst.append(view)
                grads.append(p.grad)
        if dst:
            torch._foreach_copy_(dst, grads)  # multi-tensor pack into the flat buffer
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            n = self.flat.numel()
            for start in range(0, n, self.bucket_elems):
                dist.all_reduce(self.flat[start: min(n, start + self.bucket_elems)], op=dist.ReduceOp.SUM, group=self.group)
        for p, view, idx in zip(self.active, self.views, self.sparse_idx):
            if idx is None:
                p.grad = view
            elif p.grad is not None:
                p.grad.view(-1).index_copy_(0, idx, view)  # scatter the reduced entries back (rest stays zero)


def rpb_masks(model: torch.nn.Module) -> dict:
    """{weight parameter: precision_mask} of every ResidualPrecisionBoostLinear-like module of `model`."""
    out = {}
    for m in model.modules():
        mask = getattr(m, "precision_mask", None)
        w = getattr(m, "weight", None)
        if mask is not None and isinstance(w, torch.nn.Parameter) and mask.shape == w.shape:
            out[w] = mask
    return out
