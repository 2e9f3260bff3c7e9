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
    `weight.grad` (whose other entries stay zero).  `rpb_masks(model)` builds the dictionary.

    `overlap=True` (dense gradients only): the flat buffer is laid out in REVERSE parameter order and cut into buckets
    of `bucket_bytes`; post-accumulate-grad hooks count the gradients of each bucket, and the moment a bucket is
    complete its gradients are packed (one multi-tensor copy) and all-reduced on a communication stream while
    autograd keeps running the rest of backward -- the bucketed, overlapped reduction of SURVEY section 5.  The comm
    stream first waits for an event on every stream that produced a gradient of the bucket (the two towers of the
    retrieval model run on different streams), `reduce()` launches whatever is left (parameters without a gradient
    this step contribute zeros), joins the comm stream and re-points `p.grad` at the flat views.  Works eagerly and
    under CUDA-graph capture (the comm stream becomes a branch of the captured graph).  The first step runs
    un-overlapped: it discovers which parameters receive gradients."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 256 << 20, group=None,
                 sparse_masks: Optional[dict] = None, overlap: bool = False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.overlap = bool(overlap) and not sparse_masks
        self.bucket_elems = max(1, bucket_bytes // 4)
        self._buckets: List[dict] = []       # overlap mode: {"lo", "hi", "idx": [active indices], "pending", "launched", "streams"}
        self._bucket_of: dict = {}           # id(param) -> bucket number
        self._hooks: list = []
        self._comm = None
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
        if self.overlap:
            self.active.reverse()  # gradients arrive roughly in reverse registration order: early buckets fill first
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
        if self.overlap:
            self._make_buckets(sizes)

    # ---- overlap mode ---------------------------------------------------------------------------------------
    def _make_buckets(self, sizes):
        for h in self._hooks:
            h.remove()
        self._hooks, self._buckets, self._bucket_of = [], [], {}
        off, cur = 0, None
        for i, (p, n) in enumerate(zip(self.active, sizes)):
            padded = (n + 3) // 4 * 4
            if cur is None or (off + padded - cur["lo"]) > self.bucket_elems and cur["idx"]:
                cur = {"lo": off, "hi": off, "idx": [], "pending": 0, "launched": False, "streams": {}}
                self._buckets.append(cur)
            cur["idx"].append(i)
            cur["hi"] = off + padded
            self._bucket_of[id(p)] = len(self._buckets) - 1
            off += padded
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        if self.flat.is_cuda:
            self._comm = torch.cuda.Stream(device=self.flat.device)
        self._arm()

    def _arm(self):
        for b in self._buckets:
            b["pending"], b["launched"], b["streams"] = len(b["idx"]), False, {}

    def _on_grad(self, p):
        b = self._buckets[self._bucket_of[id(p)]]
        if b["launched"]:
            return  # a second accumulation into the same parameter (shared weights): reduce() handles it
        if p.is_cuda:
            st = torch.cuda.current_stream(p.device)
            b["streams"][st.cuda_stream] = st
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b):
        b["launched"] = True
        dst, src = [], []
        for i in b["idx"]:
            p = self.active[i]
            if p.grad is None:
                self.views[i].zero_()
            else:
                dst.append(self.views[i])
                src.append(p.grad)
        piece = self.flat[b["lo"]: b["hi"]]
        distributed = dist.is_initialized() and dist.get_world_size(self.group) > 1
        if self._comm is not None:
            cur = torch.cuda.current_stream(self.flat.device)
            b["streams"][cur.cuda_stream] = cur
            for st in b["streams"].values():  # everything enqueued so far on the streams that produced these gradients
                self._comm.wait_event(st.record_event())
            with torch.cuda.stream(self._comm):
                if dst:
                    torch._foreach_copy_(dst, src)
                if distributed:
                    dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group)
        else:
            if dst:
                torch._foreach_copy_(dst, src)
            if distributed:
                dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group)

    def _reduce_overlapped(self):
        for b in self._buckets:
            if not b["launched"]:
                self._launch(b)
        if self._comm is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._comm)
        for p, view in zip(self.active, self.views):
            p.grad = view
        self._arm()

    def zero_grad(self):
        """Drop every gradient so autograd writes fresh tensors (no read-modify-write accumulation)."""
        for p in self.params:
            p.grad = None
        if self._buckets:
            self._arm()

    def reduce(self):
        if self.overlap and self.flat is not None:
            return self._reduce_overlapped()
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
