"""AdamW with the whole parameter set updated by ONE kernel launch (`atq_adamw_multi`).

Same update as `torch.optim.AdamW` (decoupled weight decay, bias correction, no amsgrad; the reference's
train_multimodal.py:361-366 optimizer).  torch's fused multi-tensor implementation reaches ~13 % of the copy
roofline on the ~120 mostly small tensors of the Flickr8k-shape model; this one walks a host-built table of
1024-element chunks and runs at the streaming rate.  Parameters whose `.grad` is None are skipped exactly like
torch does (TernaryLinear.weight, modules off the training path - SURVEY 8a G).  One step counter per parameter
group (torch keeps one per parameter: they differ only for a parameter whose first gradient arrives later than
its group's, which does not occur on this path).  CUDA-graph capturable: the step
counter lives on the device and the pointer table is uploaded from pinned host memory.
"""
from __future__ import annotations

import struct

import torch

from . import _native as nv

_CHUNK = 1024


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("FlatAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def load_state_dict(self, state_dict):
        """Resume: the moment buffers are replaced by the loaded ones, so every cached pointer table is stale; the
        per-group step counter comes back wherever torch.load put it and must live on the parameters' device."""
        super().load_state_dict(state_dict)
        self._tables = {}
        for group in self.param_groups:
            step = group.get("step")
            dev = next((p.device for p in group["params"]), None)
            if step is not None and dev is not None:
                group["step"] = torch.as_tensor(step, dtype=torch.float32).reshape(1).to(dev)

    def _table(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        if not ps:
            return None
        dev = ps[0].device
        if "step" not in group:
            group["step"] = torch.zeros(1, dtype=torch.float32, device=dev)
        for p in ps:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("FlatAdamW: parameters must be float32 CUDA tensors")
            st = self.state[p]
            if not st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        # every pointer the kernel dereferences is part of the key (a reloaded state or a re-allocated gradient
        # rebuilds the table); the cache holds no tensor references
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr())
                    for p in ps) + (group["step"].data_ptr(),)
        cached = self._tables.get(gi)
        if cached is not None and cached["key"] == key:
            return cached
        rows, chunk_tensor, chunk_off = [], [], []
        for i, p in enumerate(ps):
            st = self.state[p]
            g = p.grad
            if g.stride() != p.stride() or g.dtype != torch.float32:
                raise RuntimeError("FlatAdamW: a gradient does not share its parameter's memory layout")
            rows.append(struct.pack("<QQQQq", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
            nch = (p.numel() + _CHUNK - 1) // _CHUNK
            chunk_tensor += [i] * nch
            chunk_off += list(range(nch))
        host = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).pin_memory()
        meta = torch.tensor([chunk_tensor, chunk_off], dtype=torch.int32).pin_memory()
        table = torch.empty(host.numel(), dtype=torch.uint8, device=dev)
        meta_d = torch.empty_like(meta, device=dev)
        table.copy_(host, non_blocking=True)       # pinned -> device: legal inside a CUDA-graph capture
        meta_d.copy_(meta, non_blocking=True)
        cached = dict(key=key, table=table, meta=meta_d, host=(host, meta), n_chunks=len(chunk_tensor))
        self._tables[gi] = cached
        return cached

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            t = self._table(gi, group)
            if t is None:
                continue
            dev = nv.device_index(t["table"])
            b1, b2 = group["betas"]
            nv.call("atq_adamw_multi", dev, t["table"].data_ptr(), t["meta"][0].data_ptr(), t["meta"][1].data_ptr(), t["n_chunks"],
                    float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                    group["step"].data_ptr(), nv.stream_ptr(dev))
        return loss
