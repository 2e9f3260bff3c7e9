"""Fused attention core for the ternary transformer blocks.

`attention_core(q, k, v, num_heads, key_padding_mask, scale, dropout_p, training)` evaluates

    dropout(softmax(scale * q k^T + key mask)) v      per (batch, head)

on the [B, L, E] fp32 outputs of the ternary q/k/v projections in place (head h = columns 64h..64h+63),
replacing the reference's explicit matmul / masked_fill / softmax / dropout / matmul sequence
(models/text_encoder.py:117-163) and its autograd backward with `atq_attention_fwd` / `atq_attention_bwd`
(tcgen05 kernels, csrc/attention_sm100.cu).  Supported: head_dim 64, L <= 256 (`supported()`); there is no
fallback in here - callers keep the explicit torch sequence for other shapes.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _engine as eng
from . import _native as nv

HEAD_DIM = 64   # tile width: head dims that are a multiple of 8 up to 64 are supported (narrower heads are zero-padded in shared memory)
MAX_LEN = 256


def supported(embed_dim: int, num_heads: int, seq_len: int) -> bool:
    hd = embed_dim // max(num_heads, 1)
    return embed_dim == num_heads * hd and hd % 8 == 0 and 8 <= hd <= HEAD_DIM and 1 <= seq_len <= MAX_LEN


def _rows(t: torch.Tensor, name: str):
    """[B, L, E] fp32 CUDA tensor viewed as rows with a pitch (last dim contiguous, rows equally spaced)."""
    if t.dtype != torch.float32 or t.dim() != 3:
        raise RuntimeError(f"atq.attention: {name} must be a float32 [B, L, E] tensor")
    nv.device_index(t)
    b, l, e = t.shape
    ok = t.stride(2) == 1 and t.stride(0) == l * t.stride(1) and t.stride(1) % 4 == 0 and t.data_ptr() % 16 == 0
    if not ok:
        t = t.contiguous()
    return t, t.stride(1)


def _terms() -> int:
    return 3 if eng.get_gemm_mode() != "fast" else 1


class _AttentionCoreFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, v, num_heads, key_padding, scale, dropout_p, seed):
        q, qp = _rows(q, "q")
        k, kp = _rows(k, "k")
        v, vp = _rows(v, "v")
        b, l, e = q.shape
        if k.shape != q.shape or v.shape != q.shape:
            raise RuntimeError("atq.attention: q, k, v must have the same [B, L, E] shape (self-attention)")
        if not supported(e, num_heads, l):
            raise RuntimeError(f"atq.attention: unsupported shape E={e} heads={num_heads} L={l} (head_dim % 8 == 0, <= 64; L <= 256)")
        dev = nv.device_index(q)
        out = torch.empty((b, l, e), dtype=torch.float32, device=q.device)
        lse = torch.empty((b * num_heads, l), dtype=torch.float32, device=q.device)
        terms = _terms()
        nv.call("atq_attention_fwd", dev, b, num_heads, l, e // num_heads, q.data_ptr(), qp, k.data_ptr(), kp, v.data_ptr(), vp,
                nv.ptr(key_padding), float(scale), float(dropout_p), nv.ptr(seed), terms, out.data_ptr(), e,
                lse.data_ptr(), nv.stream_ptr(dev))
        ctx.save_for_backward(q, k, v, out, lse, key_padding, seed)
        ctx.cfg = (num_heads, float(scale), float(dropout_p), terms, qp, kp, vp)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        q, k, v, out, lse, key_padding, seed = ctx.saved_tensors
        num_heads, scale, dropout_p, terms, qp, kp, vp = ctx.cfg
        b, l, e = q.shape
        dev = nv.device_index(q)
        dout, dop = _rows(dout, "grad_output")
        dq = torch.empty((b, l, e), dtype=torch.float32, device=q.device)
        dk = torch.empty_like(dq)
        dv = torch.empty_like(dq)
        nv.call("atq_attention_bwd", dev, b, num_heads, l, e // num_heads, q.data_ptr(), qp, k.data_ptr(), kp, v.data_ptr(), vp,
                nv.ptr(key_padding), scale, dropout_p, nv.ptr(seed), terms, out.data_ptr(), e, dout.data_ptr(), dop,
                lse.data_ptr(), dq.data_ptr(), e, dk.data_ptr(), e, dv.data_ptr(), e, nv.stream_ptr(dev))
        return dq, dk, dv, None, None, None, None, None


def attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int,
                   key_padding_mask: Optional[torch.Tensor] = None, scale: Optional[float] = None,
                   dropout_p: float = 0.0, training: bool = True, seed: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q, k, v: [B, L, E] fp32 (E = num_heads * 64); key_padding_mask: [B, L] bool, True = ignore that key.
    Returns [B, L, E].  `seed` (int64 CUDA tensor [1]) pins the dropout mask; by default one is drawn from
    torch's CUDA generator (so `torch.manual_seed` governs it and CUDA-graph replays advance it)."""
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1] // int(num_heads))
    if scale <= 0:
        raise RuntimeError("atq.attention: scale must be positive")
    p = float(dropout_p) if training else 0.0
    pad = None
    if key_padding_mask is not None:
        if key_padding_mask.shape != q.shape[:2]:
            raise RuntimeError("atq.attention: key_padding_mask must be [B, L]")
        pad = key_padding_mask.to(torch.uint8).contiguous()
    if p > 0.0 and seed is None:
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=q.device)
    return _AttentionCoreFn.apply(q, k, v, int(num_heads), pad, scale, p, seed if p > 0.0 else None)


def dropout_keep_mask(seed: int, batch: int, num_heads: int, seq_len: int, dropout_p: float):
    """Host restatement of the kernels' counter-based dropout hash (tests): (keep bool [B, H, L, L], effective
    drop rate).  One 32-bit hash per pair of keys (2k, 2k+1) of a (batch, head, query) row: low / high 16 bits
    against the 16-bit threshold round(p * 65536)."""
    import numpy as np
    if dropout_p <= 0:
        return np.ones((batch, num_heads, seq_len, seq_len), dtype=bool), 0.0
    thresh = min(max(int(dropout_p * 65536.0 + 0.5), 1), 65535)
    seed_lo, seed_hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    pairs = (seq_len + 1) // 2
    with np.errstate(over="ignore"):
        row_id = np.arange(batch * num_heads * seq_len, dtype=np.uint32)
        x = (row_id * np.uint32(0x9E3779B1)) ^ seed_lo
        x ^= x >> np.uint32(16); x *= np.uint32(0x85EBCA6B); x ^= x >> np.uint32(13)
        x += seed_hi
        pair = np.arange(pairs, dtype=np.uint32)
        y = x[:, None] + pair[None, :] * np.uint32(0xC2B2AE35)
        y ^= y >> np.uint32(16); y *= np.uint32(0x7FEB352D); y ^= y >> np.uint32(15)
        y *= np.uint32(0x846CA68B); y ^= y >> np.uint32(16)
    keep = np.empty((row_id.size, 2 * pairs), dtype=bool)
    keep[:, 0::2] = (y & np.uint32(0xFFFF)) >= thresh
    keep[:, 1::2] = (y >> np.uint32(16)) >= thresh
    return keep[:, :seq_len].reshape(batch, num_heads, seq_len, seq_len), thresh / 65536.0


def dropout_keep_mask_flat(seed: int, numel: int, dropout_p: float, stream_id: int = 0x0FF1CE):
    """Host restatement of the flat-index dropout masks of csrc/streaming.cu: (keep bool [numel], effective rate).
    stream_id 0x0FF1CE = fused FFN activation (act_split_kernel), 0x6A7ED = gated residual."""
    import numpy as np
    if dropout_p <= 0:
        return np.ones(numel, dtype=bool), 0.0
    thresh = min(max(int(dropout_p * 65536.0 + 0.5), 1), 65535)
    seed_lo, seed_hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        x = (np.uint32(stream_id) * np.uint32(0x9E3779B1)) ^ seed_lo
        x ^= x >> np.uint32(16); x *= np.uint32(0x85EBCA6B); x ^= x >> np.uint32(13)
        key = x + seed_hi
        pair = np.arange((numel + 1) // 2, dtype=np.uint32)
        y = key + pair * np.uint32(0xC2B2AE35)
        y ^= y >> np.uint32(16); y *= np.uint32(0x7FEB352D); y ^= y >> np.uint32(15)
        y *= np.uint32(0x846CA68B); y ^= y >> np.uint32(16)
    keep = np.empty(2 * pair.size, dtype=bool)
    keep[0::2] = (y & np.uint32(0xFFFF)) >= thresh
    keep[1::2] = (y >> np.uint32(16)) >= thresh
    return keep[:numel], thresh / 65536.0
