"""CPU oracle for the ATQ ternary hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy (integer / byte / index work) and CPU torch
(the fp32 linear algebra and its autograd graph), the algorithm of the reference's
``atq`` package.  It is the checker the CUDA path is compared against.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
(``atq-multimodal_b200/atq``) never does and fails loudly without its CUDA library.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
``tests/golden/gen_golden.py`` imports /root/reference, runs its functions on seeded
inputs and freezes inputs+outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function below against those files (and, when /root/reference is present,
against the live reference on fresh random inputs).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# A1-A3: adaptive ternary quantization (atq/quantizers.py:7-60)
# --------------------------------------------------------------------------------------


def threshold_index(numel: int, sparsity_target: float) -> int:
    """atq/quantizers.py:28 -- ``int(sparsity_target * sorted_weights.numel())`` (Python double)."""
    return int(sparsity_target * numel)


def adaptive_threshold(w: np.ndarray, sparsity_target: float = 0.3,
                       threshold_factor: float = 0.05) -> np.float32:
    """Per-layer threshold, atq/quantizers.py:21-38.

    k-th order statistic of |W| (ascending, 0-indexed k = int(s*n)) when 0 < k < n;
    max|W| + 1 when k >= n; threshold_factor * mean|W| when k == 0.
    """
    a = np.abs(np.asarray(w, dtype=np.float32)).reshape(-1)
    n = a.size
    k = threshold_index(n, sparsity_target)
    if 0 < k < n:
        # np.partition returns the same VALUE as a full sort at position k
        return np.float32(np.partition(a, k)[k])
    if k >= n:
        return np.float32(np.float32(a.max()) + np.float32(1.0))
    # fallback branch: fp32 mean as torch.mean computes it (pairwise fp32 sums);
    # the double-precision mean rounded to fp32 is the value the GPU path targets.
    return np.float32(np.float32(threshold_factor) * np.float32(a.astype(np.float64).mean()))


def ternarize(w: np.ndarray, thr) -> np.ndarray:
    """atq/quantizers.py:41-43 -- strict compares; ties, +-thr and NaN map to 0."""
    w = np.asarray(w, dtype=np.float32)
    thr = np.float32(thr)
    t = np.zeros(w.shape, dtype=np.int8)
    with np.errstate(invalid="ignore"):
        t[w > thr] = 1
        t[w < -thr] = -1
    return t


def optimal_alpha(w: np.ndarray, t: np.ndarray) -> np.float32:
    """atq/quantizers.py:46-55 -- sum(W*T)/nnz, or mean|W| when nnz == 0 (fp64 here)."""
    w64 = np.asarray(w, dtype=np.float64)
    nnz = int(np.count_nonzero(t))
    if nnz > 0:
        return np.float32((w64 * t).sum() / nnz)
    return np.float32(np.abs(w64).mean())


def adaptive_ternary_quantization(w: np.ndarray, alpha=None, threshold_factor: float = 0.05,
                                  sparsity_target: float = 0.3):
    """Whole function, atq/quantizers.py:7-60.  Returns (T int8, alpha, thr)."""
    thr = adaptive_threshold(w, sparsity_target, threshold_factor)
    t = ternarize(w, thr)
    if alpha is None:
        alpha = optimal_alpha(w, t)
    return t, alpha, thr


# --------------------------------------------------------------------------------------
# E1-E3: 2-bit codec (atq/bit_packing.py:22-146)
# --------------------------------------------------------------------------------------


def pack2(t: np.ndarray) -> np.ndarray:
    """atq/bit_packing.py:45-69.  code = value + 1; element i sits at bits
    2*(i%4)..2*(i%4)+1 of byte i//4; flat row-major order; tail bits zero.
    Raises ValueError on non-ternary input like atq/bit_packing.py:36-39."""
    flat = np.asarray(t).reshape(-1)
    f32 = flat.astype(np.float32)
    ok = (f32 == -1.0) | (f32 == 0.0) | (f32 == 1.0)
    if not bool(ok.all()):
        raise ValueError("Input must contain only ternary values (-1, 0, 1)")
    n = flat.size
    code = (f32 + 1.0).astype(np.uint8)
    pad = (-n) % 4
    if pad:
        code = np.concatenate([code, np.zeros(pad, np.uint8)])
    c = code.reshape(-1, 4)
    out = c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)
    return out.astype(np.uint8)


def unpack2(packed: np.ndarray, num_values: int) -> np.ndarray:
    """atq/bit_packing.py:104-119.  Code 3 has no entry in the reference's encoding
    table (KeyError at :116); the oracle raises the same KeyError."""
    p = np.asarray(packed, dtype=np.uint8).reshape(-1)
    codes = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[:num_values]
    if (codes == 3).any():
        raise KeyError(3)
    return codes.astype(np.float32) - np.float32(1.0)


def compute_memory_savings(numel: int) -> dict:
    """atq/bit_packing.py:122-146 (pure arithmetic)."""
    original_bytes = numel * 4
    packed_bytes = (numel * 2 + 7) // 8
    return {
        "original_bytes": original_bytes,
        "packed_bytes": packed_bytes,
        "compression_ratio": original_bytes / packed_bytes,
        "memory_reduction": 1.0 - (packed_bytes / original_bytes),
    }


def fast_ternary_matmul(packed: np.ndarray, shape, x: np.ndarray, alpha: float = 1.0) -> np.ndarray:
    """atq/bit_packing.py:149-176: (x @ unpack(p).T) * alpha, fp32."""
    wt = unpack2(packed, int(np.prod(shape))).reshape(shape)
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return (torch.matmul(xt, torch.from_numpy(wt).t()) * alpha).numpy()


# --------------------------------------------------------------------------------------
# D2: selective gradient routing backward (atq/routing.py:36-59)
# --------------------------------------------------------------------------------------


def routing_backward(x: np.ndarray, grad_out: np.ndarray, importance_factor: float = 0.3) -> np.ndarray:
    """k = int((1-f)*numel); thr = k-th smallest |x| (1-indexed) if k < numel else 0;
    grad_in = grad_out * (|x| > thr).  k == 0 raises like torch.kthvalue does."""
    a = np.abs(np.asarray(x, dtype=np.float32)).reshape(-1)
    k = int((1 - importance_factor) * a.size)
    if k < a.size:
        if k < 1:
            raise RuntimeError("kthvalue(): selected number k out of range for dimension 0")
        thr = np.float32(np.partition(a, k - 1)[k - 1])
    else:
        thr = np.float32(0.0)
    mask = (np.abs(np.asarray(x, dtype=np.float32)) > thr).astype(np.float32)
    return np.asarray(grad_out, dtype=np.float32) * mask


# --------------------------------------------------------------------------------------
# C1: precision mask (atq/precision_boost.py:49-60)
# --------------------------------------------------------------------------------------


def precision_mask_from_weight(weight: torch.Tensor, precision_ratio: float) -> torch.Tensor:
    """mask = 1.0 at the indices torch.topk(|W0|.flatten(), int(ratio*n)) returns.
    topk tie order is implementation-defined, so the oracle calls the same
    torch.topk the reference calls."""
    flat = weight.detach().abs().reshape(-1)
    k = int(precision_ratio * flat.numel())
    mask = torch.zeros_like(flat)
    if k > 0:
        _, idx = torch.topk(flat, k)
        mask[idx] = 1.0
    return mask.reshape(weight.shape)


# --------------------------------------------------------------------------------------
# B1/B2, C2/C3: the layers as CPU torch modules (graph identical to the reference's, so
# autograd reproduces its gradient contract: T carries no grad path to W).
# Used by tests as the fp32 checker and by bench.py's reference arm as the CPU "port".
# --------------------------------------------------------------------------------------


def _quantize_torch(weight: torch.Tensor, sparsity_target: float, threshold_factor: float = 0.05):
    """atq/quantizers.py:21-43 with torch CPU ops (sort-based, like the reference)."""
    a = weight.detach().abs().reshape(-1)
    n = a.numel()
    k = threshold_index(n, sparsity_target)
    if 0 < k < n:
        thr = torch.sort(a).values[k]
    elif k >= n:
        thr = a.max() + 1.0
    else:
        thr = threshold_factor * a.mean()
    wd = weight.detach()
    t = torch.zeros_like(wd)
    t[wd > thr] = 1.0
    t[wd < -thr] = -1.0
    return t, thr


class OracleTernaryLinear(nn.Module):
    """atq/layers.py:7-43.  y = x (T*alpha)^T + b with s = 0.3 fixed (no attribute)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.alpha = nn.Parameter(torch.empty(1))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.constant_(self.alpha, 1.0)
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        t, _ = _quantize_torch(self.weight, 0.3)
        return F.linear(x, t * self.alpha, self.bias)


class OracleRPBLinear(nn.Module):
    """atq/precision_boost.py:9-92.  W_mixed = T*alpha*(1-M) + W*M."""

    def __init__(self, in_features, out_features, precision_ratio=0.05, bias=True, sparsity_target=0.3):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.precision_ratio = precision_ratio
        self.sparsity_target = sparsity_target
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.alpha = nn.Parameter(torch.empty(1))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.register_buffer("precision_mask", torch.zeros(out_features, in_features))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.constant_(self.alpha, 1.0)
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            nn.init.uniform_(self.bias, -bound, bound)
        with torch.no_grad():
            flat = self.weight.abs().view(-1)
            k = int(self.precision_ratio * flat.numel())
            _, idx = torch.topk(flat, k)
            self.precision_mask.view(-1)[idx] = 1.0

    def get_quantized_weights(self):
        t, _ = _quantize_torch(self.weight, self.sparsity_target)
        return t, self.alpha

    def forward(self, x):
        t, _ = _quantize_torch(self.weight, self.sparsity_target)
        m = self.precision_mask
        w_mixed = t * self.alpha * (1 - m) + self.weight * m
        return F.linear(x, w_mixed, self.bias)


def ternary_linear_reference(x, weight, alpha, bias, sparsity_target=0.3, mask=None,
                             grad_out=None, dtype=torch.float64):
    """Closed-form forward/backward of B1/B2 (mask None) or C2/C3 (mask given) in `dtype`
    (fp64 by default: the tolerance tests compare the GPU path to the exact answer as well
    as to the fp32 module above).  Returns dict(y, dx, dw, dalpha, dbias)."""
    t, _ = _quantize_torch(weight.float(), sparsity_target)
    x_, w_, a_ = x.to(dtype), weight.to(dtype), alpha.to(dtype)
    t_ = t.to(dtype)
    if mask is None:
        wm = t_ * a_
        tq = t_
    else:
        m_ = mask.to(dtype)
        wm = t_ * a_ * (1 - m_) + w_ * m_
        tq = t_ * (1 - m_)
    x2 = x_.reshape(-1, x_.shape[-1])
    y = x2 @ wm.t()
    if bias is not None:
        y = y + bias.to(dtype)
    out = {"y": y.reshape(*x.shape[:-1], weight.shape[0])}
    if grad_out is not None:
        g2 = grad_out.to(dtype).reshape(-1, weight.shape[0])
        G = g2.t() @ x2
        out["dx"] = (g2 @ wm).reshape(x.shape)
        out["dalpha"] = (G * tq).sum().reshape(1)
        out["dw"] = None if mask is None else G * mask.to(dtype)
        out["dbias"] = g2.sum(0)
    return out
