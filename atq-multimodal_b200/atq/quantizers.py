"""Adaptive ternary quantization on B200 (drop-in for atq/quantizers.py:7-60).

Pipeline: exact radix-select threshold (K2) -> ternarize to fp32 {-1,0,+1} (K3), optional
optimal alpha from fused statistics (A3).  Everything stays on the device: the reference's
`if nonzero_count > 0` host sync (atq/quantizers.py:49) is resolved inside a kernel.
"""
import torch

from . import _engine as eng


def adaptive_ternary_quantization(weights, alpha=None, threshold_factor=0.05, sparsity_target=0.3):
    """Returns (w_ternary, alpha).

    w_ternary: fp32 tensor shaped like `weights`, values in {-1, 0, +1}, requires_grad False,
               bit-identical to the reference's output (strict compares against the k-th order
               statistic of |W|, k = int(sparsity_target * numel); ties and NaN map to 0).
    alpha:     the object passed in (returned untouched, as the reference does), or a 0-dim fp32
               tensor sum(W*T)/nnz (mean|W| when nnz == 0) when `alpha is None`.
    """
    if not torch.is_tensor(weights):
        raise TypeError("weights must be a tensor")
    w = weights.detach()
    if w.numel() == 0:
        raise RuntimeError("adaptive_ternary_quantization: empty weight tensor")
    thr = eng.adaptive_threshold(w, sparsity_target, threshold_factor)
    stats = None
    if alpha is None:
        stats = torch.zeros(16, dtype=torch.uint8, device=w.device)
    w_ternary = eng.ternarize_f32(w, thr, stats).view(weights.shape)
    if alpha is None:
        alpha = eng.optimal_alpha(w.contiguous(), stats)
    return w_ternary, alpha
