"""Gradient routing on B200 (drop-in for atq/routing.py)."""
import torch

from . import _engine as eng


def apply_selective_routing(input, threshold=0.05, importance_factor=0.3):
    """Identity (atq/routing.py:4-20): the same tensor object comes back, no copy, no kernel."""
    return input


class SelectiveGradientRouting(torch.autograd.Function):
    """Forward identity; backward keeps the gradient only where |input| is in the top
    `importance_factor` fraction (atq/routing.py:22-59): k = int((1-f)*numel); the threshold is the
    k-th smallest |input| (exact radix select, K2) if k < numel else 0; grad_in = grad_out*(|x|>thr)
    in one fused streaming kernel (K10)."""

    @staticmethod
    def forward(ctx, input, threshold=0.05, importance_factor=0.3):
        ctx.importance_factor = importance_factor
        ctx.save_for_backward(input)
        return input

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        n = input.numel()
        k = int((1 - ctx.importance_factor) * n)
        if k < n:
            if k < 1:
                # torch.kthvalue(importance.view(-1), 0) in the reference
                raise RuntimeError("kthvalue(): selected number k out of range for dimension 0")
            thr = eng.select_kth_abs(input.detach(), k - 1)
        else:
            thr = torch.zeros((), dtype=torch.float32, device=input.device)
        return eng.route_mask_mul(input.detach(), grad_output, thr).view(grad_output.shape), None, None
