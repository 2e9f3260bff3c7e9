"""ResidualPrecisionBoostLinear on B200 (drop-in for atq/precision_boost.py:9-92)."""
import math

import torch
import torch.nn as nn

from . import _engine as eng
from .quantizers import adaptive_ternary_quantization


class ResidualPrecisionBoostLinear(nn.Module):
    """Ternary linear layer that keeps a fixed top-|W0| fraction of weights in fp32:

        W_mixed = T*alpha*(1 - mask) + W*mask ;  y = x W_mixed^T + b

    State is the reference's: weight, alpha, bias, buffer precision_mask (fp32 0/1, [out,in]),
    attributes precision_ratio / sparsity_target.  `sparsity_target` is read on every forward
    (external schedulers assign it); `precision_ratio` is only used by reset_parameters, so
    assigning it later is inert, as in the reference.
    """

    def __init__(self, in_features, out_features, precision_ratio=0.05, bias=True, sparsity_target=0.3):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.precision_ratio = precision_ratio
        self.sparsity_target = sparsity_target
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.alpha = nn.Parameter(torch.empty(1))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter('bias', None)
        self.register_buffer('precision_mask', torch.zeros(out_features, in_features))
        self._ops = eng.LayerOperands()
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.constant_(self.alpha, 1.0)
        if self.bias is not None:
            bound = 1 / math.sqrt(self.weight.shape[1])
            nn.init.uniform_(self.bias, -bound, bound)
        # the mask is STATE derived from the initial weights with the same torch.topk call as
        # atq/precision_boost.py:55-60 (tie order is implementation-defined)
        with torch.no_grad():
            magnitudes = self.weight.abs().view(-1)
            k = int(self.precision_ratio * magnitudes.numel())
            _, top = torch.topk(magnitudes, k)
            self.precision_mask.view(-1)[top] = 1.0

    def forward(self, input):
        return eng.rpb_linear(input, self.weight, self.alpha, self.bias, self.precision_mask, self._ops,
                              sparsity_target=self.sparsity_target)

    def get_quantized_weights(self):
        """(ternary_weights fp32, alpha) for analysis / bit-packing (atq/precision_boost.py:76-92)."""
        return adaptive_ternary_quantization(self.weight, alpha=self.alpha, sparsity_target=self.sparsity_target)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"precision_ratio={self.precision_ratio}, sparsity_target={self.sparsity_target}")
