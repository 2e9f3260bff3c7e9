"""TernaryLinear on B200 (drop-in for atq/layers.py:7-43)."""
import math

import torch
import torch.nn as nn

from . import _engine as eng


class TernaryLinear(nn.Module):
    """y = x (alpha * T(W))^T + b, with T re-derived from the live fp32 weight.

    Parameters/state_dict are the reference's: weight [out,in], alpha [1], bias [out] or None.
    Sparsity target 0.3 and threshold_factor 0.05 are the quantizer defaults the reference's
    forward uses (atq/layers.py:37-40); the module deliberately has NO `sparsity_target`
    attribute and no `get_quantized_weights`, because callers duck-type on them.
    Gradients: input, alpha, bias; weight.grad stays None exactly like the reference.
    """

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.alpha = nn.Parameter(torch.empty(1))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter('bias', None)
        self._ops = eng.LayerOperands()  # private packed/bf16 operand cache (not in state_dict)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.constant_(self.alpha, 1.0)
        if self.bias is not None:
            fan_in = self.weight.shape[1]
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        return eng.ternary_linear(input, self.weight, self.alpha, self.bias, self._ops)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"
