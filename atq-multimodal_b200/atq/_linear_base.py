"""Shared parameter plumbing of the two ternary linear modules.

Both keep the reference's state: `weight [out, in]`, `alpha [1]`, `bias [out]` or None, registered in that
order, initialised like a dense linear layer (kaiming-uniform weight with a = sqrt(5), bias uniform in
+-1/sqrt(fan_in)) with alpha = 1, drawing from torch's RNG in the same order as the reference so that
`torch.manual_seed` reproduces its initial weights (atq/layers.py:27-33, atq/precision_boost.py:37-46)."""
import math

import torch
import torch.nn as nn

from . import _engine as eng


class TernaryLinearBase(nn.Module):
    def _declare_parameters(self, in_features: int, out_features: int, bias: bool) -> None:
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.alpha = nn.Parameter(torch.empty(1))
        self.register_parameter('bias', nn.Parameter(torch.empty(out_features)) if bias else None)
        self._ops = eng.LayerOperands()  # private packed / bf16 operand cache (never in the state_dict)

    def _init_dense(self) -> None:
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.constant_(self.alpha, 1.0)
        if self.bias is not None:
            limit = 1.0 / math.sqrt(self.weight.shape[1])
            nn.init.uniform_(self.bias, -limit, limit)
