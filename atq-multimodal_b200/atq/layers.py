"""TernaryLinear on B200 (drop-in for atq/layers.py:7-43)."""
from . import _engine as eng
from ._linear_base import TernaryLinearBase


class TernaryLinear(TernaryLinearBase):
    """y = x (alpha * T(W))^T + b, with T re-derived from the live fp32 weight.

    Parameters/state_dict are the reference's: weight [out,in], alpha [1], bias [out] or None.
    Sparsity target 0.3 and threshold_factor 0.05 are the quantizer defaults the reference's
    forward uses (atq/layers.py:37-40); the module deliberately has NO `sparsity_target`
    attribute and no `get_quantized_weights`, because callers duck-type on them.
    Gradients: input, alpha, bias; weight.grad stays None exactly like the reference.
    """

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self._declare_parameters(in_features, out_features, bias)
        self.reset_parameters()

    def reset_parameters(self):
        self._init_dense()

    def forward(self, input):
        return eng.ternary_linear(input, self.weight, self.alpha, self.bias, self._ops)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"
