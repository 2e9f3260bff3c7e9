"""ResidualPrecisionBoostLinear on B200 (drop-in for atq/precision_boost.py:9-92)."""
import torch

from . import _engine as eng
from ._linear_base import TernaryLinearBase
from .quantizers import adaptive_ternary_quantization


class ResidualPrecisionBoostLinear(TernaryLinearBase):
    """Ternary linear layer that keeps a fixed top-|W0| fraction of weights in fp32:

        W_mixed = T*alpha*(1 - mask) + W*mask ;  y = x W_mixed^T + b

    State is the reference's: weight, alpha, bias, buffer precision_mask (fp32 0/1, [out,in]),
    attributes precision_ratio / sparsity_target.  `sparsity_target` is read on every forward
    (external schedulers assign it); `precision_ratio` is only used by reset_parameters, so
    assigning it later is inert, as in the reference.
    """

    def __init__(self, in_features, out_features, precision_ratio=0.05, bias=True, sparsity_target=0.3):
        super().__init__()
        self.precision_ratio, self.sparsity_target = precision_ratio, sparsity_target
        self._declare_parameters(in_features, out_features, bias)
        self.register_buffer('precision_mask', torch.zeros(out_features, in_features))
        self.reset_parameters()

    def reset_parameters(self):
        self._init_dense()
        # The mask is STATE: 1.0 at the int(ratio * numel) largest |W0| (torch.topk on the flattened magnitudes, the
        # call atq/precision_boost.py:55-60 makes; its tie order is implementation-defined).
        with torch.no_grad():
            flat_abs = self.weight.detach().abs().reshape(-1)
            keep = torch.topk(flat_abs, int(self.precision_ratio * flat_abs.numel())).indices
            self.precision_mask.view(-1)[keep] = 1.0

    def forward(self, input):
        return eng.rpb_linear(input, self.weight, self.alpha, self.bias, self.precision_mask, self._ops,
                              sparsity_target=self.sparsity_target)

    def get_quantized_weights(self):
        """(ternary_weights fp32, alpha) for analysis / bit-packing (atq/precision_boost.py:76-92)."""
        return adaptive_ternary_quantization(self.weight, alpha=self.alpha, sparsity_target=self.sparsity_target)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"precision_ratio={self.precision_ratio}, sparsity_target={self.sparsity_target}")
