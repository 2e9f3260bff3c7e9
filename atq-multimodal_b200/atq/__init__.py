"""atq -- B200-native (sm_100a) drop-in for the `atq` package of ak736/ATQ-Multimodal.

Same public names and signatures as the reference's atq/__init__.py:2-12; every tensor op
runs in hand-written CUDA kernels reached through the C ABI of libatq_sm100.so
(include/atq_sm100.h).  There is no CPU path: CPU tensors raise RuntimeError and a missing
library raises ImportError at import time.
"""
from . import _native  # noqa: F401  (loads libatq_sm100.so; ImportError if it was not built)
from .quantizers import adaptive_ternary_quantization
from .layers import TernaryLinear
from .routing import apply_selective_routing, SelectiveGradientRouting
from .precision_boost import ResidualPrecisionBoostLinear
from ._engine import (set_gemm_mode, get_gemm_mode, set_ste, set_packed_gemm, prepare_quantization,
                      notify_weights_changed)
from .attention import attention_core, supported as attention_core_supported
from ._engine import rpb_ffn as fused_ffn, rpb_ffn_supported as fused_ffn_supported  # GELU + dropout fused into the FFN
from ._engine import gated_residual, gated_residual_supported  # src + dropout(h) * gate in one pass
from ._engine import layer_norm, layer_norm_supported  # LayerNorm that also leaves max|y| for the operand split after it
from .contrastive import (HardNegativeMiningInfoNCE, ContrastiveLearningManager,
                          hard_negative_infonce)

__all__ = [
    'adaptive_ternary_quantization',
    'TernaryLinear',
    'SelectiveGradientRouting',
    'apply_selective_routing',
    'ResidualPrecisionBoostLinear',
]
