"""Mixed-precision allocation and gradual quantization schedule (host-side policy).

Drop-in for atq/mixed_precision_atq.py.  No kernels live here: these classes only compute the
per-layer scalars (precision_ratio, sparsity_target) that feed the radix-select threshold of
each layer, so the arithmetic below uses exactly the reference's expressions (Python doubles,
same operation order) and is pinned by tests/golden/policy_golden.json.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import TernaryLinear
from .precision_boost import ResidualPrecisionBoostLinear
from .routing import apply_selective_routing
from .quantizers import adaptive_ternary_quantization  # noqa: F401  (re-exported like the reference)

# (keywords, importance) in priority order -- atq/mixed_precision_atq.py:30-44
_IMPORTANCE_RULES = (
    (('fusion', 'cross_attention', 'projector', 'final'), 2.0),
    (('attention', 'embed', 'pool'), 1.5),
    (('intermediate', 'ffn', 'conv'), 0.8),
)


class MixedPrecisionATQ:
    """Per-layer precision/sparsity allocation from name heuristics and training progress."""

    @staticmethod
    def get_layer_importance(model, layer_name, default_importance=1.0):
        for keywords, importance in _IMPORTANCE_RULES:
            if any(k in layer_name for k in keywords):
                return importance
        return default_importance

    @staticmethod
    def get_precision_ratio(importance, base_ratio=0.05, max_ratio=0.25):
        return min(max_ratio, base_ratio * importance)

    @staticmethod
    def get_sparsity_target(importance, base_sparsity=0.3, min_sparsity=0.1):
        return max(min_sparsity, base_sparsity / importance)

    @classmethod
    def calculate_quantization_params(cls, model, layer_name, epoch, total_epochs,
                                      target_sparsity, initial_ratio=0.05):
        """-> (precision_ratio, current_sparsity) for this layer at this epoch (:82-112)."""
        importance = cls.get_layer_importance(model, layer_name)
        precision_ratio = cls.get_precision_ratio(importance, base_ratio=initial_ratio)
        final_sparsity = cls.get_sparsity_target(importance, base_sparsity=target_sparsity)
        progress = min(1.0, epoch / (total_epochs * 0.8))
        initial_sparsity = min(0.1, final_sparsity)
        current_sparsity = initial_sparsity + progress * (final_sparsity - initial_sparsity)
        return precision_ratio, current_sparsity

    @staticmethod
    def update_model_quantization(model, epoch, total_epochs, vision_threshold=0.3, text_threshold=0.2):
        """Push the epoch's parameters into every RPB module of `model` (:115-145).  Modules whose
        qualified name contains 'image' are vision components, everything else is text."""
        if hasattr(model, 'set_epoch'):
            model.set_epoch(epoch, total_epochs)
        for name, module in model.named_modules():
            if not isinstance(module, ResidualPrecisionBoostLinear):
                continue
            threshold = vision_threshold if 'image' in name else text_threshold
            ratio, sparsity = MixedPrecisionATQ.calculate_quantization_params(
                model, name, epoch, total_epochs, threshold)
            module.precision_ratio = ratio      # inert after construction, as in the reference
            module.sparsity_target = sparsity   # takes effect on the next forward


class GradualQuantizationScheduler:
    """Warm-up / linear ramp / hold schedule of the vision and text sparsity targets (:148-235)."""

    def __init__(self, model, total_epochs, vision_sparsity=0.3, text_sparsity=0.2,
                 warmup_epochs=5, final_epochs=None, verbose=False):
        self.model = model
        self.total_epochs = total_epochs
        self.vision_sparsity = vision_sparsity
        self.text_sparsity = text_sparsity
        self.warmup_epochs = warmup_epochs
        self.final_epochs = final_epochs or max(2, int(total_epochs * 0.2))
        self.verbose = verbose
        self.initial_vision_sparsity = 0.05
        self.initial_text_sparsity = 0.05
        self.vision_sparsity_schedule = self._create_schedule(self.initial_vision_sparsity, self.vision_sparsity)
        self.text_sparsity_schedule = self._create_schedule(self.initial_text_sparsity, self.text_sparsity)

    def _create_schedule(self, initial_value, final_value):
        ramp = self.total_epochs - self.warmup_epochs - self.final_epochs
        schedule = [initial_value] * self.warmup_epochs
        schedule += [initial_value + ((i + 1) / ramp) * (final_value - initial_value) for i in range(ramp)]
        schedule += [final_value] * self.final_epochs
        return schedule

    def step(self, epoch):
        if epoch >= len(self.vision_sparsity_schedule):
            vision, text = self.vision_sparsity, self.text_sparsity
        else:
            vision, text = self.vision_sparsity_schedule[epoch], self.text_sparsity_schedule[epoch]
        MixedPrecisionATQ.update_model_quantization(
            self.model, epoch, self.total_epochs, vision_threshold=vision, text_threshold=text)
        if self.verbose:
            print(f"Epoch {epoch+1}: Vision sparsity = {vision:.3f}, Text sparsity = {text:.3f}")
        return vision, text


class PrecisionControlledLinear(nn.Module):
    """Linear layer whose precision ratio / sparsity are derived from an importance score (:238-285)."""

    def __init__(self, in_features, out_features, importance=1.0,
                 base_sparsity=0.3, base_precision_ratio=0.05, bias=True, use_rpb=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.importance = importance
        self.use_rpb = use_rpb
        if use_rpb:
            self.linear = ResidualPrecisionBoostLinear(
                in_features, out_features,
                precision_ratio=MixedPrecisionATQ.get_precision_ratio(importance, base_ratio=base_precision_ratio),
                sparsity_target=MixedPrecisionATQ.get_sparsity_target(importance, base_sparsity=base_sparsity),
                bias=bias)
        else:
            self.linear = TernaryLinear(in_features, out_features, bias=bias)

    def forward(self, x):
        return self.linear(x)


class EnhancedATQTransformerLayer(nn.Module):
    """Post-norm transformer block with q/k/v/out/ff1/ff2 all ternary, importance rising with depth
    (:289-402).  Every linear runs on the tcgen05 GEMM path through PrecisionControlledLinear."""

    def __init__(self, embed_dim, num_heads, dim_feedforward=2048, dropout=0.1,
                 use_rpb=True, base_sparsity=0.3, layer_idx=0, total_layers=4):
        super().__init__()
        self.layer_idx = layer_idx
        depth = layer_idx / max(1, total_layers - 1)
        layer_importance = 1.0 + depth
        attn_importance = layer_importance * 1.2
        ff_importance = layer_importance * 0.8

        def pcl(i, o, imp):
            return PrecisionControlledLinear(i, o, importance=imp, base_sparsity=base_sparsity, use_rpb=use_rpb)

        self.query = pcl(embed_dim, embed_dim, attn_importance)
        self.key = pcl(embed_dim, embed_dim, attn_importance)
        self.value = pcl(embed_dim, embed_dim, attn_importance)
        self.attn_out = pcl(embed_dim, embed_dim, attn_importance * 1.1)
        self.ff1 = pcl(embed_dim, dim_feedforward, ff_importance)
        self.ff2 = pcl(dim_feedforward, embed_dim, ff_importance * 1.2)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.dropout = nn.Dropout(dropout)
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"

    def _split_heads(self, t, batch_size):
        return t.view(batch_size, -1, self.num_heads, self.head_dim).transpose(1, 2)

    def _attention(self, q, k, v, mask=None):
        b = q.size(0)
        q, k, v = (self._split_heads(t, b) for t in (q, k, v))
        scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(self.head_dim)
        if mask is not None:
            scores = scores.masked_fill(mask == 0, float('-inf'))
        probs = self.dropout(F.softmax(scores, dim=-1))
        out = torch.matmul(probs, v)
        return out.transpose(1, 2).contiguous().view(b, -1, self.num_heads * self.head_dim)

    def forward(self, x, mask=None):
        q, k, v = self.query(x), self.key(x), self.value(x)
        threshold = max(0.01, 0.05 * (1.0 - self.layer_idx / 10))
        q = apply_selective_routing(q, threshold=threshold)
        k = apply_selective_routing(k, threshold=threshold)
        v = apply_selective_routing(v, threshold=threshold)
        x = self.norm1(x + self.dropout(self.attn_out(self._attention(q, k, v, mask))))
        ff = self.ff2(self.dropout(F.gelu(self.ff1(x))))
        return self.norm2(x + self.dropout(ff))
