"""CPU oracle of the host-side quantization policy (atq/mixed_precision_atq.py:10-235 of the
reference), duck-typed so it can drive the oracle's CPU modules on bench.py's reference arm.
TEST INFRASTRUCTURE ONLY; pinned by tests/golden/policy_golden.json (tests/test_workloads.py)."""


def layer_importance(name, default=1.0):  # :17-46
    if any(k in name for k in ('fusion', 'cross_attention', 'projector', 'final')):
        return 2.0
    if any(k in name for k in ('attention', 'embed', 'pool')):
        return 1.5
    if any(k in name for k in ('intermediate', 'ffn', 'conv')):
        return 0.8
    return default


def quant_params(name, epoch, total_epochs, target_sparsity, initial_ratio=0.05):  # :82-112
    imp = layer_importance(name)
    ratio = min(0.25, initial_ratio * imp)
    final = max(0.1, target_sparsity / imp)
    progress = min(1.0, epoch / (total_epochs * 0.8))
    init = min(0.1, final)
    return ratio, init + progress * (final - init)


def sparsity_schedule(total_epochs, warmup_epochs, final_epochs, initial, final):  # :186-205
    final_epochs = final_epochs or max(2, int(total_epochs * 0.2))
    ramp = total_epochs - warmup_epochs - final_epochs
    out = [initial] * warmup_epochs
    out += [initial + ((i + 1) / ramp) * (final - initial) for i in range(ramp)]
    out += [final] * final_epochs
    return out


def scheduler_step(model, epoch, total_epochs, vision_sparsity=0.3, text_sparsity=0.2, warmup_epochs=5,
                   final_epochs=None):
    """GradualQuantizationScheduler.step (:207-235) + update_model_quantization (:115-145) on any
    model whose RPB modules are recognised by their `precision_mask` buffer."""
    vs = sparsity_schedule(total_epochs, warmup_epochs, final_epochs, 0.05, vision_sparsity)
    ts = sparsity_schedule(total_epochs, warmup_epochs, final_epochs, 0.05, text_sparsity)
    v, t = (vision_sparsity, text_sparsity) if epoch >= len(vs) else (vs[epoch], ts[epoch])
    if hasattr(model, 'set_epoch'):
        model.set_epoch(epoch, total_epochs)
    for name, module in model.named_modules():
        if hasattr(module, 'precision_mask') and hasattr(module, 'sparsity_target'):
            ratio, s = quant_params(name, epoch, total_epochs, v if 'image' in name else t)
            module.precision_ratio = ratio
            module.sparsity_target = s
    return v, t
