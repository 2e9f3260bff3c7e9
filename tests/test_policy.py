"""CPU: host-side policy (atq/mixed_precision_atq.py mirror) against tables produced by the reference."""
import torch

import atq
import atq.mixed_precision_atq as mp


def test_quant_params_table(policy):
    for name, epoch, thr, ratio, sparsity, importance in policy["quant_params"]:
        assert mp.MixedPrecisionATQ.get_layer_importance(None, name) == importance
        r, s = mp.MixedPrecisionATQ.calculate_quantization_params(None, name, epoch, 10, thr)
        assert r == ratio and s == sparsity, (name, epoch, thr)


class _Dummy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.image_proj = atq.ResidualPrecisionBoostLinear(16, 8, precision_ratio=0.2, sparsity_target=0.1)
        self.text_attention = atq.ResidualPrecisionBoostLinear(16, 8, precision_ratio=0.2, sparsity_target=0.1)
        self.ffn = atq.ResidualPrecisionBoostLinear(16, 8)
        self.plain = atq.TernaryLinear(16, 8)


def test_scheduler_tables(policy):
    for tab in policy["scheduler"]:
        d = _Dummy()
        masks = [m.precision_mask.clone() for m in (d.image_proj, d.text_attention, d.ffn)]
        sch = mp.GradualQuantizationScheduler(d, tab["E"], 0.3, 0.2, warmup_epochs=tab["warmup"],
                                              final_epochs=tab["final"])
        assert sch.vision_sparsity_schedule == tab["vision"]
        assert sch.text_sparsity_schedule == tab["text"]
        for row in tab["rows"]:
            ep, v, t = row[0], row[1], row[2]
            assert list(sch.step(ep)) == [v, t]
            got = [[m.precision_ratio, m.sparsity_target] for m in (d.image_proj, d.text_attention, d.ffn)]
            assert got == row[3:], (tab["E"], ep)
        assert not hasattr(d.plain, "sparsity_target") and tab["plain_has_sparsity"] is False
        # writing precision_ratio is inert: masks unchanged
        for m, old in zip((d.image_proj, d.text_attention, d.ffn), masks):
            assert torch.equal(m.precision_mask, old)


def test_precision_controlled_linear(policy):
    pcl = mp.PrecisionControlledLinear(32, 16, importance=1.44)
    assert pcl.linear.precision_ratio == policy["pcl"]["ratio"]
    assert pcl.linear.sparsity_target == policy["pcl"]["sparsity"]
    assert sorted(pcl.state_dict().keys()) == policy["pcl"]["keys"]
    assert sorted(atq.TernaryLinear(4, 4).state_dict().keys()) == policy["tl_state_keys"]
    assert sorted(atq.ResidualPrecisionBoostLinear(4, 4).state_dict().keys()) == policy["rpb_state_keys"]


def test_init_matches_reference_rng_stream(golden):
    """Same seed -> same parameters and the same precision mask as the reference constructors."""
    torch.manual_seed(12)
    rpb = atq.ResidualPrecisionBoostLinear(64, 32, precision_ratio=0.05, sparsity_target=0.3)
    assert torch.equal(rpb.weight.detach(), torch.from_numpy(golden["rpb_weight"]))
    assert torch.equal(rpb.bias.detach(), torch.from_numpy(golden["rpb_bias"]))
    assert torch.equal(rpb.precision_mask, torch.from_numpy(golden["rpb_mask"]))
    torch.manual_seed(11)
    tl = atq.TernaryLinear(64, 32)
    assert torch.equal(tl.weight.detach(), torch.from_numpy(golden["tl_weight"]))
    assert torch.equal(tl.bias.detach(), torch.from_numpy(golden["tl_bias"]))


def test_enhanced_layer_structure():
    layer = mp.EnhancedATQTransformerLayer(64, 4, dim_feedforward=128, layer_idx=1, total_layers=4)
    names = [n for n, m in layer.named_modules() if isinstance(m, atq.ResidualPrecisionBoostLinear)]
    assert names == ["query.linear", "key.linear", "value.linear", "attn_out.linear", "ff1.linear", "ff2.linear"]
    imp = 1.0 + 1 / 3
    assert layer.query.linear.precision_ratio == min(0.25, 0.05 * (imp * 1.2))
    assert layer.ff1.linear.sparsity_target == max(0.1, 0.3 / (imp * 0.8))
