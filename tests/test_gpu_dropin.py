"""GPU parity ON THE REAL BOUNDARY: the reference's own, unmodified `models/` and `utils/` (staged copy
under oracle/_ref) import THIS repo's `atq` and run on the B200; the checker is the same code on the reference's own
`atq`, run by oracle/ref_runner.py in a process of its own on the CPU, in float32 (what the reference computes) and in
float64 (how far float32 itself is from the exact answer).

Criterion per tensor (north_star: rtol 1e-2 / atol 1e-3 on logits and gradients, against the reference fp32 path):

    |gpu - f64| <= 1e-2 * |f64| + 1e-3 * max(1, scale) ... plus twice the reference's own fp32 error on that tensor,

i.e. the B200 path may not be further from the exact answer than tolerance + 2x what the reference itself is off by.
`scale` is 1 for outputs (embeddings / logits: plain atol 1e-3) and the tensor's max magnitude for gradients (the
reference network's gradients span 1e-6 .. 1e3, an absolute 1e-3 is meaningless for them).
Two documented exceptions:
  * `*.alpha` gradients are ONE number each, the sum of 10^4..10^5 cancelling products G_ij T_ij (the reference's own
    fp32 result is up to 60 % off float64 on them): 8x the reference's own error instead of 2x;
  * the parameters of the fp32 ResNet18 trunk are torch / cuDNN code on both sides (cuDNN's fp32 convolution
    algorithms differ from the CPU's by ~1e-3 relative); they are not ATQ code and are not compared -- the gradient
    that ENTERS the trunk (d loss / d features, produced by the ternary projector's dX GEMM) is.
"""
import pytest
import torch

from conftest import have_staged_reference, run_reference

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not have_staged_reference(), reason="oracle/_ref not staged (python oracle/install_ref.py)")]

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _ieee_fp32_torch_ops():
    """The fp32 trunks / attention of the reference models are torch code, not the hot path: keep them IEEE fp32."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture(scope="module")
def b200_env():
    from oracle import ref_env
    ref_env.activate("b200")
    import atq
    import models.text_encoder as te
    assert te.ResidualPrecisionBoostLinear is atq.ResidualPrecisionBoostLinear, "reference models must bind to the B200 atq"
    assert "atq-multimodal_b200" in atq.__file__
    from oracle import ref_tasks
    return ref_tasks


def anchored_errors(name, gpu, f32, f64, relative_atol, ref_mult=2.0):
    """Returns None if the tensor meets the criterion, else a description of the miss."""
    gpu, f32, f64 = gpu.detach().cpu().double(), f32.double(), f64.double()
    ref_err = float((f32 - f64).abs().max())
    scale = float(f64.abs().max()) if relative_atol else 1.0
    bound = 1e-2 * f64.abs() + 1e-3 * scale + ref_mult * ref_err
    miss = (gpu - f64).abs() - bound
    worst = float(miss.max())
    if worst <= 0:
        return None
    return f"{name}: {int((miss > 0).sum())}/{miss.numel()} elements over by up to {worst:.3e} (|f64| max {float(f64.abs().max()):.3e}, fp32 own err {ref_err:.3e}, gpu err {float((gpu - f64).abs().max()):.3e})"


def compare_grads(got, res, skip=()):
    misses, checked = [], 0
    for n, g64 in res["f64"]["grads"].items():
        if any(n.endswith(s) for s in skip):
            continue
        assert n in got, f"{n}: the reference has a gradient, the B200 path has none"
        if float(g64.abs().max()) < 1e-12:
            continue
        m = anchored_errors(n, got[n], res["f32"]["grads"][n], g64, relative_atol=True,
                            ref_mult=8.0 if n.endswith(".alpha") else 2.0)
        checked += 1
        if m:
            misses.append(m)
    extra = set(got) - set(res["f64"]["grads"])
    assert not extra, f"gradients the reference does not produce: {sorted(extra)}"
    return misses, checked


def test_reference_retrieval_model_on_b200_atq(b200_env, tmp_path):
    """BASELINE config 2 at its real size: models/multimodal_classifier.py:102 ATQMultimodalRetrieval(3000, 192, 384),
    batch 16 x 3x160x160 / 50 tokens, reference GradualQuantizationScheduler at epoch 5 of 10, reference loss."""
    R = b200_env
    cfg = dict(vocab=3000, embed_dim=192, hidden_dim=384, seed=42, schedule_epoch=5, total_epochs=10, batch=16,
               image_size=160, data_seed=1, loss_epoch=5)
    res = run_reference("retrieval", cfg, tmp_path)
    assert "oracle/_ref/atq" in res["atq_file"]
    model = R.build_retrieval(cfg["vocab"], cfg["embed_dim"], cfg["hidden_dim"], cfg["seed"])
    model.load_state_dict(res["state"], strict=True)  # same keys / shapes / dtypes as the reference's checkpoint
    R.step_schedule(model, cfg["schedule_epoch"], cfg["total_epochs"])  # THIS repo's GradualQuantizationScheduler
    got_s = {n: float(m.sparsity_target) for n, m in model.named_modules() if hasattr(m, "sparsity_target")}
    assert got_s == res["sparsity"], "per-layer sparsity targets differ from the reference's scheduler"
    model.to(DEV)
    _, man = R.build_loss(model, cfg["loss_epoch"], cfg["total_epochs"])
    out = R.retrieval_forward_backward(model, man, tuple(t.to(DEV) for t in res["batch"]))
    misses = []
    for key in ("img", "txt"):
        assert torch.allclose(out[key], res["f32"][key], rtol=1e-2, atol=1e-3), key  # plain north_star tolerance
        m = anchored_errors(key, out[key], res["f32"][key], res["f64"][key], relative_atol=False)
        if m:
            misses.append(m)
    assert abs(float(out["loss"]) - float(res["f32"]["loss"])) <= 1e-3 + 1e-2 * abs(float(res["f32"]["loss"]))
    # k_proj.bias: softmax is invariant to a shift of all keys -> this gradient is exactly zero + rounding noise
    skip = ("k_proj.bias",) + tuple(n for n in res["f64"]["grads"] if n.startswith("image_encoder.base_model."))
    gm, checked = compare_grads(out["grads"], res, skip=skip)
    misses += gm
    m = anchored_errors("d loss / d trunk features", out["dfeat"], res["f32"]["dfeat"], res["f64"]["dfeat"], relative_atol=True)
    if m:
        misses.append(m)
    assert checked > 100
    assert not misses, "\n".join(misses)


def test_reference_image_classifier_on_b200_atq(b200_env, tmp_path):
    """BASELINE config 1: models/image_classifier.py:8 ATQImageClassifier(use_rpb=True), batch 256 x 1x28x28,
    sparsity written through the attribute as train.py:146-149 does (two points of the schedule)."""
    R = b200_env
    for sparsity in (0.05, 0.3):
        cfg = dict(seed=0, data_seed=0, batch=256, sparsity=sparsity)
        res = run_reference("classifier", cfg, tmp_path)
        model = R.build_classifier(cfg["seed"])
        model.load_state_dict(res["state"], strict=True)
        model.to(DEV)
        out = R.classifier_forward_backward(model, res["x"].to(DEV), res["y"].to(DEV), sparsity)
        assert torch.allclose(out["logits"], res["f32"]["logits"], rtol=1e-2, atol=1e-3)
        misses, checked = compare_grads(out["grads"], res)
        assert checked >= 10
        assert not misses, "\n".join(misses)
        # RPB gradient contract on the reference's own model: non-zero only under the mask
        for n, m in model.named_modules():
            if hasattr(m, "precision_mask"):
                assert float((m.weight.grad * (1 - m.precision_mask)).abs().max()) == 0.0, n


def test_reference_vitb_sized_block_on_b200_atq(b200_env, tmp_path):
    """The block BASELINE config 4 is made of, at its real width: models/text_encoder.py:166 TernaryTransformerLayer
    (768, 12 heads, FFN 3072) on 4 x 197 tokens with key padding; output, dX and every parameter gradient."""
    R = b200_env
    cfg = dict(embed_dim=768, num_heads=12, dim_feedforward=3072, seed=3, data_seed=4, batch=4, seq=197,
               pad_from=[None, 150, None, 90])
    res = run_reference("block", cfg, tmp_path)
    block = R.build_block(cfg["embed_dim"], cfg["num_heads"], cfg["dim_feedforward"], cfg["seed"])
    block.load_state_dict(res["state"], strict=True)
    block.to(DEV)
    out = R.block_forward_backward(block, res["x"].to(DEV), res["pad"].to(DEV), res["gy"].to(DEV))
    misses = []
    m = anchored_errors("y", out["y"], res["f32"]["y"], res["f64"]["y"], relative_atol=False)
    if m:
        misses.append(m)
    m = anchored_errors("dx", out["dx"], res["f32"]["dx"], res["f64"]["dx"], relative_atol=True)
    if m:
        misses.append(m)
    gm, checked = compare_grads(out["grads"], res, skip=("k_proj.bias",))
    assert checked >= 24
    assert not (misses + gm), "\n".join(misses + gm)
