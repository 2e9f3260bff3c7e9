"""Which harness-level fusion costs gradient accuracy?  Harness config-2 model on the B200 atq vs the oracle model in
fp32 / fp64 (tests/test_gpu_models.py criterion), for combinations of the fusion switches."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import atq
from atq.mixed_precision_atq import GradualQuantizationScheduler
from oracle import policy as P
from workloads import models as M, train as T
from test_gpu_models import _anchored_misses

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
cfg = T.FLICKR8K_SHAPE  # BASELINE config 2 at its real size (image 160, vocab 3000, batch 16)
ref, _, man_r = T.build_retrieval(M.oracle_layers(), cfg, seed=42)
P.scheduler_step(ref, 5, 10, 0.3, 0.2, warmup_epochs=2)
ref.eval()
images, captions, lengths = T.synthetic_batches(cfg, 1, seed=1)[0]
state = {k: v.clone() for k, v in ref.state_dict().items()}

def run_cpu(dt):
    ref.zero_grad(set_to_none=True)
    i, t = ref(images.to(dt), captions, lengths)
    man_r.compute_loss(i, t).backward()
    return {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in ref.named_parameters()}
g32 = run_cpu(torch.float32)
ref.double()
g64 = run_cpu(torch.float64)
skip = ("k_proj.bias",) + tuple(n for n in g64 if n.startswith("image_encoder.base_model."))
for att, ffn, loss in ((False, False, False), (False, True, True), (True, True, True)):
    M.FUSED_ATTENTION_CORE, M.FUSED_FFN, T.FUSED_LOSS = att, ffn, loss
    mod, _, man_g = T.build_retrieval(atq, cfg, seed=42)
    mod.load_state_dict(state)
    mod.to(DEV).eval()
    GradualQuantizationScheduler(mod, 10, 0.3, 0.2, warmup_epochs=2).step(5)
    gi, gt = mod(images.to(DEV), captions.to(DEV), lengths.to(DEV))
    man_g.compute_loss(gi, gt).backward()
    grads = {n: p.grad for n, p in mod.named_parameters()}
    misses, checked, worst = _anchored_misses(grads, g32, g64, 2.0, 8.0, skip)
    print(f"attention={att} ffn/residual={ffn} loss={loss}: {len(misses)} of {checked} tensors miss, worst ratio {worst:.1f}", flush=True)
    for m in misses[:3]:
        print("    ", m)
