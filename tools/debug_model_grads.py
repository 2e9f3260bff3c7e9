import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq
from atq.mixed_precision_atq import GradualQuantizationScheduler
from oracle import policy as P
from workloads import models as M
from workloads import train as T
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
cfg = T.RetrievalCfg(name="t", vocab=500, embed_dim=192, hidden_dim=384, image_size=64, batch=16)
ref, _, man_r = T.build_retrieval(M.oracle_layers(), cfg, seed=42)
mod, _, man_g = T.build_retrieval(atq, cfg, seed=42)
mod.load_state_dict(ref.state_dict()); mod.to(DEV)
ref64, _, man_64 = T.build_retrieval(M.oracle_layers(), cfg, seed=42)
ref64.load_state_dict(ref.state_dict()); ref64.double()
P.scheduler_step(ref, 5, 10, 0.3, 0.2, warmup_epochs=2)
P.scheduler_step(ref64, 5, 10, 0.3, 0.2, warmup_epochs=2)
GradualQuantizationScheduler(mod, 10, 0.3, 0.2, warmup_epochs=2).step(5)
for m in (ref, mod, ref64): m.eval()
images, captions, lengths = T.synthetic_batches(cfg, 1, seed=1)[0]
lr = man_r.compute_loss(*ref(images, captions, lengths)); lr.backward()
lg = man_g.compute_loss(*mod(images.to(DEV), captions.to(DEV), lengths.to(DEV))); lg.backward()
try:
    l64 = man_64.compute_loss(*ref64(images.double(), captions, lengths)); l64.backward()
    g64 = dict(ref64.named_parameters())
except Exception as e:
    print("fp64 ref failed:", e); g64 = None
print("loss cpu32", float(lr), "gpu", float(lg), "cpu64", float(l64) if g64 else None)
gr = dict(ref.named_parameters())
for n, p in mod.named_parameters():
    if p.grad is None: continue
    a, b = p.grad.cpu(), gr[n].grad
    s = float(b.abs().max()) + 1e-30
    e_gpu = float((a - b).abs().max()) / s
    if g64 is not None and g64[n].grad is not None:
        c = g64[n].grad.float()
        e_cpu32_vs64 = float((b - c).abs().max()) / s
        e_gpu_vs64 = float((a - c).abs().max()) / s
    else:
        e_cpu32_vs64 = e_gpu_vs64 = float("nan")
    if "text_encoder.layers" in n and ("norm" in n or "bias" in n): continue
    print(f"{n:60s} max|g|={s:10.3e} gpu-vs-cpu32={e_gpu:8.1e} cpu32-vs-64={e_cpu32_vs64:8.1e} gpu-vs-64={e_gpu_vs64:8.1e}")
