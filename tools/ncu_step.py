"""ONE eager training step between cudaProfilerStart/Stop (ncu --profile-from-start off): the launch list of the
BASELINE config-2 (or config-4) step.   python tools/ncu_step.py [flickr8k|vitb16] [mode] [batch]"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch

import atq
from atq.mixed_precision_atq import GradualQuantizationScheduler
from atq.optim import FlatAdamW
from workloads import train as T

which = sys.argv[1] if len(sys.argv) > 1 else "flickr8k"
atq.set_gemm_mode(sys.argv[2] if len(sys.argv) > 2 else "parity")
cfg = T.VITB16 if which == "vitb16" else T.FLICKR8K_SHAPE
if len(sys.argv) > 3:
    cfg = dataclasses.replace(cfg, batch=int(sys.argv[3]))
dev = torch.device("cuda:0")
model, _, manager = T.build_retrieval(atq, cfg)
model.to(dev).train()
GradualQuantizationScheduler(model, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs).step(cfg.epoch)
opt = T.make_optimizer(model, cfg, fused=True, adamw_cls=FlatAdamW)
batches = [tuple(t.to(dev) for t in b) for b in T.synthetic_batches(cfg, 2, seed=42)]
for i in range(3):
    T.retrieval_step(model, manager, opt, batches[i % 2], prepare=atq.prepare_quantization)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = T.retrieval_step(model, manager, opt, batches[0], prepare=atq.prepare_quantization)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(loss.detach()))
