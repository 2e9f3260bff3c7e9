"""Kernel-time breakdown of one training step (torch.profiler, CUDA activity only): which kernels
outside libatq_sm100 the step spends its time in.  python tools/prof_step.py [vitb16|flickr8k] [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile

import atq
from atq.mixed_precision_atq import GradualQuantizationScheduler
from workloads import train as T

which = sys.argv[1] if len(sys.argv) > 1 else "vitb16"
atq.set_gemm_mode(sys.argv[2] if len(sys.argv) > 2 else "parity")
cfg = T.VITB16 if which == "vitb16" else T.FLICKR8K_SHAPE
dev = torch.device("cuda:0")
model, _, manager = T.build_retrieval(atq, cfg)
model.to(dev).train()
GradualQuantizationScheduler(model, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs).step(cfg.epoch)
opt = T.make_optimizer(model, cfg, fused=True)
batches = [tuple(t.to(dev) for t in b) for b in T.synthetic_batches(cfg, 2, seed=42)]
for i in range(3):
    T.retrieval_step(model, manager, opt, batches[i % 2], prepare=atq.prepare_quantization)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    T.retrieval_step(model, manager, opt, batches[0], prepare=atq.prepare_quantization)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.2f} ms over {sum(e.count for e in rows)} launches")
for e in rows[:45]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / tot:5.1f}% x{e.count:<5d} {e.key[:150]}")
