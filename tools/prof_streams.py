"""Config-2 step (eager): kernel time per CUDA stream and the top kernels of each, from a chrome trace of torch.profiler
(which tower is the critical path of the two-branch step?)."""
import collections
import json
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

cfg = T.FLICKR8K_SHAPE
dev = torch.device("cuda:0")
model, _, manager = T.build_retrieval(atq, cfg)
model.to(dev).train()
model.image_encoder.base_model.to(memory_format=torch.channels_last)  # as bench.py does (caller-side layout of the fp32 trunk)
GradualQuantizationScheduler(model, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs).step(cfg.epoch)
opt = T.make_optimizer(model, cfg, fused=True)
batches = [tuple(t.to(dev) for t in b) for b in T.synthetic_batches(cfg, 2, seed=42)]
batches = [(b[0].contiguous(memory_format=torch.channels_last),) + b[1:] for b in batches]
for i in range(3):
    T.retrieval_step(model, manager, opt, batches[i % 2], prepare=atq.prepare_quantization)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    T.retrieval_step(model, manager, opt, batches[0], prepare=atq.prepare_quantization)
    torch.cuda.synchronize()
path = "/tmp/trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
per = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0]))
tot = collections.defaultdict(float)
span = {}
for e in ev:
    s = e["args"].get("stream")
    name = e["name"].split("<")[0].split("(")[0][-60:]
    per[s][name][0] += 1
    per[s][name][1] += e["dur"]
    tot[s] += e["dur"]
    b, en = span.get(s, (1e30, 0))
    span[s] = (min(b, e["ts"]), max(en, e["ts"] + e["dur"]))
for s in sorted(tot, key=lambda k: -tot[k]):
    n = sum(v[0] for v in per[s].values())
    print(f"stream {s}: {tot[s]:.0f} us of kernels in {n} launches, span {span[s][1] - span[s][0]:.0f} us")
    for name, (c, d) in sorted(per[s].items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"    {d:8.1f} us x{c:<4d} {name}")
