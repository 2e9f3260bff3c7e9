"""Config-2 step as a CUDA graph, captured several times in one process: per-capture replay time (the step time is
bimodal across runs) and, for the fastest / slowest capture, kernel time and span per stream of one replay."""
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
from atq.optim import FlatAdamW
from workloads import train as T

cfg = T.FLICKR8K_SHAPE
dev = torch.device("cuda:0")
model, _, manager = T.build_retrieval(atq, cfg)
model.to(dev).train()
model.image_encoder.base_model.to(memory_format=torch.channels_last)
GradualQuantizationScheduler(model, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs).step(cfg.epoch)
opt = T.make_optimizer(model, cfg, capturable=True, fused=True, adamw_cls=FlatAdamW)
batches = []
for b in T.synthetic_batches(cfg, 2, seed=42):
    b = tuple(t.to(dev) for t in b)
    batches.append((b[0].contiguous(memory_format=torch.channels_last),) + b[1:])
for i in range(5):
    T.retrieval_step(model, manager, opt, batches[i % 2], None, None, atq.prepare_quantization)
torch.cuda.synchronize()
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_graph(g, n=60):
    tot = 0.0
    for i in range(n):
        flushbuf.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g(batches[i % 2]); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def stream_profile(g):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        flushbuf.zero_(); g(batches[0]); torch.cuda.synchronize()
    prof.export_chrome_trace("/tmp/t.json")
    ev = [e for e in json.load(open("/tmp/t.json"))["traceEvents"] if e.get("cat") == "kernel" and "vectorized_elementwise_kernel<4, at::native::FillFunctor<unsigned char>" not in e["name"]]
    tot, span, cnt = collections.defaultdict(float), {}, collections.defaultdict(int)
    for e in ev:
        s = e["args"].get("stream")
        tot[s] += e["dur"]; cnt[s] += 1
        b, en = span.get(s, (1e30, 0))
        span[s] = (min(b, e["ts"]), max(en, e["ts"] + e["dur"]))
    t0 = min(v[0] for v in span.values())
    return {str(s): {"kernel_us": round(tot[s]), "launches": cnt[s], "from_us": round(span[s][0] - t0), "to_us": round(span[s][1] - t0)} for s in tot}


res = []
for c in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    g = T.GraphedRetrievalStep(model, manager, opt, batches[0], None, None, prepare=atq.prepare_quantization)
    ms = time_graph(g)
    prof = stream_profile(g)
    res.append((ms, prof))
    print(json.dumps({"capture": c, "ms_per_step": round(ms, 4), "streams": prof}), flush=True)
    g.release()
    del g
