"""Device time per kernel (CUPTI activity records, warm, back-to-back) of ONE RPB layer forward + backward at the
config-2 shapes (800 tokens, 192/384 features), with CTA pairs on and off: where a launch-bound step spends its time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile

import atq
import atq._engine as eng

dev = torch.device("cuda:0")


def table(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    rows = [(e.key[:110], e.count / n, e.device_time_total / max(1, e.count)) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1] * r[2])
    tot = sum(r[1] * r[2] for r in rows)
    for k, c, us in rows:
        print(f"   {us:7.2f} us x{c:4.1f}  {k}")
    print(f"   total {tot:.1f} us per fwd+bwd")


for pairs in (1, 0):
    eng.set_cta_pairs(bool(pairs))
    for (tok, fin, fout) in ((800, 192, 192), (800, 192, 384), (16, 512, 192)):
        torch.manual_seed(0)
        mod = atq.ResidualPrecisionBoostLinear(fin, fout, 0.2, True, 0.3).to(dev)
        x = torch.randn(tok, fin, device=dev, requires_grad=True)
        gy = torch.randn(tok, fout, device=dev)

        def step():
            mod.zero_grad(set_to_none=True)
            x.grad = None
            mod(x).backward(gy)
        print(f"== RPB {fin}->{fout}, {tok} tokens, cta_pairs={pairs} (weights cached: no re-quantization)")
        table(step)
