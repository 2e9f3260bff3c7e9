"""Fused attention core timed alone (CUDA events, L2 flushed): config-4 shapes.  Useful flops: forward
4*L^2*64 per (batch, head), backward 10*L^2*64 (S and dP recomputed products count once each)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch

import atq
from atq import attention as A

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


for (b, h, l, p) in [(512, 12, 197, 0.1), (512, 12, 50, 0.1), (512, 12, 197, 0.0)]:
    e = h * 64
    q, k, v = (torch.randn(b, l, e, device=dev, requires_grad=True) for _ in range(3))
    dout = torch.randn(b, l, e, device=dev)
    for mode in ("parity", "fast"):
        atq.set_gemm_mode(mode)
        seed = torch.tensor([5], dtype=torch.int64, device=dev)
        out = A.attention_core(q, k, v, h, None, None, p, True, seed=seed)
        ms_f = timeit(lambda: A.attention_core(q.detach(), k.detach(), v.detach(), h, None, None, p, True, seed=seed))

        def fb():
            q.grad = k.grad = v.grad = None
            A.attention_core(q, k, v, h, None, None, p, True, seed=seed).backward(dout)
        ms_fb = timeit(fb)
        ff, fbw = 4.0 * l * l * 64 * b * h, 10.0 * l * l * 64 * b * h
        print(json.dumps({"B": b, "H": h, "L": l, "dropout": p, "mode": mode, "fwd_ms": round(ms_f, 3),
                          "fwd_tflops": round(ff / ms_f / 1e9, 1), "bwd_ms": round(ms_fb - ms_f, 3),
                          "bwd_tflops": round(fbw / (ms_fb - ms_f) / 1e9, 1)}), flush=True)
