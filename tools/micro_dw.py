"""Isolates where the small-shape dW GEMM spends its time: per-launch CUDA-event time (200 back-to-back
launches) of atq_tgemm_dw_masked with/without side inputs and of plain atq_tgemm in the three operand layouts."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch

import atq._engine as eng

dev = torch.device("cuda:0")


def per_launch(fn, n=20):
    """mean DEVICE time of the tgemm kernel itself (CUPTI activity records)"""
    from torch.profiler import ProfilerActivity, profile
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if "tgemm_kernel" in e.key]
    return sum(e.device_time_total for e in ev) / max(1, sum(e.count for e in ev))


for (m, k, n_tok) in [(192, 192, 16), (192, 192, 800), (768, 768, 800)]:
    g = torch.Generator(device=dev).manual_seed(0)
    dy = torch.randn(n_tok, m, device=dev, generator=g)
    x = torch.randn(n_tok, k, device=dev, generator=g)
    mask = (torch.rand(m, k, device=dev, generator=g) < 0.2).float()
    tq = torch.randint(-1, 2, (m, k), device=dev, generator=g).float()
    packed, _ = eng.pack2_from_f32(tq.reshape(-1))
    ga, xa = eng.split_bf16(dy, True), eng.split_bf16(x, True)
    gt, xt = eng.split_bf16_t(dy, True), eng.split_bf16_t(x, True)   # K-major copies [m, n_tok], [k, n_tok]
    w = eng.split_bf16(torch.randn(m, k, device=dev, generator=g), True)
    res = {
        "dw masked+tern (MM)": per_launch(lambda: eng.tgemm_dw_masked(ga + (1,), xa + (1,), m, k, n_tok, mask=mask, packed=packed)),
        "dw mask only (MM)": per_launch(lambda: eng.tgemm_dw_masked(ga + (1,), xa + (1,), m, k, n_tok, mask=mask)),
        "dw no side (MM)": per_launch(lambda: eng.tgemm_dw_masked(ga + (1,), xa + (1,), m, k, n_tok)),
        "dw masked+tern (KK)": per_launch(lambda: eng.tgemm_dw_masked(gt, xt, m, k, n_tok, mask=mask, packed=packed)),
        "tgemm linear (MM)": per_launch(lambda: eng.tgemm(ga + (1,), xa + (1,), m, k, n_tok)),
        "tgemm linear (KK)": per_launch(lambda: eng.tgemm(gt, xt, m, k, n_tok)),
        "tgemm fwd-like (KK) tokens x m": per_launch(lambda: eng.tgemm(xa, w, n_tok, m, k)),
    }
    print(f"== dW {m}x{k} over {n_tok} tokens (device us per tgemm kernel)")
    for kname, v in res.items():
        print(f"  {kname:34s} {v:8.2f}")
