"""BASELINE config 3: TernaryLinear / RPB forward+backward sweep, 4096^2 and 8192^2 weights,
tokens 1k-64k on one B200.  Prints one JSON line per point (CUDA events, L2 flushed between reps)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq
import atq._engine as eng

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = [tuple(int(b) for b in a.split("x")) if "x" in a else (int(a), int(a)) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [(4096, 4096)]  # OUTxIN
tokens = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 8192, 65536]
reps = 5


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flushbuf.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["ternary", "rpb0.05", "rpb0.2"]
for OUT, IN in sizes:
    for kind in kinds:
        torch.manual_seed(0)
        if kind == "ternary":
            mod = atq.TernaryLinear(IN, OUT).to(dev)
        else:
            mod = atq.ResidualPrecisionBoostLinear(IN, OUT, float(kind[3:]), True, 0.3).to(dev)
        for N in tokens:
            x = torch.randn(N, IN, device=dev)
            gy = torch.randn(N, OUT, device=dev)
            for mode in ("parity", "fast"):
                atq.set_gemm_mode(mode)
                mod._ops.key = None
                with torch.no_grad():
                    mod(x[:8])  # quantize + build operands (cached afterwards)
                t_q = None
                xi = x.clone().requires_grad_(True)

                def fwd():
                    global y
                    y = mod(xi)

                def fwdbwd():
                    mod.zero_grad(set_to_none=True)
                    xi.grad = None
                    mod(xi).backward(gy)

                ms_f = timeit(lambda: torch.no_grad()(lambda: mod(x))())
                ms_fb = timeit(fwdbwd)
                n_gemm = 2 if kind == "ternary" else 3
                flops_f = 2.0 * N * OUT * IN
                rec = {"M": OUT, "K": IN, "tokens": N, "layer": kind, "mode": mode,
                       "fwd_ms": round(ms_f, 4), "fwd_tflops": round(flops_f / ms_f / 1e9, 1),
                       "fwdbwd_ms": round(ms_fb, 4), "fwdbwd_tflops": round(n_gemm * flops_f / ms_fb / 1e9, 1),
                       "fwdbwd_frac_of_bf16_peak": round(n_gemm * flops_f / ms_fb / 1e9 / peaks["bf16_tflops"], 4),
                       "note": "useful flops 2NKM per GEMM; includes operand split pre-passes, bias/alpha grads; weights cached"}
                print(json.dumps(rec), flush=True)
            del x, gy
        del mod
        torch.cuda.empty_cache()
