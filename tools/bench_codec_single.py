"""Per-layer codec entry points (ternarize -> 2-bit, pack, unpack) at several layer sizes: A/B of the routing through the
whole-model kernel (default) against the simple per-layer kernels (ATQ_CODEC_PER_LAYER_KERNELS=1).  Timing as in bench.py:
median of 9, L2 flushed, host launch latency hidden behind a device-side spin."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq._engine as eng
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flushbuf.zero_()
        torch.cuda._sleep(3_000_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


route = "per-layer kernels" if os.environ.get("ATQ_CODEC_PER_LAYER_KERNELS") == "1" else "via whole-model kernel"
for (M, K) in ((512, 1024), (2048, 1024), (4096, 4096), (8192, 8192)):
    n = M * K
    g = torch.Generator(device=dev).manual_seed(0)
    w = (torch.rand(M, K, device=dev, generator=g) * 2 - 1) / K ** 0.5
    thr = eng.adaptive_threshold(w, 0.3)
    packed = eng.ternarize_pack2(w, thr)
    t = eng.ternarize_f32(w, thr)
    for name, fn in (("ternarize -> 2-bit", lambda: eng.ternarize_pack2(w, thr)), ("pack fp32 -> 2-bit", lambda: eng.pack2_from_f32(t)),
                     ("unpack 2-bit -> fp32", lambda: eng.unpack2(packed, n))):
        ms = timeit(fn)
        print(json.dumps({"route": route, "layer": f"{M}x{K}", "kernel": name, "us": round(ms * 1e3, 2),
                          "frac_of_measured_hbm": round(4.25 * n / ms / 1e6 / peaks["hbm_gbs"], 4)}), flush=True)
