"""A/B of the fused (one cooperative launch) vs per-pass exact select at config-5 layer sizes; CUDA events, L2 flushed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq._engine as eng
import atq._native as nv
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flushbuf.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); tot += s.elapsed_time(e)
    return tot / reps
for (M, K) in ((768, 768), (3072, 768), (2048, 2048), (4096, 4096), (4096, 6144), (5120, 6144)):
    n = M * K
    g = torch.Generator(device=dev).manual_seed(0)
    w = (torch.rand(M, K, device=dev, generator=g) * 2 - 1) / K ** 0.5
    ref = torch.sort(w.abs().flatten()).values[int(0.3 * n)]
    out = {}
    for fused in (1, 0):
        nv.lib.atq_set_fused_select(fused)
        thr = eng.adaptive_threshold(w, 0.3)
        assert float(thr) == float(ref), (fused, float(thr), float(ref))
        ms = timeit(lambda: eng.adaptive_threshold(w, 0.3))
        out["fused" if fused else "per_pass"] = {"us": round(ms * 1e3, 1), "frac_of_measured_hbm": round(4.0 * n / ms / 1e6 / peaks["hbm_gbs"], 3)}
    print(json.dumps({"layer": f"{M}x{K}", **out}), flush=True)
nv.lib.atq_set_fused_select(1)
# batched: 16 layers of 2048^2 in one call
ws = [(torch.rand(2048, 2048, device=dev) * 2 - 1) / 45 for _ in range(16)]
for fused in (1, 0):
    nv.lib.atq_set_fused_select(fused)
    thr = eng.adaptive_threshold_batched(ws, [0.3] * 16)
    for t, w in zip(thr, ws):
        assert float(t) == float(torch.sort(w.abs().flatten()).values[int(0.3 * w.numel())])
    ms = timeit(lambda: eng.adaptive_threshold_batched(ws, [0.3] * 16))
    print(json.dumps({"batched 16 x 2048^2": "fused" if fused else "per_pass", "us": round(ms * 1e3, 1),
                      "frac_of_measured_hbm": round(4.0 * 16 * 2048 * 2048 / ms / 1e6 / peaks["hbm_gbs"], 3)}), flush=True)
nv.lib.atq_set_fused_select(1)
