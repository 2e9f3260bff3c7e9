"""BASELINE config 5 (single GPU slice): quantize / pack / unpack throughput, per layer shape."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq._engine as eng
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
for (M, K) in ((4096, 4096), (8192, 8192), (16384, 8192)):
    n = M * K
    g = torch.Generator(device=dev).manual_seed(0)
    w = (torch.rand(M, K, device=dev, generator=g) * 2 - 1) / K ** 0.5
    thr = eng.adaptive_threshold(w, 0.3)
    packed = eng.ternarize_pack2(w, thr)
    t = eng.ternarize_f32(w, thr)
    rows = [("threshold (exact k-th |W|)", 4.0, lambda: eng.adaptive_threshold(w, 0.3)),
            ("ternarize -> fp32", 8.0, lambda: eng.ternarize_f32(w, thr)),
            ("ternarize -> 2-bit", 4.25, lambda: eng.ternarize_pack2(w, thr)),
            ("pack fp32 -> 2-bit", 4.25, lambda: eng.pack2_from_f32(t)),
            ("unpack 2-bit -> fp32", 4.25, lambda: eng.unpack2(packed, n)),
            ("unpack 2-bit -> bf16", 2.25, lambda: eng.unpack2(packed, n, torch.bfloat16)),
            ("quantize+pack layer (threshold + ternarize->2bit)", 4.25, lambda: eng.ternarize_pack2(w, eng.adaptive_threshold(w, 0.3)))]
    for name, bpe, fn in rows:
        ms = timeit(fn)
        gbs = bpe * n / ms / 1e6
        print(json.dumps({"layer": f"{M}x{K}", "kernel": name, "ms": round(ms, 4), "alg_bytes_per_elem": bpe,
                          "achieved_gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peaks["hbm_gbs"], 4),
                          "gelem_per_s": round(n / ms / 1e6, 2)}), flush=True)
    del w, t, packed
    torch.cuda.empty_cache()
