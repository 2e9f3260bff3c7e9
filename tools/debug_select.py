import os, sys, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq._native as nv
dev = torch.device("cuda:0")
for (M, K) in ((4096, 4096), (8192, 8192)):
    n = M * K
    g = torch.Generator(device=dev).manual_seed(0)
    w = (torch.rand(M, K, device=dev, generator=g) * 2 - 1) / K ** 0.5
    k = int(0.3 * n)
    thr = torch.empty((), dtype=torch.float32, device=dev)
    ws = torch.zeros(nv.lib.atq_workspace_bytes_adaptive_threshold(n), dtype=torch.uint8, device=dev)
    for rep in range(2):
        nv.call("atq_adaptive_threshold", 0, w.data_ptr(), n, k, 0.05, thr.data_ptr(), ws.data_ptr(), ws.numel(), nv.stream_ptr(0))
    torch.cuda.synchronize()
    raw = bytes(ws[0:32].cpu().numpy())
    lo, hi, ncand, fb, below, cap = struct.unpack("<IIIIQQ", raw)
    import numpy as np
    print(M, K, "lo", np.uint32(lo).view(np.float32), "hi", np.uint32(hi).view(np.float32), "ncand", ncand, "fallback", fb,
          "below", below, "cap", cap, "k", k, "thr", float(thr))
