"""Phase timing of one CTA of the attention kernels (needs the -DATQ_ATTN_PROF build: tools/build_attn_prof.sh, then
ATQ_SM100_LIB=.../libatq_sm100_prof.so).  Prints clock64 deltas between the stamps of CTA 200."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atq-multimodal_b200")):
    sys.path.insert(0, p)
import torch
import atq
import atq._native as nv
from atq.attention import attention_core

dev = torch.device("cuda:0")
B, H, L, D = int(os.environ.get("ATTN_B", "512")), 12, 197, 64
mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
atq.set_gemm_mode(mode)
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.randn(B, L, H * D, device=dev, generator=g).requires_grad_(True) for _ in range(3))
do = torch.randn(B, L, H * D, device=dev, generator=g)


def stamps():
    buf = (ctypes.c_longlong * 64)()
    assert nv.lib.atq_debug_attn_prof(buf) == 0
    return list(buf)


nv.lib.atq_debug_attn_prof.restype = ctypes.c_int
nv.lib.atq_debug_attn_prof.argtypes = [ctypes.c_void_p]
for it in range(3):
    out = attention_core(q, k, v, H, None, None, 0.1, True)
    torch.cuda.synchronize()
    f = stamps()
    out.backward(do)
    torch.cuda.synchronize()
    b = stamps()
names_f = ["start", "kv_staged"] + [f"t{t}:{n}" for t in range(2) for n in
                                    ("q_staged", "S_done", "max_done", "P_hi_stored", "PV_hi_done", "P_lo_stored", "PV_lo_done", "epilogue")]
print("forward (cycles since start, delta):")
for i, n in enumerate(names_f):
    print(f"  {n:16s} {f[i] - f[0]:8d} {f[i] - f[i - 1] if i else 0:8d}")
print("  fwd extra: setup", f[20] - f[0], "K", f[21] - f[20], "V", f[22] - f[21], "sync", f[1] - f[22], "warp arrivals", [f[24 + w] - f[0] for w in range(8)])
print("  loads landed (K, V) at", [f[32 + i] - f[0] for i in range(8)], "count", f[62])
print("backward:")
names_b = ["start"] + [f"k{i >> 1}q{i & 1}:{n}" for i in range(4) for n in
                       ("staged", "S_dP_done", "P_dS_stored", "mma_done", "dQ_stored", "P_lo_stored", "lo_done", "unused")]
prev = b[0]
for i, n in enumerate(names_b):
    if n.endswith("unused"):
        continue
    print(f"  {n:18s} {b[i] - b[0]:8d} {b[i] - prev:8d}")
    prev = b[i]
for kk in range(2):
    print(f"  k{kk}:dKdV_stored     {b[40 + kk] - b[0]:8d}")
