"""Host-side runtime of the ATQ hot path on B200: tensor-level wrappers over the C ABI,
the per-layer quantization cache, and the autograd nodes that reproduce the reference's
gradient contract:

  TernaryLinear : dX = dY (alpha T);  d(alpha) = sum(G .* T);  d(bias);  weight.grad stays None
  RPB           : dX = dY Wm;  dW = G .* mask;  d(alpha) = sum(G .* T .* (1-mask));  d(bias)

with G = dY^T X.  A true straight-through estimator (dW = G for TernaryLinear) is opt-in
(`atq.set_ste(True)`), off by default, and never used by the parity tests.

GEMM precision modes (every mode accumulates in fp32 in TMEM):
  "parity" (default)  every fp32 operand travels as a SCALED fp16 (hi, lo) pair: one power-of-two scale per tensor
                      (max|s x| in [2^14, 2^15), found by a grid-level |x| max reduction), hi = fp16(s x),
                      lo = fp16(s x - hi), 2-3 tcgen05.mma terms into one accumulator, 1/s applied in the epilogue.
                      ~22 significant bits per operand: the reference's own models keep fp32-level gradients on it.
  "parity_bf16"       the same terms on bf16 (hi, lo) pairs, no scale pass (~16 bits; meets rtol 1e-2 / atol 1e-3 on a
                      single layer, but a deep, badly conditioned network amplifies its 2^-17 to per-cent level)
  "fast"              one bf16 term.
"""
from __future__ import annotations

import collections
import ctypes
import os
from typing import Optional

import torch

from . import _native as nv

_MODE = os.environ.get("ATQ_GEMM_MODE", "parity")
_STE = os.environ.get("ATQ_STE", "0") == "1"
# B operand of the TernaryLinear GEMMs: "always" = 2-bit codec bytes expanded in shared memory,
# "never" = bf16 copy through TMA, "auto" = packed while the GEMM is weight-bandwidth bound
# (few tokens) and bf16-TMA once it is tensor-bound (the 1-CTA packed kernel is shared-memory
# bandwidth limited at ~60 % of the bf16 peak, the TMA kernel reaches ~90 %).
_PACKED = os.environ.get("ATQ_PACKED_GEMM", "auto")
_PACKED_AUTO_MAX_TOKENS = 256


def set_gemm_mode(mode: str) -> None:
    global _MODE
    if mode not in ("parity", "parity_bf16", "fast"):
        raise ValueError("mode must be 'parity', 'parity_bf16' or 'fast'")
    _MODE = mode


def get_gemm_mode() -> str:
    return _MODE


def set_ste(enabled: bool) -> None:
    global _STE
    _STE = bool(enabled)


def set_cta_pairs(enabled: bool) -> bool:
    """GEMMs with a lo operand part run on CTA pairs (cta_group::2) by default; False forces single-CTA kernels.
    Returns the previous setting."""
    return bool(nv.lib.atq_set_cta_pairs(1 if enabled else 0))


if os.environ.get("ATQ_CTA_PAIRS", "1") != "1":
    nv.lib.atq_set_cta_pairs(int(os.environ["ATQ_CTA_PAIRS"]))  # 0 = single CTA, 5 = pairs without the 256-wide tiles


def _use_lo() -> bool:
    return _MODE != "fast"


def _use_f16() -> bool:
    return _MODE == "parity"


def set_packed_gemm(policy: str) -> None:
    global _PACKED
    if policy not in ("auto", "always", "never"):
        raise ValueError("policy must be 'auto', 'always' or 'never'")
    _PACKED = policy


def _want_packed(tokens: int) -> bool:
    return _PACKED == "always" or (_PACKED == "auto" and tokens <= _PACKED_AUTO_MAX_TOKENS)


# ---------------------------------------------------------------------------------------
# tensor-level ops
# ---------------------------------------------------------------------------------------

def threshold_index(numel: int, sparsity_target) -> int:
    # atq/quantizers.py:28 -- Python double arithmetic, truncation toward zero
    return int(sparsity_target * numel)


def adaptive_threshold(w: torch.Tensor, sparsity_target, threshold_factor=0.05) -> torch.Tensor:
    """0-dim fp32 CUDA tensor holding the layer threshold (A1); no host sync."""
    w = nv.require_f32(w, "weights")
    dev = nv.device_index(w)
    n = w.numel()
    k = threshold_index(n, sparsity_target)
    thr = torch.empty((), dtype=torch.float32, device=w.device)
    ws = nv.workspace(nv.lib.atq_workspace_bytes_adaptive_threshold(n), w.device)
    nv.call("atq_adaptive_threshold", dev, w.data_ptr(), n, k, float(threshold_factor), thr.data_ptr(),
            ws.data_ptr(), ws.numel(), nv.stream_ptr(dev))
    return thr


def adaptive_threshold_batched(weights, sparsity_targets, threshold_factor=0.05):
    """Thresholds of many layers in one launch sequence.  Returns a [count] fp32 tensor."""
    count = len(weights)
    ws_list = [nv.require_f32(w, "weights") for w in weights]
    dev = nv.device_index(ws_list[0])
    thr = torch.empty(count, dtype=torch.float32, device=ws_list[0].device)
    n_arr = (ctypes.c_int64 * count)(*[w.numel() for w in ws_list])
    k_arr = (ctypes.c_int64 * count)(*[threshold_index(w.numel(), s) for w, s in zip(ws_list, sparsity_targets)])
    w_arr = (ctypes.c_void_p * count)(*[w.data_ptr() for w in ws_list])
    t_arr = (ctypes.c_void_p * count)(*[thr.data_ptr() + 4 * i for i in range(count)])
    wsb = nv.workspace(nv.lib.atq_workspace_bytes_adaptive_threshold_batched(count, n_arr), thr.device)
    nv.call("atq_adaptive_threshold_batched", dev, count, w_arr, n_arr, k_arr, float(threshold_factor), t_arr,
            wsb.data_ptr(), wsb.numel(), nv.stream_ptr(dev))
    return thr


def select_kth_abs(x: torch.Tensor, k: int) -> torch.Tensor:
    x = nv.require_f32(x, "input")
    dev = nv.device_index(x)
    n = x.numel()
    thr = torch.empty((), dtype=torch.float32, device=x.device)
    ws = nv.workspace(nv.lib.atq_workspace_bytes_select_kth_abs(n), x.device)
    nv.call("atq_select_kth_abs", dev, x.data_ptr(), n, int(k), thr.data_ptr(), ws.data_ptr(), ws.numel(),
            nv.stream_ptr(dev))
    return thr


def ternarize_f32(w: torch.Tensor, thr: torch.Tensor, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    w = nv.require_f32(w, "weights")
    dev = nv.device_index(w)
    t = torch.empty_like(w)
    nv.call("atq_ternarize_f32", dev, w.data_ptr(), w.numel(), thr.data_ptr(), t.data_ptr(), nv.ptr(stats),
            nv.stream_ptr(dev))
    return t


def ternarize_pack2(w: torch.Tensor, thr: torch.Tensor, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    w = nv.require_f32(w, "weights")
    dev = nv.device_index(w)
    n = w.numel()
    packed = torch.empty((n + 3) // 4, dtype=torch.uint8, device=w.device)
    nv.call("atq_ternarize_pack2", dev, w.data_ptr(), n, thr.data_ptr(), packed.data_ptr(), nv.ptr(stats),
            nv.stream_ptr(dev))
    return packed


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _n_array(tensors):
    return (ctypes.c_int64 * len(tensors))(*[t.numel() for t in tensors])


def ternarize_pack2_batched(weights, thresholds):
    """2-bit codec bytes of many layers in one launch (thresholds: [count] tensor or list of 0-dim tensors)."""
    ws = [nv.require_f32(w, "weights") for w in weights]
    dev = nv.device_index(ws[0])
    packed = [torch.empty((w.numel() + 3) // 4, dtype=torch.uint8, device=w.device) for w in ws]
    t_arr = (ctypes.c_void_p * len(ws))(*[thresholds[i].data_ptr() for i in range(len(ws))])
    nv.call("atq_ternarize_pack2_batched", dev, len(ws), _ptr_array(ws), _n_array(ws), t_arr, _ptr_array(packed), nv.stream_ptr(dev))
    return packed


def pack2_from_f32_batched(tensors):
    ts = [nv.require_f32(t, "ternary_weights") for t in tensors]
    dev = nv.device_index(ts[0])
    packed = [torch.empty((t.numel() + 3) // 4, dtype=torch.uint8, device=t.device) for t in ts]
    flag = torch.zeros(1, dtype=torch.int32, device=ts[0].device)
    nv.call("atq_pack2_from_f32_batched", dev, len(ts), _ptr_array(ts), _n_array(ts), _ptr_array(packed), flag.data_ptr(), nv.stream_ptr(dev))
    return packed, flag


def unpack2_batched(packed_list, numels):
    dev = nv.device_index(packed_list[0])
    outs = [torch.empty(n, dtype=torch.float32, device=p.device) for p, n in zip(packed_list, numels)]
    flag = torch.zeros(1, dtype=torch.int32, device=packed_list[0].device)
    n_arr = (ctypes.c_int64 * len(outs))(*[int(n) for n in numels])
    nv.call("atq_unpack2_to_f32_batched", dev, len(outs), _ptr_array(packed_list), n_arr, _ptr_array(outs), flag.data_ptr(), nv.stream_ptr(dev))
    return outs, flag


def optimal_alpha(w: torch.Tensor, tern_stats: torch.Tensor) -> torch.Tensor:
    """alpha* of atq/quantizers.py:46-55 from the ternarize statistics (device-resolved branch)."""
    dev = nv.device_index(w)
    abs_stats = torch.empty(16, dtype=torch.uint8, device=w.device)
    nv.call("atq_abs_stats", dev, w.data_ptr(), w.numel(), abs_stats.data_ptr(), None, 0, nv.stream_ptr(dev))
    alpha = torch.empty((), dtype=torch.float32, device=w.device)
    nv.call("atq_optimal_alpha", dev, tern_stats.data_ptr(), abs_stats.data_ptr(), w.numel(), alpha.data_ptr(),
            nv.stream_ptr(dev))
    return alpha


def pack2_from_f32(t: torch.Tensor):
    t = nv.require_f32(t, "ternary_weights")
    dev = nv.device_index(t)
    n = t.numel()
    packed = torch.empty((n + 3) // 4, dtype=torch.uint8, device=t.device)
    flag = torch.zeros(1, dtype=torch.int32, device=t.device)
    if n > 0:
        nv.call("atq_pack2_from_f32", dev, t.data_ptr(), n, packed.data_ptr(), flag.data_ptr(), nv.stream_ptr(dev))
    return packed, flag


def unpack2(packed: torch.Tensor, n: int, dtype=torch.float32, flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    if packed.dtype != torch.uint8:
        raise RuntimeError("atq: packed_weights must be uint8")
    dev = nv.device_index(packed)
    packed = packed.contiguous()
    if packed.numel() < (n + 3) // 4:
        raise RuntimeError("atq: packed_weights too short for num_values")
    out = torch.empty(n, dtype=dtype, device=packed.device)
    if n == 0:
        return out
    if dtype == torch.float32:
        nv.call("atq_unpack2_to_f32", dev, packed.data_ptr(), n, out.data_ptr(), nv.ptr(flag), nv.stream_ptr(dev))
    elif dtype == torch.bfloat16:
        nv.call("atq_unpack2_to_bf16", dev, packed.data_ptr(), n, out.data_ptr(), nv.stream_ptr(dev))
    elif dtype == torch.int8:
        nv.call("atq_unpack2_to_i8", dev, packed.data_ptr(), n, out.data_ptr(), nv.stream_ptr(dev))
    else:
        raise RuntimeError(f"atq: unsupported unpack dtype {dtype}")
    return out


def route_mask_mul(x: torch.Tensor, grad_out: torch.Tensor, thr: torch.Tensor) -> torch.Tensor:
    x = nv.require_f32(x, "input")
    g = nv.require_f32(grad_out, "grad_output")
    dev = nv.device_index(x)
    out = torch.empty_like(g)
    if x.numel():
        nv.call("atq_route_mask_mul", dev, x.data_ptr(), g.data_ptr(), thr.data_ptr(), x.numel(), out.data_ptr(),
                nv.stream_ptr(dev))
    return out


def absmax_slot(x2: torch.Tensor, bound_mul: float = 1.0, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Scale slot of a scaled-fp16 operand: one grid-level max|x| reduction; slot[1] = s, slot[2] = 1/s with
    max(max|x|, |extra|) * bound_mul * s in [2^14, 2^15).  No host synchronisation."""
    rows, cols = x2.shape
    dev = nv.device_index(x2)
    slot = nv.new_slot(x2.device)
    nv.call("atq_absmax_scale", dev, x2.data_ptr(), rows, cols, x2.stride(0), float(bound_mul), nv.ptr(extra),
            slot.data_ptr(), nv.stream_ptr(dev))
    return slot


def absmax_slots_batched(tensors, extras=None, bound_mul: float = 1.0):
    """Scale slots of many contiguous tensors in one launch (extras: per-tensor device scalars joining the maximum)."""
    count = len(tensors)
    dev = nv.device_index(tensors[0])
    slots = [nv.new_slot(tensors[0].device) for _ in range(count)]
    x_arr = (ctypes.c_void_p * count)(*[t.data_ptr() for t in tensors])
    n_arr = (ctypes.c_int64 * count)(*[t.numel() for t in tensors])
    e_arr = (ctypes.c_void_p * count)(*[None if e is None else e.data_ptr() for e in (extras or [None] * count)])
    s_arr = (ctypes.c_void_p * count)(*[sl.data_ptr() for sl in slots])
    nv.call("atq_absmax_scale_batched", dev, count, x_arr, n_arr, e_arr, s_arr, float(bound_mul), nv.stream_ptr(dev))
    return slots


def split_bf16(x2: torch.Tensor, want_lo: bool, slot: Optional[torch.Tensor] = None):
    """fp32 [rows, cols] -> (hi, lo|None, pitch) row-major: bf16 pair, or (when a scale slot is given) the scaled
    fp16 pair as (hi, lo, pitch, 0, slot)."""
    rows, cols = x2.shape
    dev = nv.device_index(x2)
    pitch = nv.round_up(cols, 8)
    dt = torch.bfloat16 if slot is None else torch.float16
    hi = torch.empty((rows, pitch), dtype=dt, device=x2.device)
    lo = torch.empty((rows, pitch), dtype=dt, device=x2.device) if want_lo else None
    nv.call("atq_split_bf16", dev, x2.data_ptr(), rows, cols, x2.stride(0), hi.data_ptr(), nv.ptr(lo), pitch,
            nv.ptr(slot), nv.stream_ptr(dev))
    return (hi, lo, pitch) if slot is None else (hi, lo, pitch, 0, slot)


_FUSED_MAX = int(nv.lib.atq_split_scaled_fused_max_elems())


# Producers that already reduced max|y| while writing y (layer_norm) leave the finished scale slot here, keyed by the
# output's memory; the operand split of that tensor then skips its own reduction pass.  A detached alias pins the memory
# (so the address cannot be recycled under the key); the version counter catches in-place edits.
_ABSMAX_HINTS: "collections.OrderedDict" = collections.OrderedDict()
_ABSMAX_HINT_STATS = {"set": 0, "hit": 0}


def _note_absmax(y: torch.Tensor, slot: torch.Tensor) -> None:
    _ABSMAX_HINTS[(nv.device_index(y), y.data_ptr())] = (y.numel(), y._version, y.detach(), slot)
    _ABSMAX_HINT_STATS["set"] += 1
    while len(_ABSMAX_HINTS) > 4:
        _ABSMAX_HINTS.popitem(last=False)


def _known_absmax(x2: torch.Tensor) -> Optional[torch.Tensor]:
    hit = _ABSMAX_HINTS.get((nv.device_index(x2), x2.data_ptr()))
    if hit is None or hit[0] != x2.numel() or hit[1] != x2._version or not x2.is_contiguous():
        return None
    _ABSMAX_HINT_STATS["hit"] += 1
    return hit[3]


def split_scaled(x2: torch.Tensor, want_lo: bool = True):
    """Scaled-fp16 operand of a contiguous fp32 [rows, cols] tensor: one cluster kernel (max|x|, scale, split) for
    small tensors, the grid-level reduction followed by the streaming split otherwise."""
    rows, cols = x2.shape
    n = rows * cols
    known = _known_absmax(x2) if _ABSMAX_HINTS else None
    if known is not None:
        return split_bf16(x2, want_lo, known)
    if _FUSED_SPLIT and n <= _FUSED_MAX and cols % 8 == 0 and x2.stride(0) == cols and x2.data_ptr() % 16 == 0:
        dev = nv.device_index(x2)
        hi = torch.empty((rows, cols), dtype=torch.float16, device=x2.device)
        lo = torch.empty((rows, cols), dtype=torch.float16, device=x2.device) if want_lo else None
        slot = nv.new_slot(x2.device)
        nv.call("atq_split_scaled_fused", dev, x2.data_ptr(), n, hi.data_ptr(), nv.ptr(lo), 1.0, None, slot.data_ptr(),
                nv.stream_ptr(dev))
        return (hi, lo, cols, 0, slot)
    return split_bf16(x2, want_lo, absmax_slot(x2))


# q_proj / k_proj / v_proj of a self-attention block receive the SAME tensor (models/text_encoder.py:82-84): the
# last operand built on each stream is kept (with a strong reference to its source, so the memory cannot be recycled
# under it) and reused while the source's (storage, version, layout, mode) still match.
_LAST_SPLIT: dict = {}
# (skipped while the stream is capturing: a captured step re-runs the kernels anyway and the saving is two small launches)
_SPLIT_REUSE = os.environ.get("ATQ_SPLIT_REUSE", "1") == "1"
_FUSED_SPLIT = os.environ.get("ATQ_FUSED_SPLIT", "1") == "1"


def split_operand(x2: torch.Tensor, owner: Optional[torch.Tensor] = None):
    """The A operand of a GEMM in the current precision mode (see the module docstring)."""
    if owner is not None and _SPLIT_REUSE and not torch.cuda.is_current_stream_capturing():
        skey = (nv.device_index(x2), nv.stream_ptr(nv.device_index(x2)))
        key = (x2.data_ptr(), owner._version, tuple(x2.shape), tuple(x2.stride()), _MODE)
        hit = _LAST_SPLIT.get(skey)
        if hit is not None and hit[0] == key:
            return hit[2]
        op = split_operand(x2)
        # pin the source MEMORY (so the key cannot match a recycled allocation) through a detached alias: holding the
        # tensor itself would keep its autograd graph -- and the AccumulateGrad nodes of every upstream parameter, bound
        # to whatever stream that step ran on -- alive into the next step (this is what broke whole-step graph capture)
        _LAST_SPLIT[skey] = (key, x2.detach(), op)
        return op
    if _use_f16():
        return split_scaled(x2)
    hi, lo, pitch = split_bf16(x2, _use_lo())
    return (hi, lo, pitch, 0, None)


def mn_view(op):
    """The same operand memory consumed MN-major ([k, rows] row-major view of a row-major tensor)."""
    return (op[0], op[1], op[2], 1, op[4] if len(op) > 4 else None)


def split_bf16_colsum(x2: torch.Tensor, want_lo: bool, slot: Optional[torch.Tensor] = None):
    """split_bf16 of a contiguous [rows, cols] tensor fused with its column sums (bias gradient).
    Falls back to the two separate kernels when the fused path's layout conditions do not hold."""
    rows, cols = x2.shape
    if cols % 8 != 0 or x2.stride(0) != cols or x2.data_ptr() % 16 != 0:
        op = split_bf16(x2, want_lo, slot)
        return (op + (0, None) if len(op) == 3 else op), colsum(x2)
    dev = nv.device_index(x2)
    dt = torch.bfloat16 if slot is None else torch.float16
    hi = torch.empty((rows, cols), dtype=dt, device=x2.device)
    lo = torch.empty((rows, cols), dtype=dt, device=x2.device) if want_lo else None
    out = torch.empty(cols, dtype=torch.float32, device=x2.device)
    ws = nv.workspace(nv.lib.atq_workspace_bytes_split_colsum(rows, cols), x2.device)
    nv.call("atq_split_bf16_colsum", dev, x2.data_ptr(), rows, cols, hi.data_ptr(), nv.ptr(lo), out.data_ptr(),
            ws.data_ptr(), ws.numel(), nv.ptr(slot), nv.stream_ptr(dev))
    return (hi, lo, cols, 0, slot), out


def split_operand_colsum(x2: torch.Tensor, want_lo: bool, f16: bool):
    """Operand split + column sums with the format of the OTHER operand of the GEMMs it feeds (saved from forward)."""
    if f16 and _FUSED_SPLIT and x2.numel() <= _FUSED_MAX and x2.shape[1] % 8 == 0 and x2.stride(0) == x2.shape[1] and x2.data_ptr() % 16 == 0:
        return split_scaled(x2, want_lo), colsum(x2)  # two launches instead of three
    return split_bf16_colsum(x2, want_lo, absmax_slot(x2) if f16 else None)


def split_bf16_t(x2: torch.Tensor, want_lo: bool):
    """fp32 [rows, cols] -> transposed (hi_t, lo_t|None, pitch_t), each [cols, pitch_t]."""
    rows, cols = x2.shape
    dev = nv.device_index(x2)
    pitch_t = nv.round_up(rows, 8)
    hi = torch.empty((cols, pitch_t), dtype=torch.bfloat16, device=x2.device)
    lo = torch.empty((cols, pitch_t), dtype=torch.bfloat16, device=x2.device) if want_lo else None
    nv.call("atq_split_bf16_t", dev, x2.data_ptr(), rows, cols, x2.stride(0), hi.data_ptr(), nv.ptr(lo), pitch_t,
            None, None, nv.stream_ptr(dev))
    return hi, lo, pitch_t


def colsum(x2: torch.Tensor) -> torch.Tensor:
    rows, cols = x2.shape
    dev = nv.device_index(x2)
    out = torch.empty(cols, dtype=torch.float32, device=x2.device)
    ws = nv.workspace(nv.lib.atq_workspace_bytes_colsum(rows, cols), x2.device)
    nv.call("atq_colsum_f32", dev, x2.data_ptr(), rows, cols, x2.stride(0), out.data_ptr(), ws.data_ptr(), ws.numel(),
            nv.stream_ptr(dev))
    return out


def tgemm(a, b, rows: int, cols: int, kdim: int, scale=None, bias=None, dot_ref=None):
    """out[rows, cols] = scale * (A . B^T) + bias;  a, b = (hi, lo|None, pitch).
    With dot_ref (fp32 [rows, cols]) also returns sum(acc .* dot_ref) as a [1] tensor."""
    hi = a[0]
    dev = nv.device_index(hi)
    out = torch.empty((rows, cols), dtype=torch.float32, device=hi.device)
    oa, ob = nv.operand(*a), nv.operand(*b)
    dot_out = torch.empty(1, dtype=torch.float32, device=hi.device) if dot_ref is not None else None
    ws = nv.workspace(nv.lib.atq_workspace_bytes_tgemm(rows, cols) if dot_ref is not None else 0, hi.device)
    nv.call("atq_tgemm", dev, rows, cols, kdim, ctypes.byref(oa), ctypes.byref(ob), nv.ptr(scale), nv.ptr(bias),
            out.data_ptr(), cols, nv.ptr(dot_ref), 0 if dot_ref is None else dot_ref.stride(0), nv.ptr(dot_out),
            ws.data_ptr(), ws.numel(), nv.stream_ptr(dev))
    return out, dot_out


def tgemm_absmax(a, b, rows: int, cols: int, kdim: int, bound_mul: float, scale=None, bias=None):
    """tgemm whose epilogue also reduces max|out|: returns (out, slot) with slot the scale slot of `out` as a scaled
    fp16 operand for bound = max|out| * bound_mul (no separate reduction pass over the output)."""
    hi = a[0]
    dev = nv.device_index(hi)
    out = torch.empty((rows, cols), dtype=torch.float32, device=hi.device)
    slot = nv.new_slot(hi.device)
    oa, ob = nv.operand(*a), nv.operand(*b)
    nv.call("atq_tgemm_absmax", dev, rows, cols, kdim, ctypes.byref(oa), ctypes.byref(ob), nv.ptr(scale), nv.ptr(bias),
            out.data_ptr(), cols, slot.data_ptr(), float(bound_mul), nv.stream_ptr(dev))
    return out, slot


def tgemm_packed(a, packed_b: torch.Tensor, rows: int, cols: int, kdim: int, scale=None, bias=None, dot_ref=None):
    """out[rows, cols] = scale * (A . T^T) + bias with T given as 2-bit codec bytes [cols, kdim/4]
    (kdim % 64 == 0): the packed weights are expanded to bf16 tiles in shared memory by the GEMM."""
    hi = a[0]
    dev = nv.device_index(hi)
    out = torch.empty((rows, cols), dtype=torch.float32, device=hi.device)
    oa = nv.operand(*a)
    dot_out = torch.empty(1, dtype=torch.float32, device=hi.device) if dot_ref is not None else None
    ws = nv.workspace(nv.lib.atq_workspace_bytes_tgemm(rows, cols) if dot_ref is not None else 0, hi.device)
    nv.call("atq_tgemm_packed", dev, rows, cols, kdim, ctypes.byref(oa), packed_b.data_ptr(), nv.ptr(scale), nv.ptr(bias),
            out.data_ptr(), cols, nv.ptr(dot_ref), 0 if dot_ref is None else dot_ref.stride(0), nv.ptr(dot_out),
            ws.data_ptr(), ws.numel(), nv.stream_ptr(dev))
    return out, dot_out


def packed_gemm_ok(kdim: int, packed: torch.Tensor) -> bool:
    return kdim % 64 == 0 and packed is not None and packed.data_ptr() % 16 == 0


def tgemm_dw_masked(dy_t, x_t, m_out: int, k_in: int, n_tok: int, mask=None, packed=None):
    hi = dy_t[0]
    dev = nv.device_index(hi)
    dw = torch.empty((m_out, k_in), dtype=torch.float32, device=hi.device)
    oa, ob = nv.operand(*dy_t), nv.operand(*x_t)
    dalpha = torch.empty(1, dtype=torch.float32, device=hi.device) if packed is not None else None
    ws = nv.workspace(nv.lib.atq_workspace_bytes_tgemm_dw(m_out, k_in, n_tok), hi.device)
    nv.call("atq_tgemm_dw_masked", dev, m_out, k_in, n_tok, ctypes.byref(oa), ctypes.byref(ob), nv.ptr(mask),
            nv.ptr(packed), dw.data_ptr(), k_in, nv.ptr(dalpha), ws.data_ptr(), ws.numel(), nv.stream_ptr(dev))
    return dw, dalpha


# ---------------------------------------------------------------------------------------
# per-layer quantization cache: everything the GEMMs consume, rebuilt only when
# the weight / alpha / mask tensors or the sparsity target changed.
# ---------------------------------------------------------------------------------------

class LayerOperands:
    __slots__ = ("key", "thr", "packed", "packed_t", "w", "w_t", "saw_grad", "hook_param")

    def __init__(self):
        self.saw_grad = False    # a gradient has reached the weight at least once (TernaryLinear: normally never)
        self.hook_param = None   # the Parameter object the gradient hook is registered on
        self.key = None
        self.thr = None
        self.packed = None    # 2-bit codec bytes of T, public layout [M*K/4]
        self.packed_t = None  # codec bytes of T^T [K, M/4] (TernaryLinear, M % 64 == 0)
        self.w = None      # (hi, lo|None, pitch)   [M, pitch]  forward B operand (None when packed is used)
        self.w_t = None    # (hi, lo|None, pitch_t) [K, pitch_t] dX B operand (None when packed_t is used)


# Fused / foreach optimizer kernels (torch.optim.AdamW(fused=True), capturable paths) update parameters
# without bumping Tensor._version, so the version alone cannot prove that cached operands are current.
# A global post-step hook counts optimizer steps; any step anywhere invalidates every layer's cache.
_OPT_STEPS = 0


def _count_optimizer_step(optimizer, args, kwargs):
    global _OPT_STEPS
    _OPT_STEPS += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook
    _register_post_hook(_count_optimizer_step)
except ImportError:  # pragma: no cover - older torch: fall back to versions only
    pass


def notify_weights_changed() -> None:
    """Call after mutating parameters behind autograd's back (e.g. through `.data`)."""
    global _OPT_STEPS
    _OPT_STEPS += 1


def _key(weight, alpha, mask, sparsity_target, threshold_factor, frozen=False):
    # frozen: a TernaryLinear weight that has never received a gradient (grad None => every optimizer skips
    # it, weight decay included) cannot be changed by an optimizer step, so the step counter is left out of its key
    return (0 if frozen else _OPT_STEPS, weight.data_ptr(), weight._version, tuple(weight.shape),
            None if alpha is None else (alpha.data_ptr(), alpha._version),
            None if mask is None else (mask.data_ptr(), mask._version),
            float(sparsity_target), float(threshold_factor), _MODE)


@torch.no_grad()
def layer_operands(cache: LayerOperands, weight, alpha, mask, sparsity_target, threshold_factor=0.05,
                   thr: Optional[torch.Tensor] = None, slot: Optional[torch.Tensor] = None) -> LayerOperands:
    """mask None -> TernaryLinear operands (T exact in bf16; alpha applied in the GEMM epilogue).
    mask given -> RPB mixed weight Wm = T*alpha*(1-mask) + W*mask as bf16 hi/lo."""
    frozen = False
    if mask is None and isinstance(weight, torch.nn.Parameter):
        if cache.hook_param is not weight:  # (re-)arm the detector on this Parameter object
            cache.hook_param, cache.saw_grad = weight, False
            if weight.requires_grad:
                # (autograd calls the hook with None when the only consumer returns no gradient: that does not count)
                weight.register_hook(lambda g, c=cache: setattr(c, "saw_grad", True) if g is not None else None)
        frozen = not cache.saw_grad and not _STE
    key = _key(weight, alpha if mask is not None else None, mask, sparsity_target, threshold_factor, frozen)
    if cache.key == key:
        return cache
    w = nv.require_f32(weight.detach(), "weight")
    dev = nv.device_index(w)
    M, K = w.shape
    if thr is None:
        thr = adaptive_threshold(w, sparsity_target, threshold_factor)
    pitch, pitch_t = nv.round_up(K, 8), nv.round_up(M, 8)
    n = M * K
    packed = torch.empty((n + 3) // 4, dtype=torch.uint8, device=w.device)
    flat_ok = (K % 4 == 0)
    st = nv.stream_ptr(dev)
    packed_t = None
    f16 = _use_f16()
    bf = torch.float16 if f16 else torch.bfloat16
    if mask is None:
        slot = None
        # TernaryLinear: the GEMMs read the 2-bit codec bytes directly whenever the contraction
        # dimension allows 16-byte codec rows per k-block; 16-bit copies only for odd shapes.  T is exact in
        # bf16 and in fp16 (no scale): the copy just has to carry the element format of the activations.
        hi = torch.empty((M, pitch), dtype=bf, device=w.device)
        hi_t = None  # dX reads `hi` in place through MN-major descriptors
        if M % 64 == 0:
            packed_t = torch.empty(n // 4, dtype=torch.uint8, device=w.device)
        lo = lo_t = None
        nv.call("atq_build_ternary_operands", dev, w.data_ptr(), M, K, thr.data_ptr(),
                packed.data_ptr() if flat_ok else None, nv.ptr(packed_t), nv.ptr(hi), pitch, nv.ptr(hi_t), pitch_t, None,
                1 if f16 else 0, st)
    else:
        want_lo = _use_lo()
        hi = torch.empty((M, pitch), dtype=bf, device=w.device)
        lo = torch.empty((M, pitch), dtype=bf, device=w.device) if want_lo else None
        hi_t = lo_t = None  # dX reads hi/lo in place through MN-major descriptors
        mk = nv.require_f32(mask, "precision_mask")
        al = nv.require_f32(alpha.detach(), "alpha")
        if not f16:
            slot = None
        elif slot is None:  # |Wm| <= max(max|W|, |alpha|): one reduction over W gives the scale of the mixed operand
            slot = absmax_slot(w, 1.0, al)
        nv.call("atq_build_mixed_operands", dev, w.data_ptr(), mk.data_ptr(), M, K, thr.data_ptr(), al.data_ptr(),
                packed.data_ptr() if flat_ok else None, hi.data_ptr(), nv.ptr(lo), pitch, None, None, pitch_t,
                nv.ptr(slot), st)
    if not flat_ok:  # rows of the flat codec do not start on byte boundaries
        nv.call("atq_ternarize_pack2", dev, w.data_ptr(), n, thr.data_ptr(), packed.data_ptr(), None, st)
    cache.key, cache.thr, cache.packed, cache.packed_t = key, thr, packed, packed_t
    cache.w = (hi, lo, pitch, 0, slot)     # [M, pitch]: forward B operand (K-major) ...
    cache.w_t = (hi, lo, pitch, 1, slot)   # ... and, read as MN-major, the dX B operand (no transposed copy)
    return cache


@torch.no_grad()
def prepare_quantization(model, threshold_factor=0.05) -> int:
    """Re-quantize every ternary layer of `model` whose weights / alpha / mask / sparsity target
    changed since its operands were built, with ONE batched threshold launch sequence for all of
    them (3 radix passes over all layers instead of 3 per layer).  Optional: the layers quantize
    lazily on their own if this is never called.  Returns the number of layers rebuilt."""
    stale = []
    for m in model.modules():
        cache = getattr(m, "_ops", None)
        if not isinstance(cache, LayerOperands) or not m.weight.is_cuda:
            continue
        mask = getattr(m, "precision_mask", None)
        s = getattr(m, "sparsity_target", 0.3)
        alpha = m.alpha if mask is not None else None
        frozen = mask is None and cache.hook_param is m.weight and not cache.saw_grad and not _STE
        if cache.key != _key(m.weight, alpha, mask, s, threshold_factor, frozen):
            stale.append((m, cache, alpha, mask, s))
    if not stale:
        return 0
    thr = adaptive_threshold_batched([m.weight.detach() for m, *_ in stale], [s for *_, s in stale], threshold_factor)
    slots = [None] * len(stale)
    rpb = [i for i, (m, cache, alpha, mask, s) in enumerate(stale) if mask is not None and m.weight.is_contiguous()]
    if _use_f16() and rpb:  # per-layer scales of the mixed weights: one launch for all layers
        got = absmax_slots_batched([stale[i][0].weight.detach() for i in rpb], [stale[i][0].alpha.detach() for i in rpb])
        for i, sl in zip(rpb, got):
            slots[i] = sl
    for i, (m, cache, alpha, mask, s) in enumerate(stale):
        layer_operands(cache, m.weight, m.alpha, mask, s, threshold_factor, thr=thr[i], slot=slots[i])
    return len(stale)


# ---------------------------------------------------------------------------------------
# autograd nodes
# ---------------------------------------------------------------------------------------

class _TernaryLinearFn(torch.autograd.Function):
    """y = x (alpha T)^T + b   (atq/layers.py:35-43)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, alpha, bias, ops: LayerOperands):
        M, K = weight.shape
        x2 = nv.require_f32(x, "input").reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        N = x2.shape[0]
        al = alpha.detach()
        if N == 0:
            y = x2.new_zeros((0, M))
        else:
            xa = split_operand(x2, x)
            b_ = None if bias is None else bias.detach()
            if _want_packed(N) and packed_gemm_ok(K, ops.packed):
                # packed 2-bit weights, expanded to bf16 tiles inside the GEMM
                y, _ = tgemm_packed(xa, ops.packed, N, M, K, scale=al, bias=b_)
            else:
                y, _ = tgemm(xa, ops.w, N, M, K, scale=al, bias=b_)
        ctx.save_for_backward(x2, al)
        ctx.ops_w_t, ctx.packed_t = ops.w_t, ops.packed_t
        ctx.has_bias = bias is not None
        ctx.wshape = (M, K)
        ctx.xshape = x.shape
        ctx.ste = _STE
        ctx.fmt = (_use_lo(), _use_f16())  # backward uses the element format the cached weight operand was built in
        return y.reshape(*x.shape[:-1], M)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        x2, al = ctx.saved_tensors
        M, K = ctx.wshape
        N = x2.shape[0]
        g2 = nv.require_f32(gy, "grad_output").reshape(-1, M)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        if N == 0:
            return (gy.new_zeros(ctx.xshape), None, al.new_zeros(1), g2.new_zeros(M) if ctx.has_bias else None, None)
        want_lo, f16 = ctx.fmt
        if ctx.has_bias:
            ga, dbias = split_operand_colsum(g2, want_lo, f16)  # one pass over dY: operand split + bias gradient
        else:
            ga, dbias = (split_scaled(g2, want_lo) if f16 else split_bf16(g2, want_lo)), None
        # dX = alpha * (dY . T);  d(alpha) = sum((dY . T) .* X) fused in the same epilogue
        if _want_packed(N) and packed_gemm_ok(M, ctx.packed_t):
            dx, dalpha = tgemm_packed(ga, ctx.packed_t, N, K, M, scale=al, dot_ref=x2)
        else:
            dx, dalpha = tgemm(ga, ctx.ops_w_t, N, K, M, scale=al, dot_ref=x2)
        dw = None
        if ctx.ste:  # opt-in straight-through estimator: dW = G (dY and X consumed MN-major, no transposes)
            xa = split_scaled(x2, want_lo) if f16 else split_bf16(x2, want_lo)
            dw, _ = tgemm_dw_masked(mn_view(ga), mn_view(xa), M, K, N)
        return dx.reshape(ctx.xshape), dw, dalpha, dbias, None


class _RPBLinearFn(torch.autograd.Function):
    """y = x Wm^T + b,  Wm = T*alpha*(1-mask) + W*mask   (atq/precision_boost.py:62-74)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, alpha, bias, mask, ops: LayerOperands):
        M, K = weight.shape
        x2 = nv.require_f32(x, "input").reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        N = x2.shape[0]
        if N == 0:
            y = x2.new_zeros((0, M))
        else:
            xa = split_operand(x2, x)
            y, _ = tgemm(xa, ops.w, N, M, K, scale=None, bias=None if bias is None else bias.detach())
        # backward consumes the SAME (hi, lo) split of x (MN-major, as the dW B operand): save it
        # instead of the fp32 activations (same bytes), so nothing is split or transposed twice
        if N == 0:
            ctx.save_for_backward(x2, mask)
        elif xa[1] is None:
            ctx.save_for_backward(xa[0], mask)
        else:
            ctx.save_for_backward(xa[0], xa[1], mask)
        ctx.x_pitch = xa[2] if N else 0
        ctx.x_slot = xa[4] if N else None
        ctx.n_tokens = N
        ctx.ops_w_t, ctx.packed = ops.w_t, ops.packed
        ctx.has_bias = bias is not None
        ctx.wshape = (M, K)
        ctx.xshape = x.shape
        return y.reshape(*x.shape[:-1], M)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        mask = saved[-1]
        M, K = ctx.wshape
        N = ctx.n_tokens
        g2 = nv.require_f32(gy, "grad_output").reshape(-1, M)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        if N == 0:
            z = g2.new_zeros
            return (gy.new_zeros(ctx.xshape), z((M, K)), z(1), z(M) if ctx.has_bias else None, None, None)
        xa = (saved[0], saved[1] if len(saved) == 3 else None, ctx.x_pitch, 0, ctx.x_slot)
        f16 = xa[0].dtype == torch.float16
        if ctx.has_bias:  # ONE pass over dY: the split feeds both backward GEMMs, the column sums are d(bias)
            ga, dbias = split_operand_colsum(g2, xa[1] is not None, f16)
        else:
            ga, dbias = (split_scaled(g2, xa[1] is not None) if f16 else split_bf16(g2, xa[1] is not None)), None
            if len(ga) == 3:
                ga = ga + (0, None)
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _ = tgemm(ga, ctx.ops_w_t, N, K, M)   # B = Wm [M, K] read MN-major
            dx = dx.reshape(ctx.xshape)
        # G = dY^T X with the mask and the d(alpha) reduction fused into the epilogue; dY [N, M] and
        # X [N, K] are consumed in place as MN-major operands
        mk = mask if mask.is_contiguous() else mask.contiguous()
        dw, dalpha = tgemm_dw_masked(mn_view(ga), mn_view(xa), M, K, N, mask=mk, packed=ctx.packed)
        return dx, dw, dalpha, dbias, None, None


def gelu_dropout_split(y: torch.Tensor, dropout_p: float, seed, want_lo: bool, slot=None):
    """d = dropout(gelu(y)) as a (hi, lo|None, pitch, 0, slot) GEMM operand; y contiguous fp32 [rows, cols], cols % 8 == 0."""
    rows, cols = y.shape
    dev = nv.device_index(y)
    dt = torch.bfloat16 if slot is None else torch.float16
    hi = torch.empty((rows, cols), dtype=dt, device=y.device)
    lo = torch.empty((rows, cols), dtype=dt, device=y.device) if want_lo else None
    nv.call("atq_gelu_dropout_split", dev, y.data_ptr(), rows, cols, float(dropout_p), nv.ptr(seed), hi.data_ptr(), nv.ptr(lo),
            nv.ptr(slot), nv.stream_ptr(dev))
    return hi, lo, cols, 0, slot


def gelu_dropout_bwd_split_colsum(g: torch.Tensor, y: torch.Tensor, dropout_p: float, seed, want_lo: bool, slot=None):
    """dy = g * keep/(1-p) * gelu'(y) as a GEMM operand + its column sums."""
    rows, cols = y.shape
    dev = nv.device_index(y)
    dt = torch.bfloat16 if slot is None else torch.float16
    hi = torch.empty((rows, cols), dtype=dt, device=y.device)
    lo = torch.empty((rows, cols), dtype=dt, device=y.device) if want_lo else None
    out = torch.empty(cols, dtype=torch.float32, device=y.device)
    ws = nv.workspace(nv.lib.atq_workspace_bytes_split_colsum(rows, cols), y.device)
    nv.call("atq_gelu_dropout_bwd_split_colsum", dev, g.data_ptr(), y.data_ptr(), rows, cols, float(dropout_p), nv.ptr(seed),
            hi.data_ptr(), nv.ptr(lo), out.data_ptr(), ws.data_ptr(), ws.numel(), nv.ptr(slot), nv.stream_ptr(dev))
    return (hi, lo, cols, 0, slot), out


def _keep_bound(p: float) -> float:
    """Upper bound of the dropout keep factor 1/(1-p_eff) (p_eff is p rounded to 1/65536), with head-room."""
    return 1.01 / max(1.0 - float(p) - 2.0 ** -16, 1e-6)


class _RPBFFNFn(torch.autograd.Function):
    """y = linear2(dropout(gelu(linear1(x)))) for two ResidualPrecisionBoostLinear layers
    (models/text_encoder.py:246): the activation, the dropout and the operand split of the hidden tensor are one
    streaming kernel per direction, the [tokens, hidden] gelu / dropout outputs never exist in fp32."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, w1, alpha1, b1, mask1, ops1, w2, alpha2, b2, mask2, ops2, p, seed):
        H, K = w1.shape
        M = w2.shape[0]
        x2 = nv.require_f32(x, "input").reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        N = x2.shape[0]
        lo, f16 = _use_lo(), _use_f16()
        xa = split_operand(x2, x)
        # |dropout(gelu(y))| <= max|y| / (1-p): the hidden operand's scale comes from max|y1|, reduced in the GEMM epilogue
        if f16:
            y1, slot_d = tgemm_absmax(xa, ops1.w, N, H, K, _keep_bound(p), bias=None if b1 is None else b1.detach())
        else:
            (y1, _), slot_d = tgemm(xa, ops1.w, N, H, K, bias=None if b1 is None else b1.detach()), None
        da = gelu_dropout_split(y1, p, seed, lo, slot_d)
        y2, _ = tgemm(da, ops2.w, N, M, H, bias=None if b2 is None else b2.detach())
        tensors = [y1, mask1, mask2, xa[0], da[0]] + ([xa[1], da[1]] if lo else [])
        ctx.save_for_backward(*tensors)
        ctx.seed = seed
        ctx.slots = (xa[4], da[4])
        ctx.cfg = (N, K, H, M, float(p), lo, xa[2], b1 is not None, b2 is not None)
        ctx.ops = (ops1.w_t, ops1.packed, ops2.w_t, ops2.packed)
        ctx.xshape = x.shape
        return y2.reshape(*x.shape[:-1], M)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        N, K, H, M, p, lo, x_pitch, has_b1, has_b2 = ctx.cfg
        saved = ctx.saved_tensors
        y1, mask1, mask2 = saved[0], saved[1], saved[2]
        xa = (saved[3], saved[5] if lo else None, x_pitch, 0, ctx.slots[0])
        da = (saved[4], saved[6] if lo else None, H, 0, ctx.slots[1])
        f16 = saved[3].dtype == torch.float16
        w1_t, packed1, w2_t, packed2 = ctx.ops
        g2 = nv.require_f32(gy, "grad_output").reshape(-1, M)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        if has_b2:
            ga2, db2 = split_operand_colsum(g2, lo, f16)
        else:
            ga2, db2 = (split_scaled(g2, lo) if f16 else split_bf16(g2, lo)), None
        # gradient w.r.t. the dropped activations, fp32 [N, H]; |dd * keep/(1-p) * gelu'(y)| <= max|dd| * 1.13 / (1-p)
        if f16:
            dd, slot_g = tgemm_absmax(ga2, w2_t, N, H, M, 1.13 * _keep_bound(p))
        else:
            (dd, _), slot_g = tgemm(ga2, w2_t, N, H, M), None
        mk2 = mask2 if mask2.is_contiguous() else mask2.contiguous()
        dw2, dalpha2 = tgemm_dw_masked(mn_view(ga2), mn_view(da), M, H, N, mask=mk2, packed=packed2)
        g1a, db1 = gelu_dropout_bwd_split_colsum(dd, y1, p, ctx.seed, lo, slot_g)
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _ = tgemm(g1a, w1_t, N, K, H)
            dx = dx.reshape(ctx.xshape)
        mk1 = mask1 if mask1.is_contiguous() else mask1.contiguous()
        dw1, dalpha1 = tgemm_dw_masked(mn_view(g1a), mn_view(xa), H, K, N, mask=mk1, packed=packed1)
        return (dx, dw1, dalpha1, db1 if has_b1 else None, None, None, dw2, dalpha2, db2, None, None, None, None)


def rpb_ffn_supported(l1, l2, x) -> bool:
    """Both layers ResidualPrecisionBoostLinear-like on CUDA, chained shapes, hidden width % 8 == 0, some tokens."""
    return (hasattr(l1, "precision_mask") and hasattr(l2, "precision_mask") and x.is_cuda and l1.weight.is_cuda
            and l1.weight.shape[0] == l2.weight.shape[1] and l1.weight.shape[0] % 8 == 0 and x.numel() > 0
            and x.shape[-1] == l1.weight.shape[1])


def rpb_ffn(l1, l2, x, dropout_p: float = 0.0, training: bool = True, seed=None, threshold_factor=0.05):
    """linear2(dropout(gelu(linear1(x)))) with the fused activation kernels; l1, l2: ResidualPrecisionBoostLinear."""
    if not rpb_ffn_supported(l1, l2, x):
        raise RuntimeError("atq.fused_ffn: needs two chained ResidualPrecisionBoostLinear layers on CUDA (hidden % 8 == 0)")
    p = float(dropout_p) if training else 0.0
    if p > 0.0 and seed is None:
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=x.device)
    ops1 = layer_operands(l1._ops, l1.weight, l1.alpha, l1.precision_mask, l1.sparsity_target, threshold_factor)
    ops2 = layer_operands(l2._ops, l2.weight, l2.alpha, l2.precision_mask, l2.sparsity_target, threshold_factor)
    return _RPBFFNFn.apply(x, l1.weight, l1.alpha, l1.bias, l1.precision_mask, ops1,
                           l2.weight, l2.alpha, l2.bias, l2.precision_mask, ops2, p, seed if p > 0.0 else None)


class _GatedResidualFn(torch.autograd.Function):
    """out = src + dropout(h) * g  (g: 1-element tensor, e.g. sigmoid(gate); models/text_encoder.py:238-249)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, src, h, g, p, seed):
        src_c = nv.require_f32(src, "src")
        h_c = nv.require_f32(h, "h")
        g_c = nv.require_f32(g.detach().reshape(1), "gate")
        dev = nv.device_index(src_c)
        out = torch.empty_like(src_c)
        nv.call("atq_gated_residual_fwd", dev, src_c.data_ptr(), h_c.data_ptr(), g_c.data_ptr(), src_c.numel(), float(p),
                nv.ptr(seed), out.data_ptr(), nv.stream_ptr(dev))
        ctx.save_for_backward(h_c, g_c)
        ctx.seed, ctx.p, ctx.gshape = seed, float(p), g.shape
        return out.view(src.shape)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        h_c, g_c = ctx.saved_tensors
        d = nv.require_f32(dout, "grad_output")
        dev = nv.device_index(d)
        dh = torch.empty_like(h_c)
        dg = torch.empty(1, dtype=torch.float32, device=d.device)
        ws = nv.workspace(nv.lib.atq_workspace_bytes_gated_residual(d.numel()), d.device)
        nv.call("atq_gated_residual_bwd", dev, d.data_ptr(), h_c.data_ptr(), g_c.data_ptr(), d.numel(), ctx.p, nv.ptr(ctx.seed),
                dh.data_ptr(), dg.data_ptr(), ws.data_ptr(), ws.numel(), nv.stream_ptr(dev))
        return dout, dh.view(dout.shape), dg.view(ctx.gshape), None, None


def gated_residual_supported(src, h, g) -> bool:
    return (src.is_cuda and src.dtype == torch.float32 and h.shape == src.shape and g.numel() == 1
            and src.numel() > 0 and src.numel() % 4 == 0)


def gated_residual(src, h, g, dropout_p: float = 0.0, training: bool = True, seed=None):
    """src + dropout(h) * g in one pass (and one pass backward); the dropout mask is regenerated from a counter hash."""
    if not gated_residual_supported(src, h, g):
        raise RuntimeError("atq.gated_residual: needs fp32 CUDA tensors of equal shape (numel % 4 == 0) and a 1-element gate")
    p = float(dropout_p) if training else 0.0
    if p > 0.0 and seed is None:
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=src.device)
    return _GatedResidualFn.apply(src, h, g, p, seed if p > 0.0 else None)


class _LayerNormFn(torch.autograd.Function):
    """LayerNorm over the last dimension (the op in front of every ternary GEMM of the transformer block,
    models/text_encoder.py:77,232,244): one pass forward that also leaves max|y| in a scale slot for the operand split
    that follows, one pass backward with deterministic gamma / beta gradients."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, eps):
        cols = x.shape[-1]
        x2 = nv.require_f32(x, "input").reshape(-1, cols)
        w_c = nv.require_f32(weight.detach(), "weight")
        b_c = nv.require_f32(bias.detach(), "bias")
        rows = x2.shape[0]
        dev = nv.device_index(x2)
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x2.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x2.device)
        slot = nv.new_slot(x2.device) if _use_f16() else None
        nv.call("atq_layernorm_fwd", dev, x2.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), rows, cols, float(eps), y.data_ptr(),
                mean.data_ptr(), rstd.data_ptr(), nv.ptr(slot), nv.stream_ptr(dev))
        ctx.save_for_backward(x2, w_c, mean, rstd)
        out = y.view(x.shape)
        if slot is not None:
            _note_absmax(out, slot)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        x2, w_c, mean, rstd = ctx.saved_tensors
        rows, cols = x2.shape
        d2 = nv.require_f32(dout, "grad_output").reshape(rows, cols)
        dev = nv.device_index(d2)
        dx = torch.empty_like(x2)
        dw = torch.empty(cols, dtype=torch.float32, device=d2.device)
        db = torch.empty(cols, dtype=torch.float32, device=d2.device)
        ws = nv.workspace(nv.lib.atq_workspace_bytes_layernorm_bwd(cols), d2.device)
        nv.call("atq_layernorm_bwd", dev, d2.data_ptr(), x2.data_ptr(), w_c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, cols,
                dx.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), nv.stream_ptr(dev))
        return dx.view(dout.shape), dw, db, None


def layer_norm_supported(x, weight, bias) -> bool:
    cols = x.shape[-1] if x.dim() else 0
    return (x.is_cuda and x.dtype == torch.float32 and weight is not None and bias is not None and x.numel() > 0
            and 4 <= cols <= 1024 and cols % 4 == 0 and weight.numel() == cols and bias.numel() == cols)


def layer_norm(x, weight, bias, eps: float = 1e-5):
    """F.layer_norm(x, (cols,), weight, bias, eps) on the fused kernels (fp32, 4 <= cols <= 1024, cols % 4 == 0)."""
    if not layer_norm_supported(x, weight, bias):
        raise RuntimeError("atq.layer_norm: needs an fp32 CUDA tensor with 4 <= last dim <= 1024 (multiple of 4), affine weight and bias")
    return _LayerNormFn.apply(x, weight, bias, eps)


def ternary_linear(x, weight, alpha, bias, cache: LayerOperands, sparsity_target=0.3, threshold_factor=0.05):
    ops = layer_operands(cache, weight, None, None, sparsity_target, threshold_factor)
    return _TernaryLinearFn.apply(x, weight, alpha, bias, ops)


def rpb_linear(x, weight, alpha, bias, mask, cache: LayerOperands, sparsity_target=0.3, threshold_factor=0.05):
    ops = layer_operands(cache, weight, alpha, mask, sparsity_target, threshold_factor)
    return _RPBLinearFn.apply(x, weight, alpha, bias, mask, ops)
