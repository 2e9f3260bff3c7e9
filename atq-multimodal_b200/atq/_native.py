"""ctypes binding of libatq_sm100.so (the C ABI declared in include/atq_sm100.h).

There is no CPU fallback and no alternative backend: importing this module without the built
library raises ImportError, and every wrapper raises RuntimeError for tensors that are not
contiguous fp32 CUDA tensors on an sm_100-class device.  PyTorch is used here only for device
memory (the caching allocator owns every buffer, including workspaces) and for the current
stream handle.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_LIB_PATH = os.environ.get("ATQ_SM100_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libatq_sm100.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(
        f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C atq-multimodal_b200/csrc`). The atq package has no CPU or eager fallback.")
_lib = ctypes.CDLL(_LIB_PATH)

ABI_VERSION = 6


class BF16Operand(Structure):
    _fields_ = [("hi", c_void_p), ("lo", c_void_p), ("pitch", c_int64), ("mn_major", c_int32), ("format", c_int32),
                ("inv_scale", c_void_p)]


_P = c_void_p
_SIGS = {
    "atq_abi_version": (c_int, []),
    "atq_last_error_string": (c_char_p, []),
    "atq_device_check": (c_int, [c_int]),
    "atq_num_sms": (c_int, [c_int]),
    "atq_kernel_launch_count": (ctypes.c_uint64, []),
    "atq_workspace_bytes_abs_stats": (c_size_t, [c_int64]),
    "atq_abs_stats": (c_int, [c_int, _P, c_int64, _P, _P, c_size_t, _P]),
    "atq_workspace_bytes_select_kth_abs": (c_size_t, [c_int64]),
    "atq_select_kth_abs": (c_int, [c_int, _P, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "atq_workspace_bytes_adaptive_threshold": (c_size_t, [c_int64]),
    "atq_adaptive_threshold": (c_int, [c_int, _P, c_int64, c_int64, c_float, _P, _P, c_size_t, _P]),
    "atq_workspace_bytes_adaptive_threshold_batched": (c_size_t, [c_int, POINTER(c_int64)]),
    "atq_adaptive_threshold_batched": (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                               c_float, POINTER(c_void_p), _P, c_size_t, _P]),
    "atq_ternarize_f32": (c_int, [c_int, _P, c_int64, _P, _P, _P, _P]),
    "atq_ternarize_pack2": (c_int, [c_int, _P, c_int64, _P, _P, _P, _P]),
    "atq_optimal_alpha": (c_int, [c_int, _P, _P, c_int64, _P, _P]),
    "atq_pack2_from_f32": (c_int, [c_int, _P, c_int64, _P, _P, _P]),
    "atq_unpack2_to_f32": (c_int, [c_int, _P, c_int64, _P, _P, _P]),
    "atq_unpack2_to_bf16": (c_int, [c_int, _P, c_int64, _P, _P]),
    "atq_unpack2_to_i8": (c_int, [c_int, _P, c_int64, _P, _P]),
    "atq_ternarize_pack2_batched": (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_void_p), POINTER(c_void_p), _P]),
    "atq_pack2_from_f32_batched": (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_void_p), _P, _P]),
    "atq_unpack2_to_f32_batched": (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_void_p), _P, _P]),
    "atq_route_mask_mul": (c_int, [c_int, _P, _P, _P, c_int64, _P, _P]),
    "atq_absmax_scale": (c_int, [c_int, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P]),
    "atq_absmax_scale_batched": (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_void_p), POINTER(c_void_p),
                                         c_float, _P]),
    "atq_split_scaled_fused_max_elems": (c_int64, []),
    "atq_split_scaled_fused": (c_int, [c_int, _P, c_int64, _P, _P, c_float, _P, _P, _P]),
    "atq_split_bf16": (c_int, [c_int, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P, _P]),
    "atq_workspace_bytes_split_colsum": (c_size_t, [c_int64, c_int64]),
    "atq_split_bf16_colsum": (c_int, [c_int, _P, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P, _P]),
    "atq_split_bf16_t": (c_int, [c_int, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P, _P, _P]),
    "atq_build_ternary_operands": (c_int, [c_int, _P, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int, _P]),
    "atq_build_mixed_operands": (c_int, [c_int, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, c_int64, _P, _P,
                                         c_int64, _P, _P]),
    "atq_set_cta_pairs": (c_int, [c_int]),
    "atq_workspace_bytes_tgemm": (c_size_t, [c_int64, c_int64]),
    "atq_tgemm": (c_int, [c_int, c_int64, c_int64, c_int64, POINTER(BF16Operand), POINTER(BF16Operand), _P, _P, _P,
                          c_int64, _P, c_int64, _P, _P, c_size_t, _P]),
    "atq_tgemm_absmax": (c_int, [c_int, c_int64, c_int64, c_int64, POINTER(BF16Operand), POINTER(BF16Operand), _P, _P, _P,
                                 c_int64, _P, c_float, _P]),
    "atq_tgemm_packed": (c_int, [c_int, c_int64, c_int64, c_int64, POINTER(BF16Operand), _P, _P, _P, _P,
                                 c_int64, _P, c_int64, _P, _P, c_size_t, _P]),
    "atq_tgemm_fwd": (c_int, [c_int, c_int64, c_int64, c_int64, POINTER(BF16Operand), POINTER(BF16Operand), _P, _P,
                              _P, c_int64, _P, c_size_t, _P]),
    "atq_tgemm_dx": (c_int, [c_int, c_int64, c_int64, c_int64, POINTER(BF16Operand), POINTER(BF16Operand), _P, _P,
                             c_int64, _P, c_int64, _P, _P, c_size_t, _P]),
    "atq_workspace_bytes_tgemm_dw": (c_size_t, [c_int64, c_int64, c_int64]),
    "atq_tgemm_dw_masked": (c_int, [c_int, c_int64, c_int64, c_int64, POINTER(BF16Operand), POINTER(BF16Operand), _P,
                                    _P, _P, c_int64, _P, _P, c_size_t, _P]),
    "atq_workspace_bytes_colsum": (c_size_t, [c_int64, c_int64]),
    "atq_colsum_f32": (c_int, [c_int, _P, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "atq_gelu_dropout_split": (c_int, [c_int, _P, c_int64, c_int64, c_float, _P, _P, _P, _P, _P]),
    "atq_gelu_dropout_bwd_split_colsum": (c_int, [c_int, _P, _P, c_int64, c_int64, c_float, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "atq_set_fused_select": (None, [c_int]),
    "atq_layernorm_fwd": (c_int, [c_int, _P, _P, _P, c_int64, c_int64, c_float, _P, _P, _P, _P, _P]),
    "atq_workspace_bytes_layernorm_bwd": (c_size_t, [c_int64]),
    "atq_layernorm_bwd": (c_int, [c_int, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "atq_workspace_bytes_gated_residual": (c_size_t, [c_int64]),
    "atq_gated_residual_fwd": (c_int, [c_int, _P, _P, _P, c_int64, c_float, _P, _P, _P]),
    "atq_gated_residual_bwd": (c_int, [c_int, _P, _P, _P, c_int64, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "atq_adamw_multi": (c_int, [c_int, _P, _P, _P, c_int, c_float, c_float, c_float, c_float, c_float, _P, _P]),
    "atq_attention_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_float, c_float,
                                  _P, c_int, _P, c_int64, _P, _P]),
    "atq_attention_bwd": (c_int, [c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_float, c_float,
                                  _P, c_int, _P, c_int64, _P, c_int64, _P, _P, c_int64, _P, c_int64, _P, c_int64, _P]),
    "atq_rowkth_largest": (c_int, [c_int, _P, c_int64, c_int64, c_int64, _P, _P]),
    "atq_infonce_row_stats": (c_int, [c_int, _P, c_int64, c_int64, _P, _P, _P, c_float, _P, _P, _P, _P, _P]),
    "atq_infonce_finalize": (c_int, [c_int, c_int64, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P]),
    "atq_infonce_grad": (c_int, [c_int, _P, c_int64, c_int64, _P, _P, _P, c_float, _P, _P, _P, _P, _P, _P, c_float, c_float, _P,
                                 _P, c_int64, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(_lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

if _lib.atq_abi_version() != ABI_VERSION:
    raise ImportError(f"libatq_sm100.so ABI {_lib.atq_abi_version()} != binding ABI {ABI_VERSION}; rebuild")

lib = _lib
_checked_devices: set = set()
gpu_launches = 0  # number of library compute calls issued


def kernel_launch_count() -> int:
    """Kernels launched by libatq_sm100 in this process (bench.py reports the per-step delta)."""
    return int(_lib.atq_kernel_launch_count())


class ATQNativeError(RuntimeError):
    pass


def last_error() -> str:
    s = _lib.atq_last_error_string()
    return s.decode() if s else ""


def check(status: int, what: str) -> None:
    if status != 0:
        raise ATQNativeError(f"{what} failed with status {status}: {last_error()}")


def device_index(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("atq (B200 build) only runs on CUDA tensors: there is no CPU fallback; "
                           f"got a tensor on {t.device}")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        check(_lib.atq_device_check(idx), "atq_device_check")
        _checked_devices.add(idx)
    return idx


def require_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"atq: {name} must be float32, got {t.dtype}")
    device_index(t)
    return t if t.is_contiguous() else t.contiguous()


def stream_ptr(dev: int) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def workspace(nbytes: int, device) -> torch.Tensor:
    # the caching allocator makes this a free-list pop; stream-ordered reuse is handled by torch
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def call(name: str, *args) -> None:
    global gpu_launches
    gpu_launches += 1
    check(getattr(_lib, name)(*args), name)


def operand(hi: torch.Tensor, lo, pitch: int, mn_major: int = 0, slot=None) -> BF16Operand:
    """mn_major=1: the tensor is [kdim, rows] row-major (a row-major activation/weight used transposed).
    The element format follows the tensor dtype (bfloat16 / float16); `slot` is the scale slot of a scaled-fp16
    operand (4 floats; slot[2] = 1/scale is handed to the GEMM epilogue)."""
    fmt = 1 if hi.dtype == torch.float16 else 0
    return BF16Operand(hi.data_ptr(), None if lo is None else lo.data_ptr(), pitch, int(mn_major), fmt,
                       None if slot is None else slot.data_ptr() + 8)


_SLOT_ARENAS: dict = {}
_ARENAS_USED_IN_CAPTURE: list = []   # never freed: a captured graph keeps raw pointers into them
_ARENA_FLOATS = 1 << 16              # 16 384 slots


def new_slot(device) -> torch.Tensor:
    """A zeroed 4-float scale slot (include/atq_sm100.h: atq_absmax_scale), carved from a per-device arena so that
    handing one out costs no kernel launch; slots are never handed out twice.  Arenas are allocated outside
    CUDA-graph capture (ordinary allocator memory) and, once a capture has drawn from one, kept alive for the life
    of the process, because graph replays write through the raw pointers."""
    key = (device.type, device.index)
    capturing = torch.cuda.is_current_stream_capturing()
    arena = _SLOT_ARENAS.get(key)
    low = arena is None or arena[1] + 4 > arena[0].numel() or (not capturing and arena[1] > arena[0].numel() // 2 and arena[2])
    if low:
        # (inside a capture this allocation lands in the graph's private pool; it is pinned below like the others)
        arena = [torch.zeros(_ARENA_FLOATS, dtype=torch.float32, device=device), 0, False]
        _SLOT_ARENAS[key] = arena
    if capturing and not arena[2]:
        arena[2] = True
        _ARENAS_USED_IN_CAPTURE.append(arena[0])
    off = arena[1]
    arena[1] = off + 4
    return arena[0][off: off + 4]


_UNIT_SLOTS: dict = {}


def unit_slot(device) -> torch.Tensor:
    """Scale slot with scale 1 (exactly representable operands such as ternary weights in fp16)."""
    key = (device.type, device.index)
    s = _UNIT_SLOTS.get(key)
    if s is None:
        s = torch.tensor([0.0, 1.0, 1.0, 0.0], dtype=torch.float32, device=device)
        _UNIT_SLOTS[key] = s
    return s
