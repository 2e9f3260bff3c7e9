"""On-disk format for packed ATQ models and the inference-only layers that load it.

The reference never persists packed weights (checkpoints are fp32 state_dicts, train.py:302; `*.atq` only
appears in .gitignore:81), so its "16x / 15-22 MB" claims (readme.md:26,30) are not measurable there.  This
module defines the missing piece:

  file  = b"ATQP" | u32 version | u64 header_len | header (UTF-8 JSON) | payload (64-byte aligned tensors)
  header = {"version": 1,
            "layers":  {module_name: {"kind": "ternary" | "rpb", "out_features", "in_features",
                                       "sparsity_target", "precision_ratio", "num_values", "encoding"}},
            "tensors": {tensor_name: {"dtype", "shape", "offset", "nbytes"}}}

Per ternary layer `name`:  `name.packed_weights` (uint8, the E1 codec of atq/bit_packing.py:60-69: code =
value + 1, four per byte LSB first, flat row-major, zero tail), `name.alpha` (fp32 [1]), `name.bias` (fp32 [M],
optional); RPB layers add the sparse fp32 residual `name.residual_index` (int32, flat positions where
precision_mask != 0, ascending) and `name.residual_value` (fp32 W at those positions).  Everything that is
not a ternary layer (embeddings, LayerNorm, convolution trunks, ...) is stored dense under its state_dict key.
All integers little-endian.

`convert_to_inference(model, meta, tensors)` swaps every TernaryLinear / ResidualPrecisionBoostLinear of a
model skeleton for `PackedTernaryLinear` / `PackedRPBLinear` and loads the dense rest; TernaryLinear inference
runs the GEMM straight on the 2-bit bytes (`atq_tgemm_packed`), RPB rebuilds its mixed bf16 operand once.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

MAGIC = b"ATQP"
VERSION = 1
_ALIGN = 64
_DTYPES = {"uint8": np.uint8, "int32": np.int32, "int64": np.int64, "float32": np.float32, "bool": np.bool_}
ENCODING = {"0": -1, "1": 0, "2": 1}


# ---------------------------------------------------------------------------------------------
# container (pure host code: numpy + struct; no GPU needed)
# ---------------------------------------------------------------------------------------------

def write_container(path: str, meta: dict, tensors: Dict[str, torch.Tensor]) -> int:
    """Returns the file size in bytes."""
    table, blobs, off = {}, [], 0
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        dt = str(t.dtype).replace("torch.", "")
        if dt not in _DTYPES:
            raise ValueError(f"packed checkpoint: unsupported dtype {t.dtype} for {name}")
        raw = t.numpy().astype(_DTYPES[dt], copy=False).tobytes()
        pad = (-off) % _ALIGN
        off += pad
        blobs.append((pad, raw))
        table[name] = {"dtype": dt, "shape": list(t.shape), "offset": off, "nbytes": len(raw)}
        off += len(raw)
    header = json.dumps({"version": VERSION, "layers": meta.get("layers", {}), "tensors": table,
                         "extra": meta.get("extra", {})}, sort_keys=True).encode()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<IQ", VERSION, len(header)) + header)
        pos = len(MAGIC) + 12 + len(header)
        f.write(b"\0" * ((-pos) % _ALIGN))  # payload starts 64-byte aligned
        for pad, raw in blobs:
            f.write(b"\0" * pad)
            f.write(raw)
        return f.tell()


def read_container(path: str) -> Tuple[dict, Dict[str, torch.Tensor]]:
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != MAGIC:
        raise ValueError(f"{path}: not an ATQP packed checkpoint")
    version, hlen = struct.unpack("<IQ", data[4:16])
    if version != VERSION:
        raise ValueError(f"{path}: packed checkpoint version {version}, this reader handles {VERSION}")
    header = json.loads(data[16:16 + hlen].decode())
    start = 16 + hlen
    start += (-start) % _ALIGN
    tensors = {}
    for name, e in header["tensors"].items():
        lo = start + e["offset"]
        arr = np.frombuffer(data, dtype=_DTYPES[e["dtype"]], count=int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1,
                            offset=lo).reshape(e["shape"]).copy()
        tensors[name] = torch.from_numpy(arr)
    return header, tensors


# ---------------------------------------------------------------------------------------------
# export (GPU: thresholds / codec bytes come from the same kernels the training forward uses)
# ---------------------------------------------------------------------------------------------

@torch.no_grad()
def export_packed(model: nn.Module) -> Tuple[dict, Dict[str, torch.Tensor]]:
    from . import _engine as eng
    from .layers import TernaryLinear
    from .precision_boost import ResidualPrecisionBoostLinear
    layers, tensors, owned = {}, {}, set()
    for name, m in model.named_modules():
        is_rpb = isinstance(m, ResidualPrecisionBoostLinear)
        if not (is_rpb or isinstance(m, TernaryLinear)):
            continue
        M, K = m.weight.shape
        s = float(m.sparsity_target) if is_rpb else 0.3
        ops = eng.layer_operands(m._ops, m.weight, m.alpha if is_rpb else None, m.precision_mask if is_rpb else None, s)
        layers[name] = {"kind": "rpb" if is_rpb else "ternary", "out_features": M, "in_features": K,
                        "sparsity_target": s, "precision_ratio": float(getattr(m, "precision_ratio", 0.0)),
                        "num_values": M * K, "encoding": ENCODING}
        tensors[name + ".packed_weights"] = ops.packed
        tensors[name + ".alpha"] = m.alpha.detach()
        if m.bias is not None:
            tensors[name + ".bias"] = m.bias.detach()
        if is_rpb:
            idx = torch.nonzero(m.precision_mask.reshape(-1) != 0).reshape(-1)
            tensors[name + ".residual_index"] = idx.to(torch.int32)
            tensors[name + ".residual_value"] = m.weight.detach().reshape(-1)[idx]
        owned.update(f"{name}.{k}" for k in ("weight", "alpha", "bias", "precision_mask"))
    for key, t in model.state_dict().items():
        if key not in owned:
            tensors[key] = t.detach().float() if t.is_floating_point() else t.detach()
    return {"layers": layers}, tensors


def save_packed(model: nn.Module, path: str) -> dict:
    """Writes the packed checkpoint; returns sizes for the compression claim."""
    meta, tensors = export_packed(model)
    size = write_container(path, meta, tensors)
    dense = sum(t.numel() * t.element_size() for t in model.state_dict().values())
    tern = sum(l["num_values"] for l in meta["layers"].values())
    return {"file_bytes": size, "fp32_state_dict_bytes": dense, "ternary_weights": tern,
            "compression_ratio": dense / size}


# ---------------------------------------------------------------------------------------------
# inference-only layers
# ---------------------------------------------------------------------------------------------

class PackedTernaryLinear(nn.Module):
    """y = alpha * (x T^T) + b with T held as 2-bit codec bytes (inference; no fp32 weight on the device)."""

    def __init__(self, in_features: int, out_features: int, packed: torch.Tensor, alpha: torch.Tensor, bias=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("packed_weights", packed)
        self.register_buffer("alpha", alpha.float().reshape(1))
        self.register_buffer("bias", None if bias is None else bias.float())
        self._tb = None  # bf16 copy of T, only for shapes the packed GEMM does not take (K % 64 != 0)

    @torch.no_grad()
    def forward(self, x):
        from . import _engine as eng
        from . import _native as nv
        M, K = self.out_features, self.in_features
        x2 = nv.require_f32(x, "input").reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        N = x2.shape[0]
        if N == 0:
            return x2.new_zeros((*x.shape[:-1], M))
        xa = eng.split_operand(x2)
        if eng.packed_gemm_ok(K, self.packed_weights):
            y, _ = eng.tgemm_packed(xa, self.packed_weights, N, M, K, scale=self.alpha, bias=self.bias)
        else:
            f16 = xa[0].dtype == torch.float16
            if self._tb is None or self._tb[0].device != x2.device or (self._tb[0].dtype == torch.float16) != f16:
                t = eng.unpack2(self.packed_weights, M * K, torch.float32).reshape(M, K)
                self._tb = eng.split_bf16(t, False, nv.unit_slot(t.device) if f16 else None)  # T is exact in either format
            y, _ = eng.tgemm(xa, self._tb, N, M, K, scale=self.alpha, bias=self.bias)
        return y.reshape(*x.shape[:-1], M)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, packed 2-bit, bias={self.bias is not None}"


class PackedRPBLinear(nn.Module):
    """y = x Wm^T + b with Wm = alpha*T off the residual positions and the stored fp32 weights on them
    (atq/precision_boost.py:72).  Stored: 2-bit T, sparse fp32 residual; the bf16 hi/lo operand of Wm is
    rebuilt once on first use."""

    def __init__(self, in_features: int, out_features: int, packed, alpha, residual_index, residual_value, bias=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("packed_weights", packed)
        self.register_buffer("alpha", alpha.float().reshape(1))
        self.register_buffer("residual_index", residual_index)
        self.register_buffer("residual_value", residual_value.float())
        self.register_buffer("bias", None if bias is None else bias.float())
        self._w = None
        self._mode = None

    @torch.no_grad()
    def _operand(self, device):
        from . import _engine as eng
        mode = eng.get_gemm_mode()
        if self._w is None or self._mode != mode or self._w[0].device != device:
            M, K = self.out_features, self.in_features
            wm = eng.unpack2(self.packed_weights, M * K, torch.float32) * self.alpha  # alpha * T
            wm[self.residual_index.long()] = self.residual_value                      # fp32 weights under the mask
            self._w = eng.split_operand(wm.reshape(M, K))
            self._mode = mode
        return self._w

    @torch.no_grad()
    def forward(self, x):
        from . import _engine as eng
        from . import _native as nv
        M, K = self.out_features, self.in_features
        x2 = nv.require_f32(x, "input").reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        N = x2.shape[0]
        if N == 0:
            return x2.new_zeros((*x.shape[:-1], M))
        xa = eng.split_operand(x2)
        y, _ = eng.tgemm(xa, self._operand(x2.device), N, M, K, bias=self.bias)
        return y.reshape(*x.shape[:-1], M)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features}, packed 2-bit + "
                f"{self.residual_index.numel()} fp32 residuals, bias={self.bias is not None}")


def convert_to_inference(model: nn.Module, meta: dict, tensors: Dict[str, torch.Tensor], device=None) -> nn.Module:
    """Swap every ternary layer named in `meta` for its packed inference module (in place) and load the dense
    remainder; returns the model in eval mode."""
    device = device if device is not None else next(model.parameters()).device
    for name, info in meta["layers"].items():
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model

        def g(key):
            t = tensors.get(f"{name}.{key}")
            return None if t is None else t.to(device)
        if info["kind"] == "rpb":
            mod = PackedRPBLinear(info["in_features"], info["out_features"], g("packed_weights"), g("alpha"),
                                  g("residual_index"), g("residual_value"), g("bias"))
        else:
            mod = PackedTernaryLinear(info["in_features"], info["out_features"], g("packed_weights"), g("alpha"), g("bias"))
        setattr(parent, leaf, mod)
    prefixes = tuple(n + "." for n in meta["layers"])
    dense = {k: v for k, v in tensors.items() if not k.startswith(prefixes)}
    missing, unexpected = model.load_state_dict(dense, strict=False)
    missing = [k for k in missing if not k.startswith(prefixes)]
    if missing or unexpected:
        raise RuntimeError(f"packed checkpoint does not match the model: missing {missing}, unexpected {unexpected}")
    return model.to(device).eval()


def load_packed(path: str, model: nn.Module, device=None) -> nn.Module:
    meta, tensors = read_container(path)
    return convert_to_inference(model, meta, tensors, device)
