"""2-bit ternary codec on B200 (drop-in for atq/bit_packing.py)."""
import torch

from . import _engine as eng
from . import _native as nv


class TernaryBitPacking:
    """-1 -> 00, 0 -> 01, +1 -> 10; four values per byte, value i in bits 2*(i%4).. of byte i//4,
    flat row-major order, zero tail bits -- the reference's format byte for byte."""

    ENCODING = {0: -1, 1: 0, 2: 1}

    @staticmethod
    def pack_ternary_weights(ternary_weights):
        t = ternary_weights.detach()
        if t.dtype != torch.float32:
            t = t.float()
        packed, flag = eng.pack2_from_f32(t.reshape(-1))
        # the only host sync of the codec: the reference validates before packing (:36-39)
        if int(flag.item()) != 0:
            raise ValueError("Input must contain only ternary values (-1, 0, 1)")
        return {
            'packed_weights': packed,
            'original_shape': ternary_weights.shape,
            'metadata': {'num_values': t.numel(), 'encoding': dict(TernaryBitPacking.ENCODING)},
        }

    @staticmethod
    def unpack_ternary_weights(packed_data):
        packed = packed_data['packed_weights']
        n = packed_data['metadata']['num_values']
        flag = torch.zeros(1, dtype=torch.int32, device=packed.device)
        out = eng.unpack2(packed, n, torch.float32, flag)
        if int(flag.item()) != 0:
            raise KeyError(3)  # code 0b11 has no entry in the encoding table (atq/bit_packing.py:116)
        return out.reshape(packed_data['original_shape'])

    @staticmethod
    def compute_memory_savings(original_tensor):
        n = original_tensor.numel()
        original_bytes = n * 4
        packed_bytes = (n * 2 + 7) // 8
        return {
            'original_bytes': original_bytes,
            'packed_bytes': packed_bytes,
            'compression_ratio': original_bytes / packed_bytes,
            'memory_reduction': 1.0 - (packed_bytes / original_bytes),
        }

    @staticmethod
    def fast_ternary_matmul(packed_data, input_tensor, alpha=1.0):
        """(input @ T^T) * alpha straight from the packed bytes: the codec is expanded to a bf16
        B operand (exact) and contracted on the tensor cores (K7); atq/bit_packing.py:149-176."""
        shape = tuple(packed_data['original_shape'])
        if len(shape) != 2:
            raise RuntimeError("fast_ternary_matmul expects 2-D packed weights")
        M, K = shape
        packed = packed_data['packed_weights']
        x = nv.require_f32(input_tensor, "input_tensor")
        if x.shape[-1] != K:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and {K}x{M})")
        x2 = x.reshape(-1, K)
        if x2.shape[0] == 0:
            return x.new_zeros(*x.shape[:-1], M) * alpha
        xa = eng.split_operand(x2)
        f16 = xa[0].dtype == torch.float16
        packed = packed.contiguous()
        if eng.packed_gemm_ok(K, packed) and eng._PACKED != "never":
            # the GEMM reads the codec bytes and expands them in shared memory
            y, _ = eng.tgemm_packed(xa, packed, x2.shape[0], M, K)
        else:
            tb = eng.unpack2(packed, M * K, torch.bfloat16).view(M, K)
            if f16:
                tb = tb.to(torch.float16)  # -1 / 0 / +1: exact in either 16-bit format
            pitch = nv.round_up(K, 8)
            if pitch != K:
                tb = torch.nn.functional.pad(tb, (0, pitch - K))
            y, _ = eng.tgemm(xa, (tb, None, pitch), x2.shape[0], M, K)
        return y.reshape(*x.shape[:-1], M) * alpha
