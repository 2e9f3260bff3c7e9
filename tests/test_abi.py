"""CPU: the C-ABI library loads and exports every symbol include/atq_sm100.h declares; the
package imports without a GPU; there is no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, PKG_DIR


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "atq_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(PKG_DIR, "atq", "libatq_sm100.so"))
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/atq_sm100.h but not exported"


def test_binding_covers_header():
    import atq._native as nv
    assert sorted(nv.EXPORTED_SYMBOLS) == _header_symbols()
    assert nv.lib.atq_abi_version() == nv.ABI_VERSION


def test_workspace_queries_are_host_only():
    import atq._native as nv
    assert nv.lib.atq_workspace_bytes_select_kth_abs(1 << 20) >= 16 * 1024
    assert nv.lib.atq_workspace_bytes_tgemm(4096, 4096) >= 4 * 32 * 64
    assert nv.lib.atq_workspace_bytes_colsum(1000, 768) == 16 * 768 * 4


def test_public_surface_matches_reference():
    import atq
    import atq.bit_packing
    import atq.mixed_precision_atq as mp
    assert atq.__all__ == ['adaptive_ternary_quantization', 'TernaryLinear', 'SelectiveGradientRouting',
                           'apply_selective_routing', 'ResidualPrecisionBoostLinear']
    for n in ("pack_ternary_weights", "unpack_ternary_weights", "compute_memory_savings", "fast_ternary_matmul"):
        assert hasattr(atq.bit_packing.TernaryBitPacking, n)
    for n in ("MixedPrecisionATQ", "GradualQuantizationScheduler", "PrecisionControlledLinear",
              "EnhancedATQTransformerLayer"):
        assert hasattr(mp, n)
    tl = atq.TernaryLinear(8, 4)
    rpb = atq.ResidualPrecisionBoostLinear(8, 4)
    # duck-typing contract
    assert not hasattr(tl, "sparsity_target") and not hasattr(tl, "get_quantized_weights")
    assert hasattr(rpb, "sparsity_target") and hasattr(rpb, "get_quantized_weights")
    assert sorted(tl.state_dict()) == ["alpha", "bias", "weight"]
    assert sorted(rpb.state_dict()) == ["alpha", "bias", "precision_mask", "weight"]
    x = torch.zeros(3, 8)
    assert atq.apply_selective_routing(x) is x


def test_no_cpu_fallback():
    import atq
    from atq.bit_packing import TernaryBitPacking
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        atq.adaptive_ternary_quantization(torch.randn(4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        atq.TernaryLinear(8, 4)(torch.randn(2, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TernaryBitPacking.pack_ternary_weights(torch.zeros(4))
    assert TernaryBitPacking.compute_memory_savings(torch.zeros(4096, 4096)) == {
        'original_bytes': 67108864, 'packed_bytes': 4194304, 'compression_ratio': 16.0, 'memory_reduction': 0.9375}


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
