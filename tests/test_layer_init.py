"""CPU: the B200 modules construct on the CPU (only their forward needs CUDA) with exactly the reference's state -
same state_dict keys in the same order, same initial values for the same torch seed (reset_parameters
must reproduce init), checked against the oracle modules and, when /root/reference is mounted, the reference itself."""
import os
import sys

import pytest
import torch

import atq
from oracle import atq_oracle as O

REF = os.environ.get("ATQ_REFERENCE", "/root/reference")


def _same_state(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]


@pytest.mark.parametrize("bias", [True, False])
def test_init_matches_oracle_modules(bias):
    torch.manual_seed(11)
    a = atq.TernaryLinear(48, 24, bias=bias)
    torch.manual_seed(11)
    b = O.OracleTernaryLinear(48, 24, bias=bias)
    _same_state(a, b)
    assert not hasattr(a, "sparsity_target") and not hasattr(a, "get_quantized_weights")
    torch.manual_seed(12)
    a = atq.ResidualPrecisionBoostLinear(40, 56, 0.15, bias, 0.25)
    torch.manual_seed(12)
    b = O.OracleRPBLinear(40, 56, 0.15, bias, 0.25)
    _same_state(a, b)
    assert int(a.precision_mask.sum()) == int(0.15 * 40 * 56)
    assert (a.precision_ratio, a.sparsity_target, a.in_features, a.out_features) == (0.15, 0.25, 40, 56)
    # reset_parameters reproduces the constructor's draw
    torch.manual_seed(12)
    a.precision_mask.zero_()
    a.reset_parameters()
    _same_state(a, b)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "atq")), reason="reference checkout not mounted")
def test_init_matches_live_reference():
    saved = {k: v for k, v in sys.modules.items() if k == "atq" or k.startswith("atq.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        from atq.layers import TernaryLinear as RefTL
        from atq.precision_boost import ResidualPrecisionBoostLinear as RefRPB
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "atq" or k.startswith("atq.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    torch.manual_seed(5)
    r1, r2 = RefTL(30, 20), RefRPB(30, 20, 0.1, True, 0.3)
    torch.manual_seed(5)
    m1, m2 = atq.TernaryLinear(30, 20), atq.ResidualPrecisionBoostLinear(30, 20, 0.1, True, 0.3)
    _same_state(m1, r1)
    _same_state(m2, r2)
