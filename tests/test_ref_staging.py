"""CPU: the staged reference (oracle/_ref) is a byte-for-byte copy of the mounted reference, and the checker process
(oracle/ref_runner.py) really runs the reference's own atq package."""
import filecmp
import os

import pytest
import torch

from conftest import REFERENCE, REF_STAGED, have_reference, have_staged_reference, run_reference

pytestmark = pytest.mark.skipif(not have_staged_reference(), reason="oracle/_ref not staged (python oracle/install_ref.py)")


@pytest.mark.skipif(not have_reference(), reason="reference not mounted (GPU box)")
def test_staged_copy_is_identical_to_the_reference():
    n = 0
    for pkg in ("atq", "models", "utils"):
        for name in sorted(os.listdir(os.path.join(REFERENCE, pkg))):
            if name.endswith(".py"):
                assert filecmp.cmp(os.path.join(REFERENCE, pkg, name), os.path.join(REF_STAGED, pkg, name), shallow=False), (pkg, name)
                n += 1
    assert n >= 15
    # nothing else is staged: no scripts, no data, and nothing of it is tracked by git
    assert sorted(d for d in os.listdir(REF_STAGED) if os.path.isdir(os.path.join(REF_STAGED, d)) and d != "__pycache__") == ["atq", "models", "utils"]


def test_checker_process_runs_the_references_own_atq(tmp_path):
    res = run_reference("classifier", dict(seed=0, data_seed=0, batch=32, sparsity=0.3), tmp_path)
    assert "oracle/_ref/atq" in res["atq_file"]
    assert sorted(res["state"]) [:2] == ["classifier.0.alpha", "classifier.0.bias"]
    # fp32 and fp64 runs of the reference agree to fp32 accuracy; the RPB gradient contract holds in its own code
    assert torch.allclose(res["f32"]["logits"].double(), res["f64"]["logits"], rtol=1e-4, atol=1e-5)
    g, mask = res["f32"]["grads"]["classifier.0.weight"], res["state"]["classifier.0.precision_mask"]
    assert int((g != 0).sum()) <= int(mask.sum()) == int(0.05 * mask.numel())
    assert float((g * (1 - mask)).abs().max()) == 0.0


def test_product_never_imports_the_staged_reference():
    pkg = os.path.join(os.path.dirname(REF_STAGED), "..", "atq-multimodal_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text and "_ref" not in text.replace("dot_ref", "").replace("x_ref", ""), os.path.join(root, f)
