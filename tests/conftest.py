import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "atq-multimodal_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    lib = os.path.join(PKG_DIR, "atq", "libatq_sm100.so")
    if not os.path.exists(lib):
        # build container: nvcc cross-compiles sm_100a without a GPU
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "atq_golden.npz"))


@pytest.fixture(scope="session")
def policy():
    import json
    with open(os.path.join(GOLDEN, "policy_golden.json")) as f:
        return json.load(f)


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "atq"))


def load_reference_atq():
    """Import the reference's atq package under an alias (build container only)."""
    import importlib.util
    import types
    if "ref_atq" in sys.modules:
        return sys.modules["ref_atq"]
    pkg = types.ModuleType("ref_atq")
    pkg.__path__ = [os.path.join(REFERENCE, "atq")]
    sys.modules["ref_atq"] = pkg
    for name in ("quantizers", "layers", "precision_boost", "routing", "bit_packing"):
        spec = importlib.util.spec_from_file_location(f"ref_atq.{name}", os.path.join(REFERENCE, "atq", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"ref_atq.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return pkg
