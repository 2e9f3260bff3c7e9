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
REF_STAGED = os.path.join(ROOT, "oracle", "_ref")  # byte-for-byte copy made by oracle/install_ref.py (travels to the GPU box)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    lib = os.path.join(PKG_DIR, "atq", "libatq_sm100.so")
    if not os.path.exists(lib):
        # build container: nvcc cross-compiles sm_100a without a GPU
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)
    if not have_staged_reference() and have_reference():
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "install_ref.py")], cwd=ROOT, check=True,
                       stdout=subprocess.DEVNULL)


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


def have_staged_reference():
    return all(os.path.isfile(os.path.join(REF_STAGED, p, "__init__.py")) for p in ("atq", "models", "utils"))


def reference_dir():
    """The reference tree to import from: the staged copy (also present on the GPU box) or the mount."""
    return REF_STAGED if have_staged_reference() else REFERENCE


def run_reference(task, cfg, tmp_path):
    """Run oracle/ref_runner.py (the unmodified reference on ITS OWN atq, CPU, own process) and load its result."""
    import json
    import torch
    out = os.path.join(str(tmp_path), f"ref_{task}.pt")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--task", task, "--out", out, "--cfg",
                    json.dumps(cfg)], cwd=ROOT, check=True, env=env, stdout=subprocess.DEVNULL)
    return torch.load(out)


def load_reference_atq():
    """Import the reference's atq core (quantizers, layers, precision_boost, routing, bit_packing) under the alias
    `ref_atq`, next to this repo's `atq`: the reference's own layer code, usable as the checker on any device."""
    import importlib.util
    import types
    if "ref_atq" in sys.modules:
        return sys.modules["ref_atq"]
    REFERENCE = reference_dir()
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
