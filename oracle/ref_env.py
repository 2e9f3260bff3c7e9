"""Import environment for the staged reference (oracle/_ref).  TEST / BASELINE INFRASTRUCTURE ONLY.

Two arrangements, one per process (both bind the package name ``atq``, so they cannot coexist):

  activate("reference") : ``atq``, ``models``, ``utils`` all come from oracle/_ref -- the reference exactly
                          as published, on the CPU.  This is the baseline arm and the checker.
  activate("b200")      : ``atq`` is THIS repo's package (atq-multimodal_b200/atq, CUDA only), ``models`` and
                          ``utils`` are the reference's unmodified files -- the drop-in boundary of SURVEY 8(b).

The only adaptations are environmental, none touches the reference's arithmetic: empty ``matplotlib`` stubs
(utils/__init__.py:2 imports the plotting helper; matplotlib is not in this image) and
``torchvision.models.resnet18(weights=None)`` (models/multimodal_classifier.py:30 would download ImageNet
weights; there is no network, and BASELINE's configs are synthetic / random-init anyway).
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
PKG_DIR = os.path.join(ROOT, "atq-multimodal_b200")


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF, p, "__init__.py")) for p in ("atq", "models", "utils"))


def _stub_matplotlib():
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))


def _offline_resnet():
    import torchvision.models as tvm
    if getattr(tvm.resnet18, "_atq_offline", False):
        return
    orig = tvm.resnet18

    def resnet18(*args, weights=None, **kw):  # random init: no download
        return orig(*args, weights=None, **kw)

    resnet18._atq_offline = True
    tvm.resnet18 = resnet18


def activate(atq_impl: str) -> None:
    if atq_impl not in ("reference", "b200"):
        raise ValueError(atq_impl)
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/install_ref.py` where /root/reference is mounted")
    want_dir = os.path.join(REF if atq_impl == "reference" else PKG_DIR, "atq")
    loaded = sys.modules.get("atq")
    if loaded is not None and os.path.dirname(os.path.abspath(loaded.__file__)) != want_dir:
        raise RuntimeError(f"another `atq` ({loaded.__file__}) is already imported in this process; "
                           f"the {atq_impl} arrangement needs a fresh process")
    paths = [REF] if atq_impl == "reference" else [PKG_DIR, REF]
    for p in (PKG_DIR, REF):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = paths
    _stub_matplotlib()
    _offline_resnet()
