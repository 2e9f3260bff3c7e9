"""Recipe: stage the UNMODIFIED reference next to the oracle.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (ak736/ATQ-Multimodal) is pure Python with no setup.py, so nothing can be pip-installed;
its three importable packages -- ``atq`` (the hot path's own implementation), ``models`` and ``utils``
(the callers either side of it) -- are copied byte for byte from /root/reference into the git-ignored
``oracle/_ref/`` so that they travel to the GPU box with the snapshot (the box has no /root/reference).
Nothing under ``oracle/_ref`` is tracked, edited, or imported by the product package.

Used by
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg (the reference's own CPU path, timed),
  * ``tests/test_gpu_dropin.py`` (the reference's models running unmodified on the B200 ``atq``, compared
    with the same models on the reference's ``atq`` on the CPU),
  * ``bench.py``'s ``dropin`` arm (the cost of staying strictly behind the reference's public names).

    python oracle/install_ref.py            # idempotent; prints what it did
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ATQ_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
PACKAGES = ("atq", "models", "utils")


def installed() -> bool:
    return all(os.path.isfile(os.path.join(REF_DST, p, "__init__.py")) for p in PACKAGES)


def install(verbose: bool = False) -> bool:
    """Copy the packages if the source tree is mounted.  Returns True when oracle/_ref is usable."""
    if not os.path.isdir(os.path.join(REF_SRC, "atq")):
        return installed()
    os.makedirs(REF_DST, exist_ok=True)
    for pkg in PACKAGES:
        src, dst = os.path.join(REF_SRC, pkg), os.path.join(REF_DST, pkg)
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            if not name.endswith(".py"):
                continue
            s, d = os.path.join(src, name), os.path.join(dst, name)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                os.chmod(d, 0o644)
                if verbose:
                    print(f"copied {pkg}/{name}")
    with open(os.path.join(REF_DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"byte-for-byte copy of {REF_SRC}/{{{','.join(PACKAGES)}}}/*.py made by oracle/install_ref.py\n"
                "git-ignored; never edited; not part of the product\n")
    return installed()


if __name__ == "__main__":
    ok = install(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT available (no /root/reference and no earlier copy)")
    sys.exit(0 if ok else 1)
