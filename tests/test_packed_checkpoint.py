"""CPU: the ATQP packed-checkpoint container (atq/packed_checkpoint.py) round-trips tensors
bit for bit; the loader rejects foreign files.  (The exporter / inference layers need the GPU: test_gpu_packed_checkpoint.py.)"""
import importlib.util
import os
import struct

import numpy as np
import pytest
import torch

from conftest import PKG_DIR


def _load_module():
    # the container code is pure host code; import it without importing the atq package (which needs the .so + CUDA)
    spec = importlib.util.spec_from_file_location("atq_packed_checkpoint_host", os.path.join(PKG_DIR, "atq", "packed_checkpoint.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_container_round_trip(tmp_path):
    pc = _load_module()
    g = torch.Generator().manual_seed(0)
    tensors = {
        "a.packed_weights": torch.randint(0, 256, (1001,), dtype=torch.uint8, generator=g),
        "a.alpha": torch.tensor([0.37]),
        "a.residual_index": torch.arange(0, 70, 7, dtype=torch.int32),
        "a.residual_value": torch.randn(10, generator=g),
        "norm.weight": torch.randn(3, 5, generator=g),
        "bn.num_batches_tracked": torch.tensor(12, dtype=torch.int64),
        "flag": torch.tensor([True, False, True]),
    }
    meta = {"layers": {"a": {"kind": "rpb", "out_features": 77, "in_features": 52, "sparsity_target": 0.3,
                             "precision_ratio": 0.05, "num_values": 4004, "encoding": pc.ENCODING}}}
    path = str(tmp_path / "m.atq")
    size = pc.write_container(path, meta, tensors)
    assert size == os.path.getsize(path)
    header, back = pc.read_container(path)
    assert header["version"] == pc.VERSION and header["layers"] == meta["layers"]
    assert set(back) == set(tensors)
    for k, t in tensors.items():
        assert back[k].dtype == t.dtype and tuple(back[k].shape) == tuple(t.shape) and torch.equal(back[k], t), k
        assert header["tensors"][k]["offset"] % 64 == 0
    raw = open(path, "rb").read()
    assert raw[:4] == b"ATQP" and struct.unpack("<I", raw[4:8])[0] == 1
    # payload of the codec bytes is stored verbatim
    e = header["tensors"]["a.packed_weights"]
    start = 16 + struct.unpack("<Q", raw[8:16])[0]
    start += (-start) % 64
    assert raw[start + e["offset"]: start + e["offset"] + e["nbytes"]] == tensors["a.packed_weights"].numpy().tobytes()


def test_container_rejects_foreign_files(tmp_path):
    pc = _load_module()
    p = tmp_path / "x.bin"
    p.write_bytes(b"NOPE" + b"\0" * 64)
    with pytest.raises(ValueError):
        pc.read_container(str(p))
    p.write_bytes(b"ATQP" + struct.pack("<IQ", 99, 2) + b"{}")
    with pytest.raises(ValueError):
        pc.read_container(str(p))
    with pytest.raises(ValueError):
        pc.write_container(str(tmp_path / "y.atq"), {}, {"h": torch.zeros(2, dtype=torch.float16)})
