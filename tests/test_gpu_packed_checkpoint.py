"""GPU: export a model to the ATQP packed format, load it into a fresh skeleton as inference-only layers and
compare with the training modules in eval mode."""
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

import atq
from atq import packed_checkpoint as PC
from atq.bit_packing import TernaryBitPacking

DEV = "cuda:0"


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.embed = nn.Embedding(50, 128)
        self.fc1 = atq.TernaryLinear(128, 192)                                  # packed GEMM path (K % 64 == 0)
        self.norm = nn.LayerNorm(192)
        self.fc2 = atq.ResidualPrecisionBoostLinear(192, 72, precision_ratio=0.2, sparsity_target=0.25)
        self.fc3 = atq.TernaryLinear(72, 10, bias=False)                        # K % 64 != 0: bf16 copy path

    def forward(self, tok):
        h = torch.relu(self.fc1(self.embed(tok)))
        return self.fc3(torch.relu(self.fc2(self.norm(h))))


@pytest.mark.parametrize("mode", ["parity", "fast"])
def test_packed_checkpoint_round_trip_and_inference(tmp_path, mode):
    atq.set_gemm_mode(mode)
    try:
        torch.manual_seed(0)
        net = _Net().to(DEV).eval()
        with torch.no_grad():
            net.fc1.alpha.fill_(0.7)
            net.fc2.alpha.fill_(1.3)
        tok = torch.randint(0, 50, (4, 9), device=DEV)
        with torch.no_grad():
            want = net(tok)
        path = str(tmp_path / "net.atq")
        info = PC.save_packed(net, path)
        assert info["file_bytes"] == os.path.getsize(path)
        assert info["ternary_weights"] == 128 * 192 + 192 * 72 + 72 * 10
        meta, tensors = PC.read_container(path)
        # the codec bytes are exactly pack(T) for the T the quantizer returns (E1 format)
        t1, _ = atq.adaptive_ternary_quantization(net.fc1.weight.detach(), net.fc1.alpha)
        assert torch.equal(tensors["fc1.packed_weights"], TernaryBitPacking.pack_ternary_weights(t1)["packed_weights"].cpu())
        assert tensors["fc2.residual_index"].numel() == int(0.2 * 192 * 72)
        assert "fc1.weight" not in tensors and "fc2.precision_mask" not in tensors and "embed.weight" in tensors
        torch.manual_seed(123)  # different init: everything must come from the file
        fresh = PC.load_packed(path, _Net(), device=DEV)
        assert isinstance(fresh.fc1, PC.PackedTernaryLinear) and isinstance(fresh.fc2, PC.PackedRPBLinear)
        with torch.no_grad():
            got = fresh(tok)
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), (got - want).abs().max()
    finally:
        atq.set_gemm_mode("parity")


def test_packed_checkpoint_compression_on_ternary_heavy_model(tmp_path):
    torch.manual_seed(0)
    net = nn.Sequential(atq.TernaryLinear(1024, 1024), nn.ReLU(), atq.TernaryLinear(1024, 1024)).to(DEV)
    info = PC.save_packed(net, str(tmp_path / "m.atq"))
    assert info["compression_ratio"] > 15.0, info   # 2 bits + alpha/bias per weight vs 32 bits
