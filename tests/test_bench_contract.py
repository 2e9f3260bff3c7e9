"""CPU: bench.py's reference arm prints the contract's JSON line (the GPU arm needs a B200)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "multimodal ATQ train samples/sec"
    assert line["unit"] == "samples/s" and line["higher_is_better"] is True and line["value"] > 0
    from conftest import have_staged_reference
    # the unmodified reference (oracle/_ref) is the baseline whenever it is staged; the port is only a fallback
    assert line["cpu_baseline"]["kind"] == ("reference" if have_staged_reference() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert set(line["config"]) == {"workload", "per_gpu_batch", "global_batch", "parallelism", "l2"}
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "train_multimodal.py synthetic Flickr8k shape" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
