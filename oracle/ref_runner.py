#!/usr/bin/env python
"""Runs the UNMODIFIED reference (oracle/_ref: its own atq + models + utils) on the CPU in a process of its own and
writes what it computed to a file.  TEST / BASELINE INFRASTRUCTURE ONLY -- the checker side of the drop-in tests.

    python oracle/ref_runner.py --task retrieval|classifier|block --out result.pt [--cfg '{"json": ...}']

Every task runs twice: in float32 (what the reference computes) and in float64 (`model.double()`, the anchor that
tells how far fp32 itself is from the exact answer; T is identical because the fp32 weights cast exactly).
The state_dict is saved so the other side loads the very same parameters and masks.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_env  # noqa: E402

ref_env.activate("reference")

import torch  # noqa: E402

from oracle import ref_tasks as R  # noqa: E402


def _to64(t):
    return t.double() if torch.is_tensor(t) and t.is_floating_point() else t


def task_retrieval(cfg):
    model = R.build_retrieval(cfg["vocab"], cfg["embed_dim"], cfg["hidden_dim"], cfg["seed"])
    if cfg.get("schedule_epoch") is not None:
        R.step_schedule(model, cfg["schedule_epoch"], cfg["total_epochs"])
    state = copy.deepcopy(model.state_dict())
    sparsity = {n: float(m.sparsity_target) for n, m in model.named_modules() if hasattr(m, "sparsity_target")}
    batch = R.synthetic_retrieval_batch(cfg["batch"], cfg["image_size"], cfg["vocab"], seed=cfg["data_seed"])
    _, man = R.build_loss(model, cfg["loss_epoch"], cfg["total_epochs"])
    f32 = R.retrieval_forward_backward(model, man, batch)
    model.double()
    f64 = R.retrieval_forward_backward(model, man, tuple(_to64(t) for t in batch))
    return {"state": state, "sparsity": sparsity, "batch": batch, "f32": f32, "f64": f64}


def task_classifier(cfg):
    model = R.build_classifier(cfg["seed"])
    state = copy.deepcopy(model.state_dict())
    g = torch.Generator().manual_seed(cfg["data_seed"])
    x = torch.randn(cfg["batch"], 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (cfg["batch"],), generator=g)
    f32 = R.classifier_forward_backward(model, x, y, cfg["sparsity"])
    model.double()
    f64 = R.classifier_forward_backward(model, x.double(), y, cfg["sparsity"])
    return {"state": state, "x": x, "y": y, "f32": f32, "f64": f64}


def task_block(cfg):
    block = R.build_block(cfg["embed_dim"], cfg["num_heads"], cfg["dim_feedforward"], cfg["seed"])
    state = copy.deepcopy(block.state_dict())
    g = torch.Generator().manual_seed(cfg["data_seed"])
    x = torch.randn(cfg["batch"], cfg["seq"], cfg["embed_dim"], generator=g)
    gy = torch.randn(cfg["batch"], cfg["seq"], cfg["embed_dim"], generator=g)
    pad = torch.zeros(cfg["batch"], cfg["seq"], dtype=torch.bool)
    for i, n in enumerate(cfg.get("pad_from", [])):
        if n is not None:
            pad[i, n:] = True
    f32 = R.block_forward_backward(block, x, pad, gy)
    block.double()
    f64 = R.block_forward_backward(block, x.double(), pad, gy.double())
    return {"state": state, "x": x, "gy": gy, "pad": pad, "f32": f32, "f64": f64}


TASKS = {"retrieval": task_retrieval, "classifier": task_classifier, "block": task_block}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", choices=sorted(TASKS), required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--cfg", default="{}")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    import atq
    assert os.path.dirname(os.path.abspath(atq.__file__)).startswith(ref_env.REF), "checker must run the reference's atq"
    out = TASKS[args.task](json.loads(args.cfg))
    out["atq_file"] = atq.__file__
    torch.save(out, args.out)


if __name__ == "__main__":
    main()
