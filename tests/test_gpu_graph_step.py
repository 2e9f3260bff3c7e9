"""GPU: the whole-step CUDA graph of the bench harness (workloads.train.GraphedRetrievalStep) against the same steps
run eagerly, and its re-capture when a host-side scalar baked into the capture moves (learning rate, per-layer
sparsity targets, loss epoch): the reference changes those between epochs (train_multimodal.py:403-413,
atq/mixed_precision_atq.py:323-401), so a replay must never train with the schedule frozen at capture."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import atq
from atq.mixed_precision_atq import GradualQuantizationScheduler
from atq.optim import FlatAdamW
from workloads import train as T

DEV = "cuda:0"
CFG = T.RetrievalCfg(name="graph-step test", vocab=300, embed_dim=64, hidden_dim=128, image_size=64, batch=8,
                     text_heads=4, text_layers=2, lr=2e-5)


def _build():
    model, crit, man = T.build_retrieval(atq, CFG, seed=7)
    model.to(DEV).train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0  # dropout streams of two separately built steps differ; everything else is deterministic
    sched = GradualQuantizationScheduler(model, CFG.total_epochs, 0.3, 0.2, warmup_epochs=CFG.warmup_epochs)
    sched.step(CFG.epoch)
    opt = T.make_optimizer(model, CFG, capturable=True, fused=True, adamw_cls=FlatAdamW)
    return model, crit, man, sched, opt


def _batches():
    return [tuple(t.to(DEV) for t in b) for b in T.synthetic_batches(CFG, 4, seed=3)]


def _schedule(i, opt, sched, crit, man):
    """Host-side schedule edits between steps, the same for both runs."""
    if i == 2:   # next epoch: new per-layer sparsity targets, new loss temperature / curriculum stage
        sched.step(CFG.epoch + 2)
        crit.set_epoch(CFG.epoch + 2, CFG.total_epochs)
        man.set_epoch(CFG.epoch + 2, CFG.total_epochs)
    if i == 4:   # learning rate 0: the weights stop moving, so the same batch must give the same loss twice
        for g in opt.param_groups:
            g["lr"] = 0.0
            g["weight_decay"] = 0.0


ORDER = [1, 2, 3, 1, 2, 2]  # batch index per step (batch 0 is the capture / warm-up batch)


def _run(graphed):
    batches = _batches()
    model, crit, man, sched, opt = _build()
    prepare = atq.prepare_quantization
    if graphed:
        step = T.GraphedRetrievalStep(model, man, opt, batches[0], prepare=prepare)  # 3 eager warm-up steps inside
    else:
        for _ in range(3):
            T.retrieval_step(model, man, opt, batches[0], None, None, prepare)

        def step(b):
            return T.retrieval_step(model, man, opt, b, None, None, prepare)
    losses = []
    for i, bi in enumerate(ORDER):
        _schedule(i, opt, sched, crit, man)
        losses.append(float(step(batches[bi]).detach()))
    return losses, (step.recaptures if graphed else None)


def test_graphed_step_tracks_eager_and_recaptures_on_schedule_change():
    eager, _ = _run(False)
    graph, recaptures = _run(True)
    assert recaptures == 2, recaptures            # once for the epoch change (step 2), once for the learning rate (step 4)
    for i, (a, b) in enumerate(zip(eager, graph)):
        # same kernels on the same data; cuDNN's backward and the atomics of a few reductions are not bit-reproducible,
        # and the network is badly conditioned at init: Adam turns rounding-level gradient differences
        # into +-lr parameter moves, so the two trajectories agree to per-cent level, not to rounding level
        assert abs(a - b) <= 2e-2 * abs(a) + 1e-3, (i, eager, graph)
    # steps 4 and 5 run the same batch with lr = 0: a stale capture (lr baked at 2e-5) would move the weights in between
    assert abs(graph[4] - graph[5]) <= 1e-5 * abs(graph[4]) + 1e-6, graph
    assert abs(eager[4] - eager[5]) <= 1e-5 * abs(eager[4]) + 1e-6, eager
    # the trajectory is a real one: the loss moved while the learning rate was non-zero
    assert abs(graph[0] - graph[3]) > 1e-5 * abs(graph[0]), graph
