"""GPU parity: atq.optim.FlatAdamW (one-launch multi-tensor AdamW kernel) against torch.optim.AdamW, eagerly and
replayed from a CUDA graph (device-side step counter)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from atq.optim import FlatAdamW

DEV = "cuda:0"
HP = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)


def _model(seed):
    torch.manual_seed(seed)
    m = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Flatten(), nn.Linear(8 * 6 * 6, 37), nn.ReLU(), nn.Linear(37, 5)).to(DEV)
    m[0].to(memory_format=torch.channels_last)
    m.unused = nn.Parameter(torch.randn(1030, device=DEV))  # never receives a gradient: must stay untouched
    return m


def _grads(m, step):
    g = torch.Generator(device=DEV).manual_seed(100 + step)
    x = torch.randn(4, 3, 6, 6, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    m.zero_grad(set_to_none=True)
    m(x).square().mean().backward()


def test_flat_adamw_matches_torch_adamw():
    a, b = _model(0), _model(0)
    oa, ob = FlatAdamW(a.parameters(), **HP), torch.optim.AdamW(b.parameters(), **HP)
    for step in range(6):
        _grads(a, step); _grads(b, step)
        oa.step(); ob.step()
    for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6), (n, (pa - pb).abs().max().item())
    assert torch.equal(a.unused, _model(0).unused)


def test_flat_adamw_cuda_graph_replays_advance_the_step_counter():
    a, b = _model(1), _model(1)
    oa, ob = FlatAdamW(a.parameters(), **HP), torch.optim.AdamW(b.parameters(), **HP)
    x = torch.randn(4, 3, 6, 6, device=DEV).contiguous(memory_format=torch.channels_last)

    def step_a():
        oa.zero_grad(set_to_none=True)
        a(x).square().mean().backward()
        oa.step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step_a()                       # warm-up step 1 (eager)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    oa.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        step_a()                       # capture does not execute
    for _ in range(3):
        graph.replay()                 # steps 2, 3, 4
    torch.cuda.synchronize()
    for _ in range(4):
        ob.zero_grad(set_to_none=True)
        b(x).square().mean().backward()
        ob.step()
    for (n, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-4, atol=1e-6), (n, (pa - pb).abs().max().item())


def test_flat_adamw_accepts_unaligned_gradient_views():
    """p.grad as a view into a flat buffer at an odd element offset (what a flat gradient all-reduce produces)."""
    torch.manual_seed(3)
    pa = nn.Parameter(torch.randn(1000, device=DEV))
    pb = nn.Parameter(pa.detach().clone())
    flat = torch.randn(1003, device=DEV)
    pa.grad = flat[3:]            # 12-byte offset
    pb.grad = flat[3:].clone()
    oa, ob = FlatAdamW([pa], **HP), torch.optim.AdamW([pb], **HP)
    for _ in range(3):
        oa.step(); ob.step()
    assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6)
