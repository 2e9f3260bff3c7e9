"""CPU, world_size 2 over gloo: the data-parallel wiring (embedding all-gather + summed gradient
all-reduce) reproduces the single-process global-batch loss and gradients."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import PKG_DIR, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from atq import parallel
    from workloads import models as M
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)  # identical init everywhere
    enc_i, enc_t = torch.nn.Linear(12, 8), torch.nn.Linear(10, 8)
    unused = torch.nn.Linear(3, 3)
    params = list(enc_i.parameters()) + list(enc_t.parameters()) + list(unused.parameters())
    g = torch.Generator().manual_seed(123)
    xi, xt = torch.randn(6 * world, 12, generator=g), torch.randn(6 * world, 10, generator=g)
    man = M.ContrastiveManager(M.HardNegativeInfoNCE())
    sync = parallel.FlatGradAllReduce(params, bucket_bytes=256)  # tiny buckets: several all-reduces
    sl = slice(6 * rank, 6 * rank + 6)
    for step in range(2):
        sync.zero_grad()
        img = parallel.gather_embeddings(enc_i(xi[sl]))
        txt = parallel.gather_embeddings(enc_t(xt[sl]))
        assert img.shape == (6 * world, 8)
        loss = man.compute_loss(img, txt)
        loss.backward()
        sync.reduce()
    assert unused.weight.grad is None  # never bound: optimizer would skip it
    torch.save({"loss": loss.detach(), "grads": [p.grad.clone() for p in params[:4]]}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from workloads import models as M
    torch.manual_seed(0)
    enc_i, enc_t = torch.nn.Linear(12, 8), torch.nn.Linear(10, 8)
    g = torch.Generator().manual_seed(123)
    xi, xt = torch.randn(6 * world, 12, generator=g), torch.randn(6 * world, 10, generator=g)
    man = M.ContrastiveManager(M.HardNegativeInfoNCE())
    loss = man.compute_loss(enc_i(xi), enc_t(xt))
    loss.backward()
    want = [p.grad for p in list(enc_i.parameters()) + list(enc_t.parameters())]
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in range(world))
    assert torch.allclose(r0["loss"], loss.detach(), atol=1e-6) and torch.allclose(r1["loss"], loss.detach(), atol=1e-6)
    for a, b, w in zip(r0["grads"], r1["grads"], want):
        assert torch.equal(a, b)  # all-reduced: identical on every rank
        assert torch.allclose(a, w, atol=1e-5, rtol=1e-4)


def test_gather_is_identity_without_process_group():
    from atq import parallel
    x = torch.randn(3, 4, requires_grad=True)
    assert parallel.gather_embeddings(x) is x
    assert parallel.init_from_env() == (0, 1, 0)


def _sparse_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from atq import parallel
    from workloads import models as M
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(0)
    L = M.oracle_layers()
    enc = L.ResidualPrecisionBoostLinear(12, 8, precision_ratio=0.3, sparsity_target=0.2)
    head = torch.nn.Linear(8, 4)
    params = list(enc.parameters()) + list(head.parameters())
    masks = parallel.rpb_masks(torch.nn.Sequential(enc, head))
    assert list(masks) == [enc.weight]
    g = torch.Generator().manual_seed(7 + rank)  # different data per rank
    x = torch.randn(5, 12, generator=g)
    results = {}
    for kind, sync in (("dense", parallel.FlatGradAllReduce(params)),
                       ("sparse", parallel.FlatGradAllReduce(params, sparse_masks=masks))):
        for _ in range(2):
            sync.zero_grad()
            head(enc(x)).square().sum().backward()
            sync.reduce()
        results[kind] = [p.grad.clone() for p in params if p.grad is not None]
        results[kind + "_flat"] = sync.flat.numel()
    nnz = int((enc.precision_mask != 0).sum())
    pad4 = lambda n: (n + 3) // 4 * 4  # slices are padded to 16 bytes
    assert results["dense_flat"] - results["sparse_flat"] == pad4(enc.weight.numel()) - pad4(nnz)
    for a, b in zip(results["dense"], results["sparse"]):
        assert torch.equal(a, b)
    wg = results["sparse"][0]
    assert torch.count_nonzero(wg * (1 - enc.precision_mask)) == 0 and torch.count_nonzero(wg) > 0
    torch.save(results["sparse"], os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sparse_rpb_gradient_all_reduce_matches_dense(tmp_path):
    """only the entries under precision_mask travel; result identical to the dense all-reduce."""
    world = 2
    mp.spawn(_sparse_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (torch.load(tmp_path / f"s{r}.pt") for r in range(world))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def _overlap_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from atq import parallel
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 9), torch.nn.Tanh(), torch.nn.Linear(9, 4))
    unused = torch.nn.Linear(3, 3)
    params = list(net.parameters()) + list(unused.parameters())
    g = torch.Generator().manual_seed(7 + rank)
    results = {}
    for kind, sync in (("plain", parallel.FlatGradAllReduce(params)),
                       ("overlap", parallel.FlatGradAllReduce(params, bucket_bytes=400, overlap=True))):
        g.manual_seed(7 + rank)
        for step in range(3):
            x = torch.randn(5, 12, generator=g)
            sync.zero_grad()
            net(x).square().sum().backward()
            sync.reduce()
        results[kind] = [p.grad.clone() for p in net.parameters()]
        if kind == "overlap":
            assert len(sync._buckets) >= 3 and all(not b["launched"] for b in sync._buckets)  # re-armed for the next step
            assert sync.active[0] is list(net.parameters())[-1]                                 # reverse layout
    assert unused.weight.grad is None
    for a, b in zip(results["plain"], results["overlap"]):
        assert torch.equal(a, b)
    torch.save(results["overlap"], os.path.join(out_dir, f"o{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_bucketed_all_reduce_matches_plain(tmp_path):
    """Buckets launched from post-accumulate-grad hooks while backward is still running give the gradients of the
    single flat all-reduce, on every rank."""
    world = 2
    mp.spawn(_overlap_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (torch.load(tmp_path / f"o{r}.pt") for r in range(world))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
