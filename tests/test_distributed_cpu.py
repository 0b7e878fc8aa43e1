"""N>1 host logic on CPU (gloo, world size 2): SyncBN statistics / gradients equal the single-process full
batch (the reference's parallel formulas, sync_batchnorm/batchnorm.py:133-150), gradient all-reduce averages."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, q):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sync_batchnorm import SynchronizedBatchNorm2d
    from gan_training import _allreduce_grads
    torch.manual_seed(0)
    x_full = torch.randn(8, 6, 5, 7)
    w_full = torch.randn(8, 6, 5, 7)
    bn = SynchronizedBatchNorm2d(6, affine=False).train()
    xs = x_full[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
    y = bn(xs)
    (y * w_full[rank * 4:(rank + 1) * 4]).sum().backward()
    # gradient all-reduce (mean) of a toy parameter set
    p1, p2 = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))
    p1.grad, p2.grad = torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))
    _allreduce_grads([p1, p2], world)
    # plain numpy payloads: tensors in an mp.Queue travel through shared-memory handles that die with the worker
    q.put((rank, y.detach().numpy(), xs.grad.numpy(), bn.running_mean.numpy().copy(), bn.running_var.numpy().copy(),
           p1.grad.numpy().copy(), p2.grad.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_and_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    res = [(r[0],) + tuple(torch.from_numpy(a) for a in r[1:]) for r in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference with the parallel formulas: inv_std = clamp(var_biased, eps)^-1/2
    torch.manual_seed(0)
    x = torch.randn(8, 6, 5, 7, requires_grad=True)
    w = torch.randn(8, 6, 5, 7)
    n = x.numel() // 6
    mean = x.sum((0, 2, 3)) / n
    var = ((x * x).sum((0, 2, 3)) - x.sum((0, 2, 3)) * mean) / n
    y = (x - mean.view(1, 6, 1, 1)) * var.clamp(min=1e-5).pow(-0.5).view(1, 6, 1, 1)
    (y * w).sum().backward()
    y_d = torch.cat([r[1] for r in res])
    g_d = torch.cat([r[2] for r in res])
    assert torch.allclose(y_d, y.detach(), atol=1e-5)
    assert torch.allclose(g_d, x.grad, atol=1e-5)
    for r in res:       # running stats identical on every rank: momentum 0.1, unbiased variance
        assert torch.allclose(r[3], 0.1 * mean.detach(), atol=1e-6)
        assert torch.allclose(r[4], 0.9 + 0.1 * var.detach() * n / (n - 1), atol=1e-5)
        assert torch.allclose(r[5], torch.full((3,), 1.5)) and torch.allclose(r[6], torch.full((2, 2), 15.0))


def test_syncbn_single_process_is_plain_batchnorm():
    from sync_batchnorm import SynchronizedBatchNorm2d, convert_model
    torch.manual_seed(1)
    x = torch.randn(4, 3, 5, 5)
    a, b = SynchronizedBatchNorm2d(3).train(), torch.nn.BatchNorm2d(3).train()
    assert torch.allclose(a(x), b(x), atol=1e-6) and torch.allclose(a.running_var, b.running_var)
    assert list(a.state_dict()) == list(b.state_dict())
    m = convert_model(torch.nn.Sequential(torch.nn.Conv2d(3, 3, 1), torch.nn.BatchNorm2d(3)))
    assert isinstance(m[1], SynchronizedBatchNorm2d)
