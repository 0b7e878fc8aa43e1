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


def _worker_golden(rank, world, port, q):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sync_batchnorm import SynchronizedBatchNorm2d
    d = np.load(os.path.join(ROOT, "tests", "golden", "syncbn_reference.npz"))
    bn = SynchronizedBatchNorm2d(6).train()
    with torch.no_grad():
        bn.weight.copy_(torch.tensor(d["weight"]))
        bn.bias.copy_(torch.tensor(d["bias"]))
    out = []
    for step in range(2):
        lo, hi = rank * 4, rank * 4 + 4
        x = torch.tensor(d[f"x{step}"][lo:hi]).requires_grad_(True)
        bn.zero_grad()
        y = bn(x)
        (y * torch.tensor(d[f"w{step}"][lo:hi])).sum().backward()
        gw, gb = bn.weight.grad.clone(), bn.bias.grad.clone()
        dist.all_reduce(gw)                        # parameter gradients: summed over the replicas (DataParallel semantics)
        dist.all_reduce(gb)
        out.append((y.detach().numpy(), x.grad.numpy(), gw.numpy(), gb.numpy(), bn.running_mean.numpy().copy(),
                    bn.running_var.numpy().copy()))
    xs = torch.tensor(d["x_small"][rank * 2:rank * 2 + 2])
    out.append(bn(xs).detach().numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_world2_equals_the_reference_parallel_branch():
    """tests/golden/syncbn_reference.npz: the reference's own parallel-branch code executed on the whole batch
    (make_golden_syncbn.py); two ranks with half a batch each must reproduce outputs, input / parameter gradients and the
    running statistics after two training steps, and the clamp-the-variance behaviour on a near-constant channel."""
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", "syncbn_reference.npz"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_golden, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        y = np.concatenate([res[r][step][0] for r in range(2)])
        dx = np.concatenate([res[r][step][1] for r in range(2)])
        assert np.abs(y - d[f"y{step}"]).max() < 2e-5 and np.abs(dx - d[f"dx{step}"]).max() < 2e-5
        for r in range(2):
            _, _, gw, gb, rm, rv = res[r][step]
            assert np.abs(gw - d[f"dweight{step}"]).max() < 1e-3 * np.abs(d[f"dweight{step}"]).max()
            assert np.abs(gb - d[f"dbias{step}"]).max() < 1e-3 * np.abs(d[f"dbias{step}"]).max() + 1e-5
            assert np.abs(rm - d[f"running_mean{step}"]).max() < 1e-6 and np.abs(rv - d[f"running_var{step}"]).max() < 1e-5
    ys = np.concatenate([res[r][2] for r in range(2)])
    assert np.abs(ys - d["y_small"]).max() < 1e-3 * np.abs(d["y_small"]).max()
