"""Generate tests/golden/syncbn_reference.npz by EXECUTING the reference's parallel SyncBN branch on the CPU
(authoring container only):
    python tests/golden/make_golden_syncbn.py
/root/reference/code/sync_batchnorm/batchnorm.py: forward :68-98 takes its parallel branch when `_is_parallel` is set; the
cross-replica part is `_sync_master.run_master(msg)` -> `_data_parallel_master` (:110-131) = ReduceAddCoalesced over the
replicas' (sum, ssum) + `_compute_mean_std` (:133-150) + Broadcast.  With the WHOLE batch in one replica the reduce-add and
the broadcast are identities, so run_master is bound straight to the module's own `_compute_mean_std` — every formula
(sums, mean, biased / unbiased variance, clamp(eps)^-1/2, momentum update, fused affine) is the reference's code; autograd
gives the backward.  Two training steps, so the running statistics carry the momentum recursion.
The world-size-2 gloo test feeds each rank half of the batch and must reproduce these numbers.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/code")
from sync_batchnorm.batchnorm import SynchronizedBatchNorm2d        # noqa: E402  (reference, unmodified)


def main():
    g = torch.Generator().manual_seed(41)
    C = 6
    bn = SynchronizedBatchNorm2d(C).train()
    with torch.no_grad():
        bn.weight.copy_(0.5 + torch.rand(C, generator=g))
        bn.bias.copy_(0.2 * torch.randn(C, generator=g))
    weight0, bias0 = bn.weight.detach().clone(), bn.bias.detach().clone()
    bn._is_parallel, bn._parallel_id = True, 0
    bn._sync_master.run_master = lambda msg: bn._compute_mean_std(msg.sum, msg.ssum, msg.sum_size)
    out = {"weight": weight0.numpy(), "bias": bias0.numpy()}
    for step in range(2):
        x = (torch.randn(8, C, 5, 7, generator=g) * (1 + step) + 0.3).requires_grad_(True)
        w = torch.randn(8, C, 5, 7, generator=g)
        bn.zero_grad()
        y = bn(x)
        (y * w).sum().backward()
        out.update({f"x{step}": x.detach().numpy(), f"w{step}": w.numpy(), f"y{step}": y.detach().numpy(),
                    f"dx{step}": x.grad.numpy(), f"dweight{step}": bn.weight.grad.numpy().copy(),
                    f"dbias{step}": bn.bias.grad.numpy().copy(), f"running_mean{step}": bn.running_mean.numpy().copy(),
                    f"running_var{step}": bn.running_var.numpy().copy()})
    # a channel whose variance is below eps: the reference clamps the VARIANCE (not var + eps)
    x = torch.randn(4, C, 3, 3, generator=g)
    x[:, 0] = 0.25 + 1e-4 * torch.randn(4, 3, 3, generator=g)
    out["x_small"], out["y_small"] = x.numpy(), bn(x).detach().numpy()
    p = os.path.join(HERE, "syncbn_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
