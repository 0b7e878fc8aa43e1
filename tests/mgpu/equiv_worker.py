"""Worker of tests/test_multigpu_gpu.py: one G step and one D step of GANTrainer on `world` GPUs (torchrun, NCCL) over a
FIXED global batch (each rank takes its contiguous shard, explicit noise) -> rank 0 writes losses, gradient norms,
BatchNorm running statistics and a parameter probe to an .npz.  world = 1 runs the same global batch on one GPU:
SURVEY §8e's parity recipe ("noise must be drawn so that shard r of N reproduces the 1-GPU stream")."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "2dimageto3dmodel_b200"), ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)


def main(out_path, global_batch):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import bench
    import gan_common as GC
    from gan_training import GANTrainer
    args = bench.gan_args()
    torch.manual_seed(4321)                                   # identical replicas
    tr = GANTrainer(args, mesh_template=None, device=dev)
    g = torch.Generator().manual_seed(99)
    B = global_batch
    z, c, alpha, tex, mesh = GC.inputs(GC.make_args(256, 2), B=B, seed=17)
    noise = [torch.randn(B, 64, generator=g) for _ in range(2)]
    lo, hi = rank * B // world, (rank + 1) * B // world
    sh = lambda t: t[lo:hi].to(dev)
    out = {}
    lg = tr.g_step(sh(alpha), sh(c), noise=sh(noise[0]))
    G, D = tr.trainer.generator, tr.trainer.discriminator
    out["g_names"] = np.array([n for n, p in G.named_parameters() if p.grad is not None])
    out["g_grad_norms"] = np.array([float(p.grad.norm()) for n, p in G.named_parameters() if p.grad is not None])
    ld = tr.d_step(sh(tex), sh(alpha), sh(mesh), sh(c), noise=sh(noise[1]))
    out["d_names"] = np.array([n for n, p in D.named_parameters() if p.grad is not None])
    out["d_grad_norms"] = np.array([float(p.grad.norm()) for n, p in D.named_parameters() if p.grad is not None])
    losses = torch.stack((lg.reshape(()), ld.reshape(())))
    if world > 1:
        dist.all_reduce(losses)
        losses /= world                                       # per-replica loss then .mean() (main.py:704,719-720)
    out["losses"] = losses.cpu().numpy()
    out["bn_mean"] = G.blk6.norm2.norm.running_mean.cpu().numpy()
    out["bn_var"] = G.blk3a.norm1.norm.running_var.cpu().numpy()
    out["sn_u"] = D.d1.conv3.weight_u.cpu().numpy()
    out["w_probe"] = G.blk5.conv1.weight_orig.detach()[:4, :4].cpu().numpy()          # after Adam
    out["fc_probe"] = G.fc.weight.detach()[:8, :8].cpu().numpy()
    if rank == 0:
        np.savez(out_path, **out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
