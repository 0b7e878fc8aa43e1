"""Generate tests/golden/recon_reference.npz from the REFERENCE's own models/reconstruction.py (authoring container only):
    python tests/golden/make_golden_recon.py
The reference module runs on the CPU unmodified.  One training-mode forward + backward of
ReconstructionNetwork(symmetric=True, texture_res=128) on a seeded batch of 8 RGBA images, and DatasetParams lookups."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/code")
import recon_common as RC                                   # noqa: E402
from models import reconstruction as ref                    # noqa: E402  (reference)


def main():
    torch.set_num_threads(8)
    net = RC.build(ref).train()
    x, w_tex, w_mesh = RC.inputs()
    tex, mesh_map = net(x)
    RC.loss_of(tex, mesh_map, w_tex, w_mesh).backward()
    out = {"tex_probe": tex.detach()[:, :, ::8, ::8].numpy(), "tex_sum": np.float64(tex.detach().double().sum()),
           "mesh_map": mesh_map.detach().numpy()}
    names = [n for n, p in net.named_parameters() if p.grad is not None]
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([float(dict(net.named_parameters())[n].grad.norm()) for n in names])
    out["grad_probe"] = net.conv3e.weight.grad[:4, :4].numpy()
    out["bn4e_mean"] = net.bn4e.running_mean.numpy().copy()
    out["bn_blk2_var"] = net.blk2.bn2.running_var.numpy().copy()
    # DatasetParams (pure index arithmetic + exp)
    dp = ref.DatasetParams(RC.dataset_args(), 10)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        dp.ds_translation.copy_(torch.randn(10, 2, generator=g) * 0.1)
        dp.ds_scale.copy_(torch.randn(10, 1, generator=g) * 0.1)
        dp.ds_z0.copy_(torch.randn(10, 1, generator=g))
    out["dp_state"] = np.concatenate([dp.ds_translation.detach().numpy(), dp.ds_scale.detach().numpy(), dp.ds_z0.detach().numpy()], 1)
    idx = torch.tensor([0, 3, 12, 19, 7])
    t, s = dp(idx, 'deltas')
    out["dp_idx"], out["dp_t"], out["dp_s"] = idx.numpy(), t.detach().numpy(), s.detach().numpy()
    out["dp_z0"] = dp(idx, 'z0').detach().numpy()
    t, s = dp(None, 'deltas')
    out["dp_t_mean"], out["dp_s_mean"], out["dp_z0_mean"] = t.detach().numpy(), s.detach().numpy(), dp(None, 'z0').detach().numpy()
    path = os.path.join(HERE, "recon_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; tex_sum", out["tex_sum"], "grad norm range", out["grad_norms"].min(), out["grad_norms"].max())


if __name__ == "__main__":
    main()
