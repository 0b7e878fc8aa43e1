"""Generate tests/golden/gan_reference.npz from the REFERENCE's own modules (authoring container only):
    python tests/golden/make_golden_gan.py
Imports models/gan.py and utils/losses.py unmodified from /root/reference/code (they run on CPU),
builds G and D with the seeds of tests/golden/gan_common.py, runs one generator step and one discriminator step
(ModelWrapper.forward modes 'g' and 'd', main.py:476-521) in training mode and stores output probes, losses and
per-parameter gradient norms.  The CUDA modules are built with the same seeds on the GPU box (their initial
values equal the reference's — checked by tests/test_gan_hostlogic.py when /root/reference is present)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/code")
import gan_common as GC                                     # noqa: E402
from models import gan as ref_gan                           # noqa: E402  (reference)
from utils.losses import GANLoss                            # noqa: E402  (reference)


def main():
    torch.set_num_threads(8)
    args = GC.make_args(256, 2)
    G, D = GC.build(ref_gan, args)
    G.train(); D.train()
    crit = GANLoss('hinge', tensor=torch.FloatTensor)
    z, c, alpha, tex, mesh = GC.inputs(args)
    out = {}
    # ---- generator step
    loss, pred_tex, pred_mesh, dout, mask = GC.g_step(G, D, crit, z, c, alpha)
    loss.mean().backward()
    out["g_loss"] = loss.detach().numpy()
    out["tex_probe"] = pred_tex.detach()[:, :, ::16, ::16].numpy()
    out["tex_sum"] = np.float64(pred_tex.detach().double().sum())
    out["mesh"] = pred_mesh.detach().numpy()
    out["d_out0"], out["d_out1"] = dout[0].detach().numpy(), dout[1].detach().numpy()
    out["mask0"], out["mask1"] = mask[0].numpy(), mask[1].numpy()
    names = [n for n, p in G.named_parameters() if p.grad is not None]
    out["g_grad_names"] = np.array(names)
    out["g_grad_norms"] = np.array([float(dict(G.named_parameters())[n].grad.norm()) for n in names])
    out["g_grad_probe"] = G.blk5.conv1.weight_orig.grad[:4, :4].numpy()
    out["sn_u_blk1"] = G.blk1.conv1.weight_u.numpy().copy()
    out["bn_mean_blk6"] = G.blk6.norm2.norm.running_mean.numpy().copy()
    G.zero_grad(); D.zero_grad()
    # ---- discriminator step (same modules, second forward: spectral-norm u/v advance again, as in training)
    lf, lr, dout = GC.d_step(G, D, crit, z, c, alpha, tex, mesh)
    (lf.mean() + lr.mean()).backward()
    out["d_loss_fake"], out["d_loss_real"] = lf.detach().numpy(), lr.detach().numpy()
    out["dd_out0"] = dout[0].detach().numpy()
    names = [n for n, p in D.named_parameters() if p.grad is not None]
    out["d_grad_names"] = np.array(names)
    out["d_grad_norms"] = np.array([float(dict(D.named_parameters())[n].grad.norm()) for n in names])
    path = os.path.join(HERE, "gan_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; g_loss", out["g_loss"], "d losses", out["d_loss_fake"], out["d_loss_real"])


if __name__ == "__main__":
    main()
