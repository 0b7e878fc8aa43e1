"""oracle/gan.py reproduces the golden vectors the reference's own modules produced (CPU, travels to any box):
weights are re-created from the seeds (module construction = the drop-in's, whose initial values equal the
reference's — tests/test_gan_hostlogic.py), then only oracle code runs."""
import os
import sys

import numpy as np
import torch

from conftest import GOLDEN
from oracle import gan as OG

sys.path.insert(0, GOLDEN)
import gan_common as GC          # noqa: E402


def test_oracle_reproduces_reference_golden():
    from models import gan
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    d = np.load(os.path.join(GOLDEN, "gan_reference.npz"))
    args = GC.make_args(256, 2)
    G, D = GC.build(gan, args)
    sg = {k: v.clone() for k, v in G.state_dict().items()}
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    for s in (sg, sd):
        for k in OG.trainable(s):
            s[k].requires_grad_(True)
    z, c, alpha, tex, mesh = GC.inputs(args)
    loss, pred_tex, pred_mesh, out, mask = OG.g_loss(sg, sd, args, z, c, alpha)
    np.testing.assert_allclose(pred_tex.detach()[:, :, ::16, ::16].numpy(), d["tex_probe"], atol=2e-5)
    np.testing.assert_allclose(pred_mesh.detach().numpy(), d["mesh"], atol=2e-6)
    np.testing.assert_allclose(out[0].detach().numpy(), d["d_out0"], atol=2e-5)
    np.testing.assert_allclose(out[1].detach().numpy(), d["d_out1"], atol=2e-5)
    np.testing.assert_allclose(mask[0].numpy(), d["mask0"], atol=1e-7)
    assert abs(float(loss) - float(d["g_loss"][0])) < 1e-5
    names = [str(n) for n in d["g_grad_names"]]
    grads = torch.autograd.grad(loss, [sg[n] for n in names], allow_unused=True)
    for n, g, ref in zip(names, grads, d["g_grad_norms"]):
        assert abs(float(g.norm()) - ref) <= 1e-3 * ref + 1e-7, (n, float(g.norm()), ref)
    np.testing.assert_allclose(sg["blk1.conv1.weight_u"].detach().numpy(), d["sn_u_blk1"], atol=1e-6)
    np.testing.assert_allclose(sg["blk6.norm2.norm.running_mean"].numpy(), d["bn_mean_blk6"], atol=1e-6)
    lf, lr, dout = OG.d_loss(sg, sd, args, z, c, alpha, tex, mesh)
    assert abs(float(lf) - float(d["d_loss_fake"][0])) < 1e-5 and abs(float(lr) - float(d["d_loss_real"][0])) < 1e-5
    np.testing.assert_allclose(dout[0].detach().numpy(), d["dd_out0"], atol=2e-5)
    names = [str(n) for n in d["d_grad_names"]]
    grads = torch.autograd.grad(lf + lr, [sd[n] for n in names])
    for n, g, ref in zip(names, grads, d["d_grad_norms"]):
        assert abs(float(g.norm()) - ref) <= 1e-3 * ref + 1e-7, (n, float(g.norm()), ref)
