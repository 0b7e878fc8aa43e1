"""Generate tests/golden/mesh_*.npz from the REFERENCE's own Python (authoring container only):
    python tests/golden/make_golden_mesh.py
Imports, unmodified, from /root/reference/code: rendering/utils.py (qrot, circpad, symmetrize_texture,
adjust_poles, qmul), rendering/fragment_shader.py (fragmentshader), rendering/monkey_patches.py
(compute_adjacency_info_patched -> ff), utils/losses.py (loss_flat, GANLoss).
mesh_template.py / renderer.py need kaolin and cannot be imported.
The mesh used is the procedural UV sphere of oracle/mesh.py (travels to the GPU box, unlike the OBJ files).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/code")

from rendering.utils import qrot, qmul, circpad, symmetrize_texture, adjust_poles   # noqa: E402  (reference)
from rendering.fragment_shader import fragmentshader                                  # noqa: E402  (reference)
from rendering.monkey_patches import compute_adjacency_info_patched                   # noqa: E402  (reference)
from utils.losses import loss_flat, GANLoss                                           # noqa: E402  (reference)
from oracle import mesh as M                                                          # noqa: E402


def main():
    g = torch.Generator().manual_seed(21)
    out = {}
    q = torch.nn.functional.normalize(torch.randn(3, 4, generator=g), dim=-1)
    v = torch.randn(3, 7, 3, generator=g)
    out["qrot_q"], out["qrot_v"], out["qrot_out"] = q.numpy(), v.numpy(), qrot(q, v).numpy()
    r = torch.randn(3, 4, generator=g)
    out["qmul_r"], out["qmul_out"] = r.numpy(), qmul(q, r).numpy()

    tex = torch.rand(2, 3, 6, 10, generator=g) * 2 - 1
    out["tex"] = tex.numpy()
    out["circpad2"] = circpad(tex, 2).numpy()
    out["symmetrize"] = symmetrize_texture(tex).numpy()
    out["adjust_poles"] = adjust_poles(tex).numpy()
    uv = torch.rand(2, 8, 8, 2, generator=g)
    mask = (torch.rand(2, 8, 8, 1, generator=g) > 0.4).float()
    bg = torch.rand(2, 8, 8, 3, generator=g)
    out["fs_uv"], out["fs_mask"], out["fs_bg"] = uv.numpy(), mask.numpy(), bg.numpy()
    out["fs_out"] = fragmentshader(uv, tex, mask).numpy()
    out["fs_out_bg"] = fragmentshader(uv, tex, mask, background_image=bg).numpy()

    tmp = tempfile.mkdtemp()
    mesh = M.load_obj(M.write_uvsphere_obj(os.path.join(tmp, "uvsphere_16rings.obj"), rings=16))
    ff = compute_adjacency_info_patched(mesh["vertices"], mesh["faces"])[8]
    out["ff16"] = ff.numpy().astype(np.int16)
    norms = torch.nn.functional.normalize(torch.randn(2, 960, 3, generator=g), dim=-1)
    holder = types.SimpleNamespace(ff=ff, faces=mesh["faces"])
    out["flat_norms"], out["flat_loss"] = norms.numpy(), loss_flat(holder, norms).numpy()

    # GANLoss (hinge, masked, weighted) on list inputs — utils/losses.py:100-120
    crit = GANLoss("hinge", tensor=torch.FloatTensor)
    preds = [torch.randn(4, 1, 8, 8, generator=g), torch.randn(4, 1, 4, 4, generator=g)]
    masks = [torch.rand(4, 1, 8, 8, generator=g), torch.rand(4, 1, 4, 4, generator=g)]
    out["gl_p0"], out["gl_p1"], out["gl_m0"], out["gl_m1"] = [t.numpy() for t in preds + masks]
    out["gl_g"] = crit(preds, True, for_discriminator=False, mask=masks, weight=[2, 1]).numpy()
    out["gl_d_fake"] = crit(preds, False, for_discriminator=True, mask=masks, weight=None).numpy()
    out["gl_d_real"] = crit(preds, True, for_discriminator=True, mask=masks, weight=[2, 1]).numpy()
    out["gl_g_nomask"] = crit(preds, True, for_discriminator=False).numpy()

    path = os.path.join(HERE, "mesh_reference_pieces.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
