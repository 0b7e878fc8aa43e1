"""Generate tests/golden/renderer_reference.npz by running the REFERENCE's rendering/renderer.py:Renderer.forward,
unmodified, on the CPU (authoring container only):
    python tests/golden/make_golden_renderer.py
renderer.py imports two kaolin functions (renderer.py:1-2).  kaolin's CUDA rasteriser cannot exist here, so
the stand-in module hands the class the ORACLE's restatement of `linear_rasterizer` (oracle/mesh.py:rasterize, parity
unpinned) and of `datanormalize`.  What this pins is therefore everything AROUND the rasteriser, executed by the reference's
own code: ortho_projection (:9-28), the (height, width) argument order, the (u, v, 1) attribute packing, the use of
imfeat[..., :2] / [..., 2:3] as texture coordinates / hard mask, the fragment shader call with and without a background,
return_hardmask, and the returned unit normals — against which oracle/mesh.py:render (the function the CUDA renderer is
compared with) must agree exactly.
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

from oracle import mesh as M                      # noqa: E402

calls = []


def linear_rasterizer(width, height, points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, vertex_attr_bxfx3d):
    # the reference passes (self.height, self.width) into (width, height) — SURVEY App. A D13; square images only
    calls.append((width, height))
    imfeat, improb, _, _ = M.rasterize(points3d_bxfx9, points2d_bxfx6, normalz_bxfx1, vertex_attr_bxfx3d, height, width)
    return imfeat, improb


mods = {n: types.ModuleType(n) for n in ("kaolin", "kaolin.graphics", "kaolin.graphics.dib_renderer",
                                         "kaolin.graphics.dib_renderer.rasterizer", "kaolin.graphics.dib_renderer.utils")}
mods["kaolin.graphics.dib_renderer.rasterizer"].linear_rasterizer = linear_rasterizer
mods["kaolin.graphics.dib_renderer.utils"].datanormalize = M.datanormalize
sys.modules.update(mods)

from rendering.renderer import Renderer, ortho_projection       # noqa: E402  (reference, unmodified)


def main():
    out = {}
    path = M.write_uvsphere_obj(os.path.join(tempfile.mkdtemp(), "uvsphere_16rings.obj"), rings=16)
    T = M.TemplateData(M.load_obj(path), path)
    g = torch.Generator().manual_seed(29)
    B, H = 2, 48
    mesh_map = torch.randn(B, 3, 32, 32, generator=g) * 0.05
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    s, t = 0.5 + 0.3 * torch.rand(B, 1, generator=g), (torch.rand(B, 3, generator=g) - 0.5) * 0.3
    tex = torch.rand(B, 3, 16, 16, generator=g) * 2 - 1
    bg = torch.rand(B, H, H, 3, generator=g)
    vtx = M.transform_vertices(M.get_vertex_positions(T, mesh_map), s, t, q)
    uvs, padded = M.adjust_uv_and_texture(T, tex)
    r = Renderer(H, H)
    img, alpha, n1 = r([vtx, T.faces], uvs, padded, ft_fx3=T.face_textures)
    img_bg, hard, _ = r([vtx, T.faces], uvs, padded, ft_fx3=T.face_textures, background_image=bg, return_hardmask=True)
    img_noft, _, _ = r([vtx, T.faces[:, [0, 1, 2]]], uvs[:, :T.vertices.shape[0]], padded)       # ft_fx3=None -> faces index the uvs
    p3d, p2d, nrm = ortho_projection(vtx, T.faces)
    out.update(mesh_map=mesh_map.numpy(), q=q.numpy(), s=s.numpy(), t=t.numpy(), tex=tex.numpy(), bg=bg.numpy(), H=np.array([H]),
               img=img.numpy(), alpha=alpha.numpy(), normal1=n1.numpy(), img_bg=img_bg.numpy(), hard=hard.numpy(),
               img_noft=img_noft.numpy(), p3d=p3d.numpy(), p2d=p2d.numpy(), normal=nrm.numpy())
    assert all(c == (H, H) for c in calls) and len(calls) == 3
    p = os.path.join(HERE, "renderer_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes; coverage", float(hard.mean()))


if __name__ == "__main__":
    main()
