"""§8f rank 3: InverseRenderer + the texel-visibility mask on the CUDA rasteriser against the oracle's restatement of the
reference's construction (run_reconstruction.py:506-527, :571-585), and a 1024^2 render (the export's resolution,
run_reconstruction.py:84) against the oracle on one sample.  Visibility / index buffers: bit exact."""
import os
import tempfile

import pytest
import torch

from oracle import mesh as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def tpl():
    from rendering.mesh_template import MeshTemplate
    path = M.write_uvsphere_obj(os.path.join(tempfile.mkdtemp(), "uvsphere_16rings.obj"), rings=16)
    return MeshTemplate(path, device=DEV), M.TemplateData(M.load_obj(path), path)


def scene(B, seed, tex_res):
    g = torch.Generator().manual_seed(seed)
    mesh_map = torch.randn(B, 3, 32, 32, generator=g) * 0.05
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    s = 0.5 + 0.3 * torch.rand(B, 1, generator=g)
    t = (torch.rand(B, 3, generator=g) - 0.5) * 0.3
    tex = torch.rand(B, 3, tex_res, tex_res, generator=g) * 2 - 1
    return mesh_map, q, s, t, tex


def test_texel_visibility_mask_matches_oracle(tpl):
    from data.pseudo_gt import visibility_to_mask
    from rendering.inverse_renderer import texel_visibility
    from rendering.renderer import Renderer
    mt, T = tpl
    B, H, R = 2, 256, 64
    mesh_map, q, s, t, tex = scene(B, 3, 32)
    vtx = M.transform_vertices(M.get_vertex_positions(T, mesh_map), s, t, q)
    to = tex.clone().requires_grad_(True)
    img_o, _, idx_o = M.forward_renderer(T, vtx, to, H, H)
    vis_o, = torch.autograd.grad(img_o, to, torch.ones_like(img_o))
    r = Renderer(H, H)
    vis, img, alpha = texel_visibility(mt, r, vtx.to(DEV), tex.to(DEV))
    assert torch.equal(r.last_face_index.cpu(), idx_o)
    assert float((vis.cpu() - vis_o).abs().max()) <= 1e-4 * float(vis_o.abs().max())
    assert torch.equal(vis.cpu() > 0, vis_o > 0), "texel-visibility mask differs from the reference construction"
    m, mo = visibility_to_mask(vis, R).cpu(), visibility_to_mask(vis_o, R)
    assert torch.equal(m, mo)
    assert 0.1 < float(m.mean()) < 0.9


def test_inverse_renderer_matches_oracle(tpl):
    from rendering.inverse_renderer import InverseRenderer
    mt, T = tpl
    B, R = 2, 64
    mesh_map, q, s, t, _ = scene(B, 5, 32)
    vtx = M.transform_vertices(M.get_vertex_positions(T, mesh_map), s, t, q)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(B, 4, 96, 96, generator=g) * 2 - 1                 # RGBA photograph
    # oracle: the same UV-space render through the oracle's Renderer restatement, three channels at a time
    uvs = (vtx[..., :2] + 1) / 2
    verts = torch.cat((T.uvs.unsqueeze(0) * 2 - 1, torch.zeros(1, T.uvs.shape[0], 1)), dim=-1).expand(B, -1, -1)
    outs = []
    for idx in ([0, 1, 2], [3, 3, 3]):
        img_o, hard_o, _, _ = M.render(verts, T.face_textures, uvs, target[:, idx], ft=T.faces, H=R, W=R, return_hardmask=True)
        outs.append(img_o)
    ref = torch.cat((outs[0], outs[1][..., :1]), dim=3)
    inv = InverseRenderer(mt.mesh, R, R)
    img, hard = inv(vtx.to(DEV), target.to(DEV))
    assert img.shape == (B, R, R, 4) and hard.shape == (B, R, R, 1)
    assert torch.equal(hard.cpu() > 0.5, hard_o > 0.5)
    assert float((img.cpu() - ref).abs().max()) < 1e-4      # 96-texel "texture": fp32 uv rounding x T |d tex| (values in [-1, 1])


def test_render_1024_matches_oracle_on_one_sample(tpl):
    """renderer_res = max(1024, 2 * pseudogt_resolution) (run_reconstruction.py:84): 4096 tiles per sample."""
    from rendering.renderer import Renderer
    mt, T = tpl
    B, H = 2, 1024
    mesh_map, q, s, t, tex = scene(B, 7, 128)
    vtx = M.transform_vertices(M.get_vertex_positions(T, mesh_map), s, t, q)
    r = Renderer(H, H)
    img, alpha = mt.forward_renderer(r, vtx.to(DEV), tex.to(DEV))
    img_o, alpha_o, idx_o = M.forward_renderer(T, vtx[1:2], tex[1:2], H, H)
    nbad = int((r.last_face_index[1].cpu() != idx_o[0]).sum())
    assert nbad == 0, f"face-index buffer differs from the oracle in {nbad} of {idx_o.numel()} pixels"
    assert float((img[1].cpu() - img_o[0]).abs().max()) < 1e-4
    assert float((alpha[1].cpu() - alpha_o[0]).abs().max()) < 2e-5
