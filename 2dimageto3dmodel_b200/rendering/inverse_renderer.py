"""The pseudo-ground-truth projection of /root/reference/code/run_reconstruction.py as importable pieces (the reference
defines them inside the `--generate_pseudogt` branch of the script, :506-527 and :542-585):

  InverseRenderer          renders IN UV SPACE: the template's uv coordinates are the "vertices", the predicted camera-space
                           vertex positions are the "uvs" and the photograph is the "texture" — every texel receives the
                           colour of the image point its surface point projects to (:506-527).
  texel_visibility         d(render) / d(texture) with an all-ones upstream gradient: > 0 exactly for the texels some
                           covered pixel samples (:571-572); `visibility_to_mask` (data/pseudo_gt.py) turns it into the
                           mask that gates the projected texture (:581-585).

Both run on the libb3d rasteriser (tile-binned DIB-R + fused shader, csrc/mesh_kernels.cu); the face-index buffer of either
render is available as `renderer.last_face_index`."""
import torch
import torch.nn as nn

from .renderer import Renderer


class InverseRenderer(nn.Module):
    def __init__(self, mesh, res_h, res_w):
        super().__init__()
        self.res = (res_h, res_w)
        self.inverse_renderer = Renderer(res_h, res_w)
        self.mesh = mesh

    def forward(self, predicted_vertices, target):
        """predicted_vertices [B,V,3] camera space, target [B,C,H,W] image -> (projected [B,R,R,C], hard mask [B,R,R,1])."""
        with torch.no_grad():
            B = target.shape[0]
            uvs = ((predicted_vertices[..., :2] + 1) / 2).contiguous()
            vertices = self.mesh.uvs.unsqueeze(0) * 2 - 1
            vertices = torch.cat((vertices, torch.zeros_like(vertices[..., :1])), dim=-1).expand(B, -1, -1).contiguous()
            outs, alpha = [], None
            C = target.shape[1]
            for c0 in range(0, C, 3):                    # the shader kernels take RGB textures: three channels per pass
                idx = [min(c0 + k, C - 1) for k in range(3)]
                tex = target[:, idx].contiguous()
                img, alpha, _ = self.inverse_renderer(points=[vertices, self.mesh.face_textures], uv_bxpx2=uvs,
                                                      texture_bx3xthxtw=tex, ft_fx3=self.mesh.faces, return_hardmask=True)
                outs.append(img[..., :min(3, C - c0)])
            return (outs[0] if len(outs) == 1 else torch.cat(outs, dim=3)), alpha


def texel_visibility(mesh_template, renderer, vtx, pred_tex):
    """-> (visibility [B,3,Th,Tw] = autograd.grad(image_pred, pred_tex, ones), image_pred, alpha_pred).  The mask the
    reference derives from it is `visibility > 0` after a bilinear resize (data.pseudo_gt.visibility_to_mask)."""
    tex = pred_tex.detach().requires_grad_(True)
    image_pred, alpha_pred = mesh_template.forward_renderer(renderer, vtx.detach(), tex)
    visibility, = torch.autograd.grad(image_pred, tex, torch.ones_like(image_pred))
    return visibility, image_pred.detach(), alpha_pred.detach()
