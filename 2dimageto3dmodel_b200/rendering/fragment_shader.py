"""Drop-in for /root/reference/code/rendering/fragment_shader.py.

On the render path the shader is fused into the rasteriser kernel (csrc/mesh_kernels.cu,
mesh_raster_fwd_kernel<SHADE=true>); these two functions keep the stand-alone call surface for callers
that shade pre-computed UV images (same maths: uv -> [-1,1], v flipped, bilinear, corners aligned)."""
import torch

from .utils import grid_sample_bilinear


def texinterpolation(imtexcoord_bxhxwx2, texture_bx3xthxtw, filtering='bilinear'):
    g = imtexcoord_bxhxwx2 * 2 - 1
    g = torch.stack((g[..., 0], -g[..., 1]), dim=-1)
    if filtering == 'bilinear':
        col = grid_sample_bilinear(texture_bx3xthxtw, g)
    else:
        col = torch.nn.functional.grid_sample(texture_bx3xthxtw, g, mode=filtering)
    return col.permute(0, 2, 3, 1)


def fragmentshader(imtexcoord_bxhxwx2, texture_bx3xthxtw, improb_bxhxwx1, filtering='bilinear',
                   background_image=None):
    col = texinterpolation(imtexcoord_bxhxwx2, texture_bx3xthxtw, filtering=filtering)
    if background_image is None:
        return col * improb_bxhxwx1
    return torch.lerp(background_image, col, improb_bxhxwx1)
