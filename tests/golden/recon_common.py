"""Shared recipe of the reconstruction-network golden test: identical construction (same seed -> same initial weights,
verified equal to the reference's in the authoring container) and identical synthetic inputs for the reference (CPU) and
the CUDA module."""
import types

import torch


def build(recon_module, seed=321, texture_res=128):
    torch.manual_seed(seed)
    net = recon_module.ReconstructionNetwork(symmetric=True, texture_res=texture_res, mesh_res=32)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():       # the mesh head is zero-initialised: give it a signal so the test sees it
        net.conv_mesh.weight.copy_(torch.randn(net.conv_mesh.weight.shape, generator=g) * 0.02)
        net.conv_mesh.bias.copy_(torch.randn(net.conv_mesh.bias.shape, generator=g) * 0.02)
    return net


def inputs(B=8, seed=11, texture_res=128):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 4, 256, 256, generator=g) * 2 - 1
    w_tex = torch.randn(B, 3, texture_res, texture_res, generator=g)
    w_mesh = torch.randn(B, 3, 32, 32, generator=g)
    return x, w_tex, w_mesh


def loss_of(tex, mesh_map, w_tex, w_mesh):
    return (tex * w_tex).mean() + 10.0 * (mesh_map * w_mesh).mean()


def dataset_args():
    return types.SimpleNamespace(optimize_deltas=True, optimize_z0=True)
