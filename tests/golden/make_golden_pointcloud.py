"""Generate tests/golden/pointcloud_*.npz by running the REFERENCE's own classes.

Run in the authoring container only (needs /root/reference, which does not exist on
the GPU box):      python tests/golden/make_golden_pointcloud.py

Imports /root/reference/code through the shim recipe of SURVEY.md §8c and applies, at
run time and without touching the reference tree, only the *execution* patches of
SURVEY.md App. A (numerics untouched = oracle mode "R"):
  P1  drop `assert not len(xyz_triplet) != 3`   (points_quaternions.py:23, tests the batch dim)
  P2  kernels=() -> separate_kernels(sigma, k)  (effective_loss_function.py:77)
  P3  TrilinearInterpolation(size=voxel_size)   (effective_loss_function.py:72)
Outputs: inputs, silhouette, d(sum(sil*wts))/d(points, rotation, scale), the int64
index buffer of corner (0,0,0) and the in-bounds mask, all from the reference code.
"""
import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [REF]
    sys.modules["refpkg"] = pkg
    for sub in ("utils", "models"):
        sys.path.insert(0, os.path.join(REF, sub))
    elf = importlib.import_module("refpkg.utils.effective_loss_function")
    pq = importlib.import_module("refpkg.quaternions.points_quaternions")
    tri = importlib.import_module("trilinear_interpolation")
    sv = importlib.import_module("smooth_voxels")

    # P1
    pq.PointsQuaternionsConverter.points_to_quaternions = staticmethod(
        lambda t: torch.nn.functional.pad(input=t, pad=(1, 0, 0, 0)))
    return elf, tri, sv


def run_reference(elf, tri, sv, points, q, scale, V, ksize, sigma):
    mod = elf.EffectiveLossFunction(voxel_size=V, kernel_size=ksize, smooth_sigma=sigma)

    # P3: the forward constructs TrilinearInterpolation() with no arguments
    orig_tri_init = tri.TrilinearInterpolation.__init__

    def tri_init(self, epsilon=1e-6, size=V):
        orig_tri_init(self, epsilon=epsilon, size=size)

    # P2: the forward passes kernels=()
    orig_smooth = sv.VoxelsSmooth.smooth

    def smooth(self, voxels, kernels, scale=None):
        if len(kernels) == 0:
            kernels = self.separate_kernels(mod.sigma, mod.kernel_size)
        return orig_smooth(self, voxels, kernels, scale)

    elf.TrilinearInterpolation.__init__ = tri_init
    elf.VoxelsSmooth.smooth = smooth
    try:
        with contextlib.redirect_stdout(io.StringIO()):      # the classes print on construction
            sil = mod(points, q, scale)
            # index buffer of corner 0 exactly as positions_update builds it (:47-52)
            cu = elf.CameraUtilities()
            c = cu.transformation_3d_coord_to_camera_coord(point_cloud=points, rotation=q,
                                                           field_of_view=1.875, camera_view_distance=2.0)
            t = tri.TrilinearInterpolation(size=V)
            inb = t.get_point_cloud_object_borders(c)
            base = t.get_grid(point_cloud=c, voxel_size=c.new(3).fill_(V)).floor().long()
            occ = t.trilinear_interpolation(point_cloud=c)
    finally:
        elf.TrilinearInterpolation.__init__ = orig_tri_init
        elf.VoxelsSmooth.smooth = orig_smooth
    return sil, c, base, inb, occ


def make(name, B, N, V, ksize, sigma, with_scale, seed, spread=0.45):
    elf, tri, sv = import_reference()
    g = torch.Generator().manual_seed(seed)
    points = ((torch.rand(B, N, 3, generator=g) * 2 - 1) * spread).requires_grad_(True)
    q = torch.randn(B, 4, generator=g).requires_grad_(True)
    scale = (0.5 + 0.5 * torch.rand(B, 1, generator=g)).requires_grad_(True) if with_scale else None
    wts = torch.rand(B, V, V, generator=g)
    sil, c, base, inb, occ = run_reference(elf, tri, sv, points, q, scale, V, ksize, sigma)
    loss = (sil * wts).sum()
    grads = torch.autograd.grad(loss, [points, q] + ([scale] if with_scale else []))
    out = dict(points=points.detach().numpy(), q=q.detach().numpy(), wts=wts.numpy(),
               V=np.int64(V), ksize=np.int64(ksize), sigma=np.float64(sigma),
               sil=sil.detach().numpy(), coords=c.detach().numpy(),
               base=base.numpy().astype(np.int16), inb=inb.view(B, N).numpy(),
               occ_sum=occ.detach().double().sum().numpy(),
               occ_probe=occ.detach()[:, ::7, ::5, ::3].numpy(),
               d_points=grads[0].numpy(), d_q=grads[1].numpy())
    if with_scale:
        out["scale"] = scale.detach().numpy()
        out["d_scale"] = grads[2].numpy()
    path = os.path.join(HERE, f"pointcloud_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "sil range", float(sil.min()), float(sil.max()), "inb", int(inb.sum()), "/", B * N,
          os.path.getsize(path), "bytes")


if __name__ == "__main__":
    torch.set_num_threads(4)
    # small, V=32 (exercises P3), with and without scale
    make("v32_scale", B=2, N=300, V=32, ksize=21, sigma=3.0, with_scale=True, seed=11)
    make("v32_noscale", B=3, N=200, V=32, ksize=21, sigma=1.0, with_scale=False, seed=12)
    # the reference default geometry (V=64), BASELINE config-1 shape cut to B=2 to keep it small
    make("v64_cfg1", B=2, N=1024, V=64, ksize=21, sigma=3.0, with_scale=True, seed=13)
    # points outside the frustum: spread 0.7 puts many points out of bounds
    make("v32_oob", B=2, N=256, V=32, ksize=21, sigma=1.5, with_scale=True, seed=14, spread=0.7)
