"""Drop-in for /root/reference/code/utils/trilinear_interpolation.py (class TrilinearInterpolation, :11-74):
8-corner trilinear scatter of camera-space points into a [B,V,V,V] occupancy grid, clamped to [0,1], on libb3d
(one atomic-scatter kernel + clamp instead of eight dense index_put_ grids and a stack/sum).
The differentiable route is EffectiveLossFunction (the fused kernels own the adjoint); this class is the
forward-only stand-alone surface."""
import torch

from b3d import check, dev, lib, mode_id, ptr, stream_ptr


class TrilinearInterpolation(object):
    def __init__(self, epsilon=1e-6, size=64, semantics="R"):
        self.epsilon, self.size, self.semantics = epsilon, size, semantics
        mode_id(semantics)

    def get_point_cloud_object_borders(self, point_cloud):
        inside = (point_cloud < 0.5 - self.epsilon) & (point_cloud > -0.5 + self.epsilon)
        return inside.all(dim=-1).view(-1)

    def get_grid(self, point_cloud, voxel_size):
        return (voxel_size - 1) * (point_cloud + 0.5)

    def trilinear_interpolation(self, point_cloud):
        """point_cloud [B,N,3] = camera coordinates (z, y, x) -> occupancy [B,V,V,V] (index order b, z, y, x)."""
        c = dev(point_cloud.detach(), "point_cloud")
        B, N, _ = c.shape
        g = self.get_grid(c, float(self.size))
        flag = self.get_point_cloud_object_borders(c).view(B, N, 1).to(c.dtype)
        pg = torch.cat((g, flag), dim=2).contiguous()
        grid = torch.empty(B, self.size, self.size, self.size, device=c.device, dtype=torch.float32)
        check(lib.b3d_pc_splat_grid(ptr(pg), B, N, self.size, mode_id(self.semantics), ptr(grid), stream_ptr(c)))
        return grid
