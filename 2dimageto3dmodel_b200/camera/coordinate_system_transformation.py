"""Drop-in for /root/reference/code/camera/coordinate_system_transformation.py (CameraUtilities, :16-39): rotate the
cloud by the (normalised) quaternion and apply the pin-hole perspective factor fov / (z + d) to x and y, on libb3d's
projection kernel (csrc/pc_kernels.cu: pc_project_kernel / pc_project_bwd_kernel)."""
import torch

from b3d import check, dev, lib, ptr, stream_ptr


class _ToCamera(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, quat, fov, dist):
        p, q = dev(points.detach(), "point_cloud"), dev(quat.detach(), "rotation")
        B, N, _ = p.shape
        pg = torch.empty(B, N, 4, device=p.device, dtype=torch.float32)
        coords = torch.empty(B, N, 3, device=p.device, dtype=torch.float32)
        check(lib.b3d_pc_project(ptr(p), ptr(q), B, N, 2, fov, dist, ptr(pg), ptr(coords), None, None, None, None, stream_ptr(p)))
        ctx.save_for_backward(p, q)
        ctx.cfg = (fov, dist)
        return coords

    @staticmethod
    def backward(ctx, g):
        p, q = ctx.saved_tensors
        fov, dist = ctx.cfg
        B, N, _ = p.shape
        # V = 2 makes grid coordinates = camera coordinates + 0.5, so d/d(grid) = d/d(coords); every point takes part
        dpg = torch.cat((dev(g, "grad"), torch.zeros(B, N, 1, device=p.device)), dim=2).contiguous()
        live = torch.ones(B, N, 4, device=p.device)
        dp, dq = torch.empty_like(p), torch.empty_like(q)
        check(lib.b3d_pc_project_bwd(ptr(p), ptr(q), ptr(live), ptr(dpg), B, N, 2, fov, dist, ptr(dp), ptr(dq), stream_ptr(p)))
        return dp, dq, None, None


class CameraUtilities(object):
    def transformation_3d_coord_to_camera_coord(self, point_cloud, rotation, field_of_view, camera_view_distance):
        """point_cloud [B,N,3] (columns z, y, x), rotation [B,4] (w, x, y, z) -> camera coordinates [B,N,3]."""
        return _ToCamera.apply(point_cloud, rotation, float(field_of_view), float(camera_view_distance))
