"""Point-cloud effective-loss ops over libb3d (autograd wrappers; plumbing only).

Reference path: /root/reference/code/utils/effective_loss_function.py:58-81 and the modules it
calls (camera/, quaternions/, utils/trilinear_interpolation.py, utils/smooth_voxels.py).
"""
import ctypes

import torch

from . import B3DError, check, dev, host_floats, lib, mode_id, ptr, stream_ptr

FIELD_OF_VIEW = 1.875          # effective_loss_function.py:69
CAMERA_VIEW_DISTANCE = 2.0     # effective_loss_function.py:70


def smoothing_taps(sigma, kernel_size=21, mode="R"):
    """The 1-D kernel of VoxelsSmooth.separate_kernels (smooth_voxels.py:24-31), on the CPU in fp32
    with the reference's own expression (mode R keeps its positive exponent)."""
    a, b = (-kernel_size // 2, kernel_size // 2)
    x = torch.arange(a + 1.0, b + 1.0)
    s = torch.as_tensor(float(sigma), dtype=torch.float32)
    if mode_id(mode) == 0:
        k = torch.exp(pow(-x, 2) / (2 * pow(s, 2)))
    else:
        k = torch.exp(-pow(x, 2) / (2 * pow(s, 2)))
    k = k / k.sum()
    return [float(v) for v in k]


def project(points, quat, V, fov=FIELD_OF_VIEW, cam_dist=CAMERA_VIEW_DISTANCE, want_aux=False, want_bins=False):
    """-> pg [B,N,4] (+ coords [B,N,3], base int32 [B,N,3], inb uint8 [B,N] when want_aux)
    (+ sorted [B,N,4], bin_start [B,nbins+1] when want_bins)."""
    points = dev(points, "point_cloud")
    quat = dev(quat, "rotation")
    B, N, three = points.shape
    if three != 3 or quat.shape != (B, 4):
        raise B3DError(f"bad shapes: point_cloud {tuple(points.shape)}, rotation {tuple(quat.shape)}")
    pg = torch.empty(B, N, 4, device=points.device, dtype=torch.float32)
    coords = base = inb = None
    if want_aux:
        coords = torch.empty(B, N, 3, device=points.device, dtype=torch.float32)
        base = torch.empty(B, N, 3, device=points.device, dtype=torch.int32)
        inb = torch.empty(B, N, device=points.device, dtype=torch.uint8)
    srt = bins = None
    if want_bins:
        srt = torch.empty(B, N, 4, device=points.device, dtype=torch.float32)
        bins = torch.empty(B, lib.b3d_pc_bin_count(V) + 1, device=points.device, dtype=torch.int32)
    check(lib.b3d_pc_project(ptr(points), ptr(quat), B, N, V, fov, cam_dist, ptr(pg), ptr(coords), ptr(base),
                             ptr(inb), ptr(srt), ptr(bins), stream_ptr(points)))
    if want_bins:
        return pg, srt, bins
    return (pg, coords, base, inb) if want_aux else pg


def splat_grid(pg, V, mode="R"):
    """Materialised, clamped occupancy grid [B,V,V,V] (trilinear_interpolation.py:62-74)."""
    pg = dev(pg, "pg")
    B, N, _ = pg.shape
    grid = torch.empty(B, V, V, V, device=pg.device, dtype=torch.float32)
    check(lib.b3d_pc_splat_grid(ptr(pg), B, N, V, mode_id(mode), ptr(grid), stream_ptr(pg)))
    return grid


class _EffectiveLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, quat, scale, taps, V, mode, fov, cam_dist):
        points = dev(points.detach(), "point_cloud")
        quat = dev(quat.detach(), "rotation")
        B, N, _ = points.shape
        sc = None
        if scale is not None:
            sc = dev(scale.detach(), "scale").reshape(-1)
            if sc.numel() != B:
                raise B3DError(f"scale must hold one value per sample, got shape {tuple(scale.shape)}")
        pg, srt, bins = project(points, quat, V, fov, cam_dist, want_bins=True)
        sil = torch.empty(B, V, V, device=points.device, dtype=torch.float32)
        h = host_floats(taps)
        check(lib.b3d_pc_silhouette_fwd_hosttaps(ptr(srt), ptr(bins), ctypes.cast(h, ctypes.c_void_p), len(taps),
                                                 ptr(sc), B, N, V, mode, ptr(sil), None, 0, stream_ptr(points)))
        ctx.save_for_backward(points, quat, pg, srt, bins, sc if sc is not None else torch.empty(0))
        ctx.cfg = (taps, V, mode, fov, cam_dist, scale.shape if scale is not None else None)
        return sil

    @staticmethod
    def backward(ctx, dsil):
        points, quat, pg, srt, bins, sc = ctx.saved_tensors
        taps, V, mode, fov, cam_dist, scale_shape = ctx.cfg
        has_scale = scale_shape is not None
        B, N, _ = points.shape
        dsil = dev(dsil, "grad_output")
        dpg = torch.empty_like(pg)
        dscale = torch.empty(B, device=points.device, dtype=torch.float32) if has_scale else None
        h = host_floats(taps)
        st = stream_ptr(points)
        check(lib.b3d_pc_silhouette_bwd_hosttaps(ptr(srt), ptr(bins), ctypes.cast(h, ctypes.c_void_p), len(taps),
                                                 ptr(sc) if has_scale else None, ptr(dsil), B, N, V, mode, ptr(dpg),
                                                 ptr(dscale), None, 0, st))
        dpoints = torch.empty_like(points)
        dquat = torch.empty_like(quat)
        check(lib.b3d_pc_project_bwd(ptr(points), ptr(quat), ptr(pg), ptr(dpg), B, N, V, fov, cam_dist, ptr(dpoints),
                                     ptr(dquat), st))
        return (dpoints, dquat, dscale.view(scale_shape) if has_scale else None, None, None, None, None, None)


def effective_loss(points, quat, scale=None, V=64, taps=None, mode="R", fov=FIELD_OF_VIEW,
                   cam_dist=CAMERA_VIEW_DISTANCE):
    """points [B,N,3] (z,y,x), quat [B,4], scale [B,1]|None -> silhouette [B,V,V] (differentiable)."""
    if taps is None:
        taps = smoothing_taps(3.0, 21, mode)
    return _EffectiveLoss.apply(points, quat, scale, list(taps), int(V), mode_id(mode), float(fov), float(cam_dist))
