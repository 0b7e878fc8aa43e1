"""Fused vertex pipeline over libb3d (csrc/vertex_kernels.cu): displacement map -> template vertices -> camera space.

Reference chain: MeshTemplate.get_vertex_positions (/root/reference/code/rendering/mesh_template.py:125-149) and
transform_vertices (run_reconstruction.py:237-252) with qrot (rendering/utils.py:36-46)."""
import struct

import torch

from . import B3DError, check, dev, lib, ptr, stream_ptr


def pack_records(sites, padded_w, h, w, wrap, frames, v0, sign):
    """Per-vertex records for the kernels.  sites [V,2] grid_sample coordinates into the PADDED map of width padded_w
    (wrap(xp) -> column of the unpadded map), frames [V,3,3], v0 [V,3], sign [V] (x factor)."""
    sites = sites.detach().cpu().float()
    ix = ((sites[:, 0] + 1) / 2) * (padded_w - 1)            # F.grid_sample, align_corners=True (fp32 like ATen)
    iy = ((sites[:, 1] + 1) / 2) * (h - 1)
    x0, y0 = ix.floor(), iy.floor()
    x1, y1 = x0 + 1, y0 + 1
    taps = [(x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)), (x0, y1, (x1 - ix) * (iy - y0)),
            (x1, y1, (ix - x0) * (iy - y0))]
    rec = bytearray()
    fr, vv = frames.detach().cpu().float(), v0.detach().cpu().float()
    for v in range(sites.shape[0]):
        ti, tw = [], []
        for xs, ys, ws in taps:
            x, y, wt = int(xs[v]), int(ys[v]), float(ws[v])
            if 0 <= x < padded_w and 0 <= y < h:
                ti.append(y * w + wrap(x))
                tw.append(wt)
            else:
                ti.append(0)
                tw.append(0.0)
        rec += struct.pack("<4i4f9f3f", *ti, *tw, *fr[v].reshape(-1).tolist(), *vv[v].tolist())
    sg = torch.zeros(sites.shape[0], 4)
    sg[:, 0] = sign
    return torch.frombuffer(rec, dtype=torch.uint8).clone(), sg


class _VertexPipeline(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dmap, scale, trans, rot, z0, rec, sgn, V):
        d = dev(dmap.detach(), "displacement_map") if dmap.is_contiguous() else dmap.detach()
        if not d.is_cuda or d.dtype != torch.float32:
            raise B3DError("vertex pipeline runs on CUDA fp32 tensors only (no CPU fallback)")
        B, C, h, w = d.shape
        if C != 3:
            raise B3DError("vertex pipeline: displacement map must have 3 channels")
        sn, sc, sy, sx = d.stride()
        raw = torch.empty(B, V, 3, device=d.device, dtype=torch.float32)
        pose = scale is not None
        vtx = torch.empty_like(raw) if pose else None
        s = dev(scale.detach().reshape(B), "scale") if pose else None
        t = dev(trans.detach().reshape(B, 3), "translation") if pose else None
        q = dev(rot.detach().reshape(B, 4), "rotation") if pose else None
        z = dev(z0.detach().reshape(B), "z0") if z0 is not None else None
        check(lib.b3d_vertex_pipeline_fwd(ptr(d), sn, sc, sy, sx, h, w, ptr(rec), ptr(sgn), B, V, ptr(s), ptr(t), ptr(q), ptr(z),
                                          ptr(raw), ptr(vtx), stream_ptr(d)))
        ctx.save_for_backward(raw, s, t, q, z, rec, sgn)
        ctx.cfg = (B, V, h, w, pose, scale.shape if pose else None, trans.shape if pose else None, z0.shape if z0 is not None else None)
        if pose:
            return raw, vtx
        ctx.mark_non_differentiable()
        return raw, raw.new_empty(0)

    @staticmethod
    def backward(ctx, g_raw, g_vtx):
        raw, s, t, q, z, rec, sgn = ctx.saved_tensors
        B, V, h, w, pose, sshape, tshape, zshape = ctx.cfg
        dd = torch.zeros(B, 3, h, w, device=raw.device) if ctx.needs_input_grad[0] else None
        need_s = pose and ctx.needs_input_grad[1]
        need_t = pose and ctx.needs_input_grad[2]
        need_z = z is not None and ctx.needs_input_grad[4]
        if pose and ctx.needs_input_grad[3]:
            raise B3DError("vertex pipeline: no gradient w.r.t. the rotation (the reference's poses are data)")
        ds = torch.zeros(B, device=raw.device) if need_s else None
        dt = torch.zeros(B, 3, device=raw.device) if need_t else None
        dz = torch.zeros(B, device=raw.device) if need_z else None
        gr = dev(g_raw, "grad_raw") if g_raw is not None else None
        gv = dev(g_vtx, "grad_vtx") if (g_vtx is not None and pose) else None
        if gr is None and gv is None:
            return (None,) * 8
        check(lib.b3d_vertex_pipeline_bwd(ptr(gr), ptr(gv), ptr(raw), ptr(rec), ptr(sgn), B, V, h, w, ptr(s), ptr(t), ptr(q), ptr(z),
                                          ptr(dd), ptr(ds), ptr(dt), ptr(dz), stream_ptr(raw)))
        return (dd, ds.view(sshape) if need_s else None, dt.view(tshape) if need_t else None, None,
                dz.view(zshape) if need_z else None, None, None, None)


def vertex_pipeline(dmap, rec, sgn, V, scale=None, trans=None, rot=None, z0=None):
    """-> (raw [B,V,3], vtx [B,V,3] or None)."""
    raw, vtx = _VertexPipeline.apply(dmap, scale, trans, rot, z0, rec, sgn, V)
    return raw, (vtx if scale is not None else None)
