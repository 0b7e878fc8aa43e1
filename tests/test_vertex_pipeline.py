"""Fused vertex pipeline (csrc/vertex_kernels.cu) against the torch composition that
tests/test_mesh_oracle.py / test_pose.py pin to the reference (MeshTemplate.get_vertex_positions, mesh_template.py:125-149;
transform_vertices, run_reconstruction.py:237-252).  fp32 tolerance 2e-6 absolute on positions of O(1), 1e-5 relative on
gradients (different summation order of the bilinear taps / atomics)."""
import os
import struct
import tempfile

import pytest
import torch

from tools.uvsphere import write_uvsphere_obj


def template(sym, device, rings=16):
    from rendering.mesh_template import MeshTemplate
    path = write_uvsphere_obj(os.path.join(tempfile.mkdtemp(), f"uvsphere_{rings}rings.obj"), rings=rings)
    return MeshTemplate(path, is_symmetric=sym, device=device)


@pytest.mark.parametrize("sym", [True, False])
def test_records_reproduce_get_vertex_positions_on_cpu(sym):
    """Host logic: the per-vertex records (taps, weights, frames, sign) restate get_vertex_positions exactly."""
    mt = template(sym, "cpu")
    rec, sgn = mt._vertex_records(32, 32)
    V = mt.topo_map.shape[0]
    raw = rec.numpy().tobytes()
    D = torch.randn(2, 3, 32, 32) * 0.05
    ref = mt.get_vertex_positions(D)
    out = torch.zeros(2, V, 3)
    flat = D.reshape(2, 3, -1)
    for v in range(V):
        f = struct.unpack_from("<4i4f9f3f", raw, v * 80)
        local = sum(f[4 + i] * flat[:, :, f[i]] for i in range(4))
        m = local @ torch.tensor(f[8:17]).view(3, 3)
        m[:, 0] *= sgn[v, 0]
        out[:, v] = torch.tensor(f[17:20]) + m
    assert float((out - ref).abs().max()) < 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("sym,rings,use_z0,channels_last", [(True, 16, False, False), (True, 31, True, True), (False, 16, True, False)])
def test_fused_pipeline_matches_torch_composition(sym, rings, use_z0, channels_last):
    from rendering.pose import transform_vertices
    dev = "cuda:0"
    mt = template(sym, dev, rings)
    g = torch.Generator().manual_seed(rings + int(sym))
    B = 5
    D0 = torch.randn(B, 3, 32, 32, generator=g) * 0.05
    s0 = 0.5 + 0.4 * torch.rand(B, 1, generator=g)
    t0 = (torch.rand(B, 3, generator=g) - 0.5) * 0.4
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1).to(dev)
    z00 = 1 + torch.exp(torch.rand(B, 1, generator=g))
    wr, wv = torch.randn(B, mt.topo_map.shape[0], 3, generator=g).to(dev), torch.randn(B, mt.topo_map.shape[0], 3, generator=g).to(dev)

    class Params:                                                 # stands in for DatasetParams (z0 lookup)
        def __init__(self, z):
            self.z = z

        def __call__(self, idx, kind):
            return self.z

    res = []
    for fused in (False, True):
        D = D0.to(dev).requires_grad_(True)
        Dm = D.contiguous(memory_format=torch.channels_last) if channels_last else D
        s, t, z0 = s0.to(dev).requires_grad_(True), t0.to(dev).requires_grad_(True), z00.to(dev).requires_grad_(True)
        if fused:
            raw, vtx = mt.vertices_and_pose(Dm, s, t, q, z0 if use_z0 else None)
        else:
            mt.disable_fused_vertices = True
            raw = mt.get_vertex_positions(Dm)
            mt.disable_fused_vertices = False
            vtx = transform_vertices(raw, s, t, q, None, Params(z0), optimize_deltas=False, optimize_z0=use_z0)
        loss = (raw * wr).sum() + (vtx * wv).sum()
        grads = torch.autograd.grad(loss, [D, s, t] + ([z0] if use_z0 else []))
        res.append([raw.detach(), vtx.detach()] + list(grads))
    for a, b, name in zip(res[0], res[1], ("raw", "vtx", "d_map", "d_scale", "d_trans", "d_z0")):
        assert a.shape == b.shape, name
        err, ref = float((a - b).abs().max()), float(a.abs().max())
        assert err <= (2e-6 * max(ref, 1.0) if name in ("raw", "vtx") else 2e-5 * ref), (name, err, ref)
    # raw vertices only (no pose)
    raw_only, none = mt.vertices_and_pose(D0.to(dev))
    assert none is None and torch.equal(raw_only, res[1][0])


@pytest.mark.gpu
@pytest.mark.parametrize("B,V,F", [(3, 50, 96), (32, 482, 960)])
def test_face_normals_kernel_matches_torch(B, V, F):
    """b3d.mesh.face_normals (csrc/loss_kernels.cu) against the torch composition of MeshTemplate.compute_normals
    (rendering/mesh_template.py:113-123: gather, cross, F.normalize): forward 1e-6, vertex gradient 1e-5 of the largest
    magnitude (fp32 atomics order)."""
    import torch.nn.functional as Fn
    from b3d.mesh import face_normals
    g = torch.Generator().manual_seed(B + F)
    verts = torch.randn(B, V, 3, generator=g).cuda().requires_grad_(True)
    faces = torch.stack([torch.randperm(V, generator=g)[:3] for _ in range(F)]).cuda()
    got = face_normals(verts, faces)
    a, b, c = verts[:, faces[:, 0]], verts[:, faces[:, 1]], verts[:, faces[:, 2]]
    ref = Fn.normalize(torch.cross(b - a, c - a, dim=2), dim=2)
    assert float((got - ref).abs().max()) <= 1e-6
    gy = torch.randn(B, F, 3, generator=g).cuda()
    ga, = torch.autograd.grad(got, verts, gy)
    gb, = torch.autograd.grad(ref, verts, gy)
    assert float((ga - gb).abs().max()) <= 1e-5 * float(gb.abs().max())
