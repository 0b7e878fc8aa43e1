"""oracle/mesh.py: pinned pieces against golden vectors from the reference's Python; the DIB-R
restatement (parity unpinned, kaolin absent) for self-consistency.  CPU only."""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mesh as M


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "mesh_reference_pieces.npz"))


@pytest.fixture(scope="module")
def sphere16():
    tmp = tempfile.mkdtemp()
    path = M.write_uvsphere_obj(os.path.join(tmp, "uvsphere_16rings.obj"), rings=16)
    return M.TemplateData(M.load_obj(path), path)


def test_qrot_circpad_fragmentshader(gold):
    t = lambda k: torch.tensor(gold[k])
    np.testing.assert_allclose(M.qrot(t("qrot_q"), t("qrot_v")).numpy(), gold["qrot_out"], atol=1e-6)
    assert np.array_equal(M.circpad(t("tex"), 2).numpy(), gold["circpad2"])
    np.testing.assert_allclose(M.fragmentshader(t("fs_uv"), t("tex"), t("fs_mask")).numpy(), gold["fs_out"], atol=1e-6)
    np.testing.assert_allclose(M.fragmentshader(t("fs_uv"), t("tex"), t("fs_mask"), t("fs_bg")).numpy(),
                               gold["fs_out_bg"], atol=1e-6)


def test_face_adjacency_and_flat_loss(gold, sphere16):
    assert np.array_equal(sphere16.ff.numpy(), gold["ff16"].astype(np.int64))
    loss = M.loss_flat(sphere16.ff, 960, torch.tensor(gold["flat_norms"]))
    np.testing.assert_allclose(loss.numpy(), gold["flat_loss"], rtol=1e-6)


def test_template_structure(sphere16):
    # SURVEY §8c (2) / App. E facts of the shipped 16-ring template hold for the procedural one
    T = sphere16
    assert T.vertices.shape == (482, 3) and T.faces.shape == (960, 3) and T.uvs.shape == (559, 2)
    assert len(T.neg_indices) == 225 and len(T.pos_indices) == 225 and len(T.zero_indices) == 32
    assert T.ff.shape == (960, 3) and int((T.ff < 0).sum()) == 0
    # zero displacement leaves the template where it is; symmetric displacement keeps x-symmetry
    v0 = M.get_vertex_positions(T, torch.zeros(1, 3, 32, 32))
    assert torch.allclose(v0[0], T.vertices, atol=1e-7)
    v = M.get_vertex_positions(T, torch.randn(2, 3, 32, 32) * 0.05)
    assert torch.allclose(v[:, T.neg_indices] * torch.tensor([-1.0, 1, 1]), v[:, T.pos_indices], atol=1e-6)
    assert float(v[:, T.zero_indices, 0].abs().max()) == 0.0


def _scene(T, B=2, dtype=torch.float32, seed=0):
    g = torch.Generator().manual_seed(seed)
    vtx = M.get_vertex_positions(T, (torch.randn(B, 3, 32, 32, generator=g) * 0.05))
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    s = 0.5 + 0.3 * torch.rand(B, 1, generator=g)
    t = (torch.rand(B, 3, generator=g) - 0.5) * 0.3
    vtx = M.transform_vertices(vtx, s, t, q).to(dtype)
    tex = (torch.rand(B, 3, 16, 16, generator=g) * 2 - 1).to(dtype)
    return vtx, tex


def test_rasterizer_self_consistency(sphere16):
    T = sphere16
    vtx, tex = _scene(T)
    H = W = 48
    p3d, p2d, normal = M.ortho_projection(vtx, T.faces)
    uvs, texp = M.adjust_uv_and_texture(T, tex)
    c = [uvs[:, T.face_textures[:, i], :] for i in range(3)]
    one = torch.ones_like(c[0][:, :, :1])
    uv9 = torch.cat((c[0], one, c[1], one, c[2], one), dim=2)
    imfeat, improb, imidx, imwei = M.rasterize(p3d, p2d, normal[:, :, 2:3], uv9, H, W)
    cov = imidx > 0
    assert 0.05 < float(cov.float().mean()) < 0.9
    # barycentrics: non-negative, sum to one on covered pixels; hard mask channel == coverage
    assert float(imwei[cov].min()) >= 0
    assert torch.allclose(imwei[cov].sum(-1), torch.ones(int(cov.sum())), atol=1e-5)
    assert torch.equal(imfeat[..., 2] > 0.5, cov)
    # the winning face is front facing, contains the pixel centre, and is the nearest such face
    x0, y0 = M.pixel_centres(H, W, torch.float32, "cpu")
    b, y, x = [t[:50] for t in torch.nonzero(cov, as_tuple=True)]
    for bi, yi, xi in zip(b.tolist(), y.tolist(), x.tolist()):
        f = int(imidx[bi, yi, xi]) - 1
        assert float(normal[bi, f, 2]) >= 0
        px, py = float(x0[0, 0, xi]), float(y0[0, yi, 0])
        best, bestz = -1, M.DEPTH_INIT
        for g in range(T.faces.shape[0]):
            if float(normal[bi, g, 2]) < 0:
                continue
            P = (M.MULTIPLIER * p2d[bi, g]).tolist()
            m, p, n, q = P[2] - P[0], P[3] - P[1], P[4] - P[0], P[5] - P[1]
            s, t = px - P[0], py - P[1]
            k3 = m * q - n * p
            if k3 == 0:
                continue
            w1, w2 = (s * q - n * t) / k3, (m * t - s * p) / k3
            w0 = 1 - w1 - w2
            if min(w0, w1, w2) < -1e-6:
                continue
            z = w0 * float(p3d[bi, g, 2]) + w1 * float(p3d[bi, g, 5]) + w2 * float(p3d[bi, g, 8])
            if z > bestz + 1e-6:
                best, bestz = g, z
        assert best == f or abs(bestz - (imwei[bi, yi, xi] * p3d[bi, f, 2::3]).sum()) < 1e-4
    # soft silhouette: 1 inside, in [0,1) outside, decaying away from the object
    assert torch.all(improb[cov] == 1)
    out = improb[..., 0][~cov]
    assert float(out.min()) >= 0 and float(out.max()) <= 1 and float(out.mean()) < 0.5
    assert float(improb[:, 0, 0].max()) < 1e-3          # image corner is far from the object


def test_rasterizer_gradients_match_finite_differences(sphere16):
    T = sphere16
    vtx, tex = _scene(T, B=1, dtype=torch.float64, seed=3)
    H = W = 24
    uvs, texp = M.adjust_uv_and_texture(T, tex)
    g = torch.Generator().manual_seed(5)
    wi = torch.rand(1, H, W, 3, generator=g, dtype=torch.float64)
    wa = torch.rand(1, H, W, 1, generator=g, dtype=torch.float64)

    def loss_fn(v, tx):
        img, alpha, _, _ = M.render(v, T.faces, uvs.double(), M.circpad(tx, 1), T.face_textures, H, W)
        return (img * wi).sum() + (alpha * wa).sum()

    v = vtx.clone().requires_grad_(True)
    tx = tex.clone().requires_grad_(True)
    loss = loss_fn(v, tx)
    gv, gt = torch.autograd.grad(loss, [v, tx])
    assert float(gv[..., 2].abs().max()) == 0           # no gradient to depth (kaolin semantics)
    idx = torch.nonzero(gv[0, :, 0].abs() > 1e-6)[:6, 0].tolist()
    eps = 1e-6
    for i in idx:
        for c in range(2):
            vp = vtx.clone(); vp[0, i, c] += eps
            vm = vtx.clone(); vm[0, i, c] -= eps
            fd = (loss_fn(vp, tex) - loss_fn(vm, tex)) / (2 * eps)
            assert abs(float(fd) - float(gv[0, i, c])) < 1e-3 * max(1.0, abs(float(fd))), (i, c, float(fd), float(gv[0, i, c]))
    ti = torch.nonzero(gt.abs() > 1e-6)[:4].tolist()
    for (b, ch, yy, xx) in ti:
        tp = tex.clone(); tp[b, ch, yy, xx] += eps
        tm = tex.clone(); tm[b, ch, yy, xx] -= eps
        fd = (loss_fn(vtx, tp) - loss_fn(vtx, tm)) / (2 * eps)
        assert abs(float(fd) - float(gt[b, ch, yy, xx])) < 1e-5 * max(1.0, abs(float(fd)))


def test_transform_vertices_and_iou():
    g = torch.Generator().manual_seed(9)
    v = torch.randn(2, 5, 3, generator=g)
    q = torch.nn.functional.normalize(torch.randn(2, 4, generator=g), dim=-1)
    s, t = torch.rand(2, 1, generator=g) + 0.5, torch.randn(2, 3, generator=g) * 0.1
    out = M.transform_vertices(v, s, t, q)
    ref = (M.qrot(q, s.unsqueeze(-1) * v) + t.unsqueeze(1)) * torch.tensor([1.0, -1, -1])
    assert torch.allclose(out, ref)
    z0 = torch.tensor([3.7, 2.5])
    o2 = M.transform_vertices(v, s, t, q, z0=z0)
    fac = (z0.view(2, 1) + ref[..., 2] / 2) / (z0.view(2, 1) - ref[..., 2] / 2)
    assert torch.allclose(o2[..., 0], ref[..., 0] * fac) and torch.allclose(o2[..., 2], ref[..., 2])
    a = torch.zeros(1, 4, 4); a[0, :2] = 1
    b = torch.zeros(1, 4, 4); b[0, 1:3] = 1
    assert abs(float(M.mean_iou(a, b)) - 4 / 12) < 1e-6
