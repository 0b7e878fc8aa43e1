"""Parity of the CUDA point-cloud path (through the drop-in module and the C ABI) with the oracle and
with the golden vectors produced by the reference's own classes.

Tolerances.  Index / visibility buffers: bit exact.  Floating point: mode R is ill-conditioned (its
corner weights are O(100)), so fp32 results of *any* evaluation order differ from
the exact value by what the fp32-vs-fp64 oracle gap shows; the CUDA result must stay within
TOL_K x that gap (+1e-5 absolute) of the fp64 oracle, i.e. be as good as the reference's own fp32.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import pointcloud as O

pytestmark = pytest.mark.gpu
TOL_K = 4.0


def _modules():
    import b3d
    from b3d import pointcloud as pc
    from utils.effective_loss_function import EffectiveLossFunction
    return b3d, pc, EffectiveLossFunction


def oracle_both(points, q, scale, wts, V, ksize, sigma, mode="R"):
    out = {}
    for dt in (torch.float32, torch.float64):
        p = points.detach().cpu().to(dt).requires_grad_(True)
        r = q.detach().cpu().to(dt).requires_grad_(True)
        s = scale.detach().cpu().to(dt).requires_grad_(True) if scale is not None else None
        sil = O.effective_loss_forward(p, r, s, V=V, kernel_size=ksize, sigma=sigma, mode=mode)
        loss = (sil * wts.cpu().to(dt)).sum()
        gs = torch.autograd.grad(loss, [p, r] + ([s] if s is not None else []))
        out[dt] = [sil.detach()] + [g for g in gs]
    return out


def assert_close_to_oracle(cuda_vals, orc, names):
    for v, o32, o64, name in zip(cuda_vals, orc[torch.float32], orc[torch.float64], names):
        gap = float((o32.double() - o64).abs().max())
        scale = max(1.0, float(o64.abs().max()))
        tol = TOL_K * gap + 1e-5 * scale
        err = float((v.detach().cpu().double() - o64).abs().max())
        assert err <= tol, f"{name}: |cuda - oracle64| = {err:.3e} > {tol:.3e} (fp32 oracle gap {gap:.3e})"


@pytest.mark.parametrize("name", ["v32_scale", "v32_noscale", "v64_cfg1", "v32_oob"])
def test_golden_vectors_from_reference(name):
    b3d, pc, ELF = _modules()
    d = np.load(os.path.join(GOLDEN, f"pointcloud_{name}.npz"))
    V = int(d["V"])
    dev = torch.device("cuda:0")
    p = torch.tensor(d["points"], device=dev, requires_grad=True)
    q = torch.tensor(d["q"], device=dev, requires_grad=True)
    s = torch.tensor(d["scale"], device=dev, requires_grad=True) if "scale" in d else None
    wts = torch.tensor(d["wts"], device=dev)

    # index / visibility buffers of the reference: bit exact
    pg, coords, base, inb = pc.project(p.detach(), q.detach(), V, want_aux=True)
    assert np.array_equal(inb.cpu().numpy().astype(bool), d["inb"])
    assert np.array_equal(base.cpu().numpy()[d["inb"]], d["base"][d["inb"]].astype(np.int32))
    np.testing.assert_array_equal(coords.cpu().numpy(), d["coords"])

    m = ELF(voxel_size=V, kernel_size=int(d["ksize"]), smooth_sigma=float(d["sigma"])).to(dev)
    sil = m(p, q, s)
    loss = (sil * wts).sum()
    grads = torch.autograd.grad(loss, [p, q] + ([s] if s is not None else []))
    orc = oracle_both(p, q, s, wts, V, int(d["ksize"]), float(d["sigma"]))
    # the golden values ARE the fp32 oracle column (pinned in test_pointcloud_oracle.py)
    np.testing.assert_allclose(orc[torch.float32][0].numpy(), d["sil"], atol=2e-6)
    assert_close_to_oracle([sil] + list(grads), orc, ["sil", "d_points", "d_q", "d_scale"])

    occ = pc.splat_grid(pg, V, "R")
    np.testing.assert_allclose(occ.cpu().numpy()[:, ::7, ::5, ::3], d["occ_probe"], atol=2e-4)


@pytest.mark.parametrize("B,N,V,ksize,sigma,with_scale", [
    (4, 1024, 64, 21, 3.0, True),      # BASELINE config 1
    (1, 8000, 64, 21, 2.0, True),      # 8000-point cloud, reference default V
    (1, 4000, 128, 21, 3.0, False),    # V = 128 (256^2 images / 2)
    (3, 333, 40, 21, 1.2, True),       # ragged: V not a multiple of the patch, N not of the block
    (2, 500, 32, 9, 1.0, True),        # generic tap count
])
def test_random_clouds_against_oracle(B, N, V, ksize, sigma, with_scale):
    b3d, pc, ELF = _modules()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + N + V)
    pts = (torch.rand(B, N, 3, generator=g) * 2 - 1) * 0.45
    if N >= 4000:   # half of the points on a noisy sphere shell: realistic collision pattern
        sph = torch.nn.functional.normalize(torch.randn(B, N // 2, 3, generator=g), dim=-1)
        pts[:, : N // 2] = sph * (0.33 + 0.01 * torch.randn(B, N // 2, 1, generator=g))
    p = pts.to(dev).requires_grad_(True)
    q = torch.randn(B, 4, generator=g).to(dev).requires_grad_(True)
    s = (0.5 + 0.5 * torch.rand(B, 1, generator=g)).to(dev).requires_grad_(True) if with_scale else None
    wts = torch.rand(B, V, V, generator=g).to(dev)
    m = ELF(voxel_size=V, kernel_size=ksize, smooth_sigma=sigma).to(dev)
    sil = m(p, q, s)
    grads = torch.autograd.grad((sil * wts).sum(), [p, q] + ([s] if s is not None else []))
    orc = oracle_both(p, q, s, wts, V, ksize, sigma)
    assert_close_to_oracle([sil] + list(grads), orc, ["sil", "d_points", "d_q", "d_scale"])
    # index buffer against the oracle, bit exact
    pg, coords, base, inb = pc.project(p.detach(), q.detach(), V, want_aux=True)
    c = O.project(p.detach().cpu(), q.detach().cpu())
    assert np.array_equal(inb.cpu().numpy().astype(bool), O.inbounds(c).numpy())
    ob = ((V - 1) * (c + 0.5)).floor().to(torch.int32)
    assert torch.equal(base.cpu(), ob)


def test_edge_cases():
    b3d, pc, ELF = _modules()
    dev = torch.device("cuda:0")
    V = 32
    m = ELF(voxel_size=V).to(dev)
    q = torch.tensor([[1.0, 0, 0, 0], [0.3, -0.2, 0.9, 0.1]], device=dev)
    # empty cloud: every cell has occupancy 0 -> clamp eps -> analytic value
    empty = m(torch.zeros(2, 0, 3, device=dev), q, torch.ones(2, 1, device=dev))
    exp_empty = O.effective_loss_forward(torch.zeros(2, 0, 3), q.cpu(), torch.ones(2, 1), V=V)
    assert torch.allclose(empty.cpu(), exp_empty, atol=1e-7)
    # all points outside the frustum give the same image as the empty cloud
    far = m(torch.full((2, 64, 3), 3.0, device=dev), q, None)
    assert torch.allclose(far, empty, atol=1e-7)
    # batch of zero samples
    assert m(torch.zeros(0, 10, 3, device=dev), torch.zeros(0, 4, device=dev)).shape == (0, V, V)
    # NaN taps of the reference for small sigma propagate (D4), mode R
    m2 = ELF(voxel_size=V, smooth_sigma=0.5).to(dev)
    out = m2(torch.rand(1, 50, 3, device=dev) * 0.5 - 0.25, q[:1], None)
    assert torch.isnan(out).all()
    # the sigma schedule re-assigns the buffer (training_test_shape_net.py:29)
    m.sigma = torch.empty_like(m.sigma).fill_(2.0)
    a = m(torch.rand(1, 50, 3, device=dev) * 0.5 - 0.25, q[:1], None)
    assert torch.isfinite(a).all()


def test_full_size_properties():
    """BASELINE config 2 size (B=16, N=8000, V=128): properties that need no oracle run."""
    b3d, pc, ELF = _modules()
    dev = torch.device("cuda:0")
    B, N, V = 16, 8000, 128
    g = torch.Generator().manual_seed(7)
    p = ((torch.rand(B, N, 3, generator=g) * 2 - 1) * 0.45).to(dev).requires_grad_(True)
    q = torch.randn(B, 4, generator=g).to(dev).requires_grad_(True)
    s = (0.5 + 0.5 * torch.rand(B, 1, generator=g)).to(dev).requires_grad_(True)
    m = ELF(voxel_size=V).to(dev)
    sil = m(p, q, s)
    assert sil.shape == (B, V, V) and torch.isfinite(sil).all()
    assert float(sil.min()) >= 0.0 and float(sil.max()) <= 1.0 + 1e-4
    # Mode R multiplies three corner weights of magnitude ~2g, so a 1-ulp change of a
    # coordinate moves single cells by O(0.1): invariances hold in the mean, not per pixel.
    def close(a, b):
        d = (a - b).abs()
        return float(d.mean()) < 1e-3 and float((d > 2e-2).float().mean()) < 1e-2
    # permutation of the points changes only the fp summation order
    perm = torch.randperm(N, generator=g).to(dev)
    assert close(sil, m(p[:, perm], q, s))
    # quaternion scale invariance: q is normalised inside (points_quaternions.py:53)
    assert close(sil, m(p, q * 3.0, s))
    # sample independence: sample 3 alone gives the same image
    assert close(sil[3:4], m(p[3:4], q[3:4], s[3:4]))
    gp, gq, gs = torch.autograd.grad(sil.square().sum(), [p, q, s])
    assert torch.isfinite(gp).all() and torch.isfinite(gq).all() and torch.isfinite(gs).all()
    # d/dq is orthogonal to q (normalisation removes the radial component)
    rel = (gq * q).sum(-1).abs() / (gq.norm(dim=-1) * q.norm(dim=-1) + 1e-12)
    assert float(rel.max()) < 1e-2
