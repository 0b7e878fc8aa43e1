"""Dense-grid kernels (mode P, VoxelsSmooth / TrilinearInterpolation / CameraUtilities / termination_probs drop-ins)
against oracle/pointcloud.py.  Mode P has no executable reference (parity unpinned); mode R dense must agree with
the fused mode-R kernel and with the reference-pinned oracle."""
import pytest
import torch

from oracle import pointcloud as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cloud(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    p = (torch.rand(B, N, 3, generator=g) * 2 - 1) * 0.45
    return p, torch.randn(B, 4, generator=g), 0.5 + 0.5 * torch.rand(B, 1, generator=g), g


@pytest.mark.parametrize("B,N,V,with_scale", [(2, 700, 32, True), (1, 3000, 64, False), (2, 400, 40, True)])
def test_mode_p_matches_oracle(B, N, V, with_scale):
    from utils.effective_loss_function import EffectiveLossFunction
    p, q, s, g = cloud(B, N, V + N)
    s = s if with_scale else None
    wts = torch.rand(B, V, V, generator=g)
    outs = {}
    for dt in (torch.float32, torch.float64):
        po, qo = p.to(dt).requires_grad_(True), q.to(dt).requires_grad_(True)
        so = s.to(dt).requires_grad_(True) if s is not None else None
        sil = O.effective_loss_forward(po, qo, so, V=V, kernel_size=21, sigma=2.0, mode="P")
        grads = torch.autograd.grad((sil * wts.to(dt)).sum(), [po, qo] + ([so] if s is not None else []))
        outs[dt] = [sil.detach()] + list(grads)
    m = EffectiveLossFunction(voxel_size=V, smooth_sigma=2.0, semantics="P").to(DEV)
    pc, qc = p.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    sc = s.to(DEV).requires_grad_(True) if s is not None else None
    sil = m(pc, qc, sc)
    grads = torch.autograd.grad((sil * wts.to(DEV)).sum(), [pc, qc] + ([sc] if s is not None else []))
    for got, o32, o64, name in zip([sil] + list(grads), outs[torch.float32], outs[torch.float64], ["sil", "dp", "dq", "ds"]):
        gap = float((o32.double() - o64).abs().max())
        tol = 4 * gap + 2e-5 * max(1.0, float(o64.abs().max()))
        err = float((got.detach().cpu().double() - o64).abs().max())
        assert err <= tol, (name, err, tol)


def test_dense_mode_r_agrees_with_fused_kernel():
    from b3d.pointcloud import effective_loss, effective_loss_dense, smoothing_taps
    p, q, s, g = cloud(2, 2000, 5)
    p, q, s = p.to(DEV), q.to(DEV), s.to(DEV)
    taps = smoothing_taps(3.0, 21, "R")
    a = effective_loss(p, q, s, V=64, taps=taps, mode="R")
    b = effective_loss_dense(p, q, s, V=64, taps=taps, mode="R")
    d = (a - b).abs()
    assert float(d.mean()) < 1e-3 and float((d > 2e-2).float().mean()) < 1e-2      # mode R is ill-conditioned (D5)


def test_standalone_dropins():
    from camera.coordinate_system_transformation import CameraUtilities
    from quaternions.points_quaternions import PointsQuaternionsRotator
    from utils.effective_loss_function import EffectiveLossFunction
    from utils.smooth_voxels import VoxelsSmooth
    from utils.trilinear_interpolation import TrilinearInterpolation
    p, q, s, g = cloud(2, 500, 9)
    # CameraUtilities == oracle.project, with gradients
    po, qo = p.clone().requires_grad_(True), q.clone().requires_grad_(True)
    co = O.project(po, qo)
    w = torch.rand(co.shape, generator=g)
    go = torch.autograd.grad((co * w).sum(), [po, qo])
    pc, qc = p.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    cc = CameraUtilities().transformation_3d_coord_to_camera_coord(pc, qc, 1.875, 2.0)
    gc = torch.autograd.grad((cc * w.to(DEV)).sum(), [pc, qc])
    assert torch.equal(cc.detach().cpu(), co.detach())
    assert torch.allclose(gc[0].cpu(), go[0], atol=1e-5) and torch.allclose(gc[1].cpu(), go[1], atol=1e-4)
    assert torch.allclose(PointsQuaternionsRotator.rotate_points(p, q, False), O.rotate_points(p, q), atol=1e-6)
    # TrilinearInterpolation == oracle.splat (mode R), VoxelsSmooth == oracle.smooth, termination_probs
    V = 32
    occ_o, _, _ = O.splat(co.detach(), V, "R")
    occ = TrilinearInterpolation(size=V).trilinear_interpolation(cc.detach())
    assert float((occ.cpu() - occ_o).abs().max()) < 2e-3
    vs = VoxelsSmooth()
    ker = vs.separate_kernels(3.0, 21)
    sm = vs.smooth(occ, ker, s.to(DEV))
    sm_o = O.smooth(occ.cpu(), O.kernel_1d(3.0, 21, "R"), s, "R")
    assert float((sm.cpu() - sm_o).abs().max()) < 1e-5
    probs = EffectiveLossFunction(voxel_size=V).to(DEV).termination_probs(sm)
    assert float((probs.cpu() - O.termination_probs(sm.cpu(), "R")).abs().max()) < 1e-5
    sm_p = VoxelsSmooth("P").smooth(occ, VoxelsSmooth("P").separate_kernels(1.5, 21), None)
    assert float((sm_p.cpu() - O.smooth(occ.cpu(), O.kernel_1d(1.5, 21, "P"), None, "P")).abs().max()) < 1e-5
