"""the drop-in UnsupervisedLoss / SupervisedLoss against golden values produced by the reference's own
classes (tests/golden/make_golden_silhouette_losses.py; execution patch D6 only) — CPU, device-agnostic torch code."""
import os
import sys

import numpy as np
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden_silhouette_losses import inputs          # noqa: E402  (the seeded inputs only; no reference import)


def test_losses_match_reference_golden():
    from models.supervised_part import SupervisedLoss
    from models.unsupervised_part import UnsupervisedLoss
    d = np.load(os.path.join(GOLDEN, "silhouette_losses.npz"))
    projection, masks, ensemble, student = inputs()
    projection.requires_grad_(True); student.requires_grad_(True)
    loss = UnsupervisedLoss(4, 20.0)
    out = loss((projection, ensemble, student), masks, True)
    out["total_loss"].backward()
    for k in ("projection_loss", "student_loss", "total_loss"):
        assert abs(float(out[k]) - float(d[k])) <= 1e-6 * abs(float(d[k])), k
    assert np.array_equal(loss.minimum_indexes.numpy(), d["minimum_indexes"])          # argmin indices: exact
    assert np.allclose(projection.grad.numpy(), d["d_projection"], atol=1e-7)
    assert np.allclose(student.grad.numpy(), d["d_student"], atol=1e-6)
    ev = loss((projection.detach()[:5],), masks, False)["projection_loss"]
    assert abs(float(ev) - float(d["eval_projection_loss"])) <= 1e-6 * float(d["eval_projection_loss"])
    sup = SupervisedLoss()(projection.detach()[:5], masks)["full_loss"]
    assert abs(float(sup) - float(d["supervised_full_loss"])) <= 1e-6 * float(d["supervised_full_loss"])


def test_repeat_tensor_for_each_element_in_batch():
    from utils.batch_repetition import repeat_tensor_for_each_element_in_batch
    t = torch.arange(6.).reshape(2, 3)
    r = repeat_tensor_for_each_element_in_batch(t, 3)
    assert r.shape == (6, 3) and torch.equal(r[:3], t[:1].expand(3, 3)) and torch.equal(r[3:], t[1:].expand(3, 3))


def test_oracle_agrees_with_the_drop_in():
    """oracle/pointcloud.py's restatement (used by the effective-loss parity tests) and the drop-in give the same number."""
    sys.path.insert(0, os.path.dirname(GOLDEN.rstrip("/")).rsplit("/tests", 1)[0])
    from oracle import pointcloud as O
    from models.unsupervised_part import UnsupervisedLoss, half_resolution_masks
    projection, masks, ensemble, student = inputs(seed=2)
    out = UnsupervisedLoss(4, 20.0)((projection, ensemble, student), masks, True)
    ref = O.candidate_min_loss(projection, O.downsample_mask_half(masks), 4)
    ref_loss = ref[0] if isinstance(ref, tuple) else ref
    assert abs(float(out["projection_loss"]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert torch.equal(half_resolution_masks(masks), O.downsample_mask_half(masks))


import pytest  # noqa: E402


@pytest.mark.gpu
def test_losses_match_reference_golden_on_cuda():
    """The same golden (reference classes, CPU) with every tensor on the GPU: the drop-ins are device agnostic; fp32 reductions
    run in another order there (1e-5 relative on the losses, 1e-6 absolute on the gradients), the argmin indices stay exact."""
    from models.supervised_part import SupervisedLoss
    from models.unsupervised_part import UnsupervisedLoss
    d = np.load(os.path.join(GOLDEN, "silhouette_losses.npz"))
    projection, masks, ensemble, student = [t.cuda() for t in inputs()]
    projection.requires_grad_(True); student.requires_grad_(True)
    loss = UnsupervisedLoss(4, 20.0)
    out = loss((projection, ensemble, student), masks, True)
    out["total_loss"].backward()
    for k in ("projection_loss", "student_loss", "total_loss"):
        assert abs(float(out[k]) - float(d[k])) <= 1e-5 * abs(float(d[k])), k
    assert np.array_equal(loss.minimum_indexes.cpu().numpy(), d["minimum_indexes"])
    assert np.allclose(projection.grad.cpu().numpy(), d["d_projection"], atol=1e-6)
    assert np.allclose(student.grad.cpu().numpy(), d["d_student"], atol=1e-5)
    sup = SupervisedLoss()(projection.detach()[:5], masks)["full_loss"]
    assert abs(float(sup) - float(d["supervised_full_loss"])) <= 1e-5 * float(d["supervised_full_loss"])


@pytest.mark.gpu
def test_candidate_loss_on_kernel_silhouettes_matches_the_oracle_pipeline():
    """Rows a1-a6 end to end on the GPU: K pose candidates per sample -> EffectiveLossFunction (libb3d point-cloud kernels)
    -> UnsupervisedLoss (min over candidates + student pose term), against the oracle's silhouettes + candidate_min_loss on the
    CPU: chosen candidate per sample exact, losses to 1e-4 relative (fp32 silhouettes of two implementations)."""
    sys.path.insert(0, os.path.dirname(GOLDEN.rstrip("/")).rsplit("/tests", 1)[0])
    from oracle import pointcloud as O
    from models.unsupervised_part import UnsupervisedLoss, half_resolution_masks
    from utils.batch_repetition import repeat_tensor_for_each_element_in_batch
    from utils.effective_loss_function import EffectiveLossFunction
    B, K, N, V = 3, 4, 600, 32
    g = torch.Generator().manual_seed(11)
    pts = (torch.rand(B, N, 3, generator=g) * 2 - 1) * 0.4
    quats = torch.nn.functional.normalize(torch.randn(B * K, 4, generator=g), dim=-1)
    student = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    masks = (torch.rand(B, 2 * V, 2 * V, generator=g) > 0.6).float()
    pk = repeat_tensor_for_each_element_in_batch(pts, K)                    # every candidate sees its sample's cloud
    elf = EffectiveLossFunction(voxel_size=V, kernel_size=21, smooth_sigma=3.0).cuda()
    p = pk.cuda().requires_grad_(True)
    sil = elf(p, quats.cuda(), None)
    crit = UnsupervisedLoss(K, 20.0)
    out = crit((sil, quats.cuda(), student.cuda().requires_grad_(True)), masks.cuda(), True)
    out["total_loss"].backward()
    ref_sil = O.effective_loss_forward(pk, quats, None, V=V, kernel_size=21, sigma=3.0, mode="R")
    ref_loss, ref_idx = O.candidate_min_loss(ref_sil, O.downsample_mask_half(masks), K)
    assert torch.equal(crit.minimum_indexes.cpu(), ref_idx)
    assert abs(float(out["projection_loss"]) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    assert torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0
    # only the chosen candidate of every sample receives a gradient (the min is a selection)
    per_cand = p.grad.abs().sum(dim=(1, 2)).view(B, K).cpu()
    chosen = torch.zeros(B, K, dtype=torch.bool)
    chosen[torch.arange(B), ref_idx] = True
    assert bool((per_cand[~chosen] == 0).all()) and bool((per_cand[chosen] > 0).all())
    assert torch.allclose(half_resolution_masks(masks.cuda()).cpu(), O.downsample_mask_half(masks), atol=1e-6)
