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
