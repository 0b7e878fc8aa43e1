"""SURVEY §8 rows a11 / a12 helpers: rendering.pose.{transform_vertices, mean_iou} against golden values obtained by running
the reference's own script-level functions (tests/golden/make_golden_pose.py), and against the oracle's restatement."""
import os
import sys

import numpy as np
import torch

from conftest import GOLDEN, ROOT

sys.path.insert(0, GOLDEN)
from make_golden_pose import inputs, load_params          # noqa: E402  (seeded inputs; no reference import at module level)


def test_pose_helpers_match_reference_golden():
    from models import reconstruction
    from rendering.pose import mean_iou, transform_vertices
    d = np.load(os.path.join(GOLDEN, "pose_reference.npz"))
    vtx, scale, trans, rot, idx, state, alpha_p, alpha_r = inputs()
    for tag, deltas, z0 in (("plain", False, False), ("deltas", True, False), ("full", True, True)):
        dp = load_params(reconstruction, state, deltas, z0)
        out = transform_vertices(vtx, scale, trans, rot, idx, dataset_params=dp, optimize_deltas=deltas, optimize_z0=z0)
        assert np.allclose(out.detach().numpy(), d["vtx_" + tag], atol=1e-6), tag
    assert abs(float(mean_iou(alpha_p, alpha_r)) - float(d["iou"])) < 1e-7


def test_pose_helpers_match_oracle():
    sys.path.insert(0, ROOT)
    from oracle import mesh as M
    from rendering.pose import mean_iou, transform_vertices
    vtx, scale, trans, rot, _, _, alpha_p, alpha_r = inputs(seed=8)
    assert torch.allclose(transform_vertices(vtx, scale, trans, rot), M.transform_vertices(vtx, scale, trans, rot), atol=1e-7)
    assert float(mean_iou(alpha_p, alpha_r)) == float(M.mean_iou(alpha_p, alpha_r))
