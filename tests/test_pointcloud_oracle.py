"""oracle/pointcloud.py (mode R) against golden vectors produced by the reference's own classes
(tests/golden/make_golden_pointcloud.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import pointcloud as O

CASES = ["v32_scale", "v32_noscale", "v64_cfg1", "v32_oob"]


def load(name):
    return np.load(os.path.join(GOLDEN, f"pointcloud_{name}.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fp32(name):
    d = load(name)
    p = torch.tensor(d["points"], requires_grad=True)
    q = torch.tensor(d["q"], requires_grad=True)
    s = torch.tensor(d["scale"], requires_grad=True) if "scale" in d else None
    sil, aux = O.effective_loss_forward(p, q, s, V=int(d["V"]), kernel_size=int(d["ksize"]),
                                        sigma=float(d["sigma"]), mode="R", return_aux=True)
    # index / visibility buffers: bit exact
    assert np.array_equal(aux["inbounds"].numpy(), d["inb"])
    assert np.array_equal(aux["base"].numpy()[d["inb"]], d["base"][d["inb"]].astype(np.int64))
    np.testing.assert_array_equal(aux["coords"].detach().numpy(), d["coords"])
    # same fp32 ops in the same order: the silhouette reproduces to the last few ulps
    np.testing.assert_allclose(sil.detach().numpy(), d["sil"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(aux["occupancy"].detach().numpy()[:, ::7, ::5, ::3], d["occ_probe"], atol=1e-6)
    loss = (sil * torch.tensor(d["wts"])).sum()
    grads = torch.autograd.grad(loss, [p, q] + ([s] if s is not None else []))
    for g, key in zip(grads, ["d_points", "d_q", "d_scale"]):
        ref = d[key]
        assert np.abs(g.numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_mode_p_weights_partition_unity():
    torch.manual_seed(0)
    c = (torch.rand(2, 500, 3, dtype=torch.float64) - 0.5) * 0.9
    g = (31) * (c + 0.5)
    occ_sum = 0.0
    # unclamped total mass = number of in-bounds points
    B, N, _ = c.shape
    grid = torch.zeros(B, 32, 32, 32, dtype=torch.float64)
    f = g.floor()
    r = g - f
    inb = O.inbounds(c)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                w = (r[..., 0] if i else 1 - r[..., 0]) * (r[..., 1] if j else 1 - r[..., 1]) * \
                    (r[..., 2] if k else 1 - r[..., 2])
                occ_sum += w[inb].sum()
    assert abs(float(occ_sum) - int(inb.sum())) < 1e-9


def test_silhouette_is_one_minus_transmittance_mode_p():
    torch.manual_seed(1)
    vox = torch.rand(2, 16, 8, 8, dtype=torch.float64)
    sil = O.silhouette_from_voxels(vox, "P")
    o = vox.clamp(O.TERM_EPS, 1 - O.TERM_EPS)
    expect = (1 - (1 - o).prod(dim=1)).flip(1)
    assert torch.allclose(sil, expect, atol=1e-12)


def test_reference_kernel_overflows_for_small_sigma():
    # exp(+x^2/2s^2) overflows fp32 below sigma ~0.75 -> NaN taps (kept in mode R)
    assert torch.isnan(O.kernel_1d(0.5, 21, "R")).any()
    assert torch.isfinite(O.kernel_1d(0.5, 21, "P")).all()


def test_candidate_min_loss():
    torch.manual_seed(2)
    proj = torch.rand(6, 8, 8)
    masks = torch.rand(2, 8, 8)
    loss, idx = O.candidate_min_loss(proj, masks, 3)
    per = ((proj.view(2, 3, 8, 8) - masks[:, None]) ** 2).sum((2, 3))
    assert torch.equal(idx, per.argmin(1))
    assert torch.allclose(loss, per.min(1).values.sum() / 2)
