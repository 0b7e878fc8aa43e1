"""CPU oracle for the FID evaluation path (main.py:188-412, utils/fid.py, utils/inception.py).

TEST INFRASTRUCTURE ONLY.

* `calculate_stats` / `calculate_frechet_distance` restate utils/fid.py:27-82 as written (np.mean / np.cov,
  scipy.linalg.sqrtm of the covariance product, eps fallback, imaginary-part check).  PINNED: the reference's own
  functions were run on the same inputs by tests/golden/make_golden_fid.py (-> tests/golden/fid_reference.npz,
  including the principal 128-d block of the real-image CUB statistics shipped with the reference).  One execution
  patch: SciPy >= 1.16 has no `disp=` keyword (utils/fid.py:67 raises TypeError as written).
* `inception_forward` restates the network the reference takes from torchvision (`models.inception_v3`, wrapped by
  utils/inception.py:54-139 into four blocks) as functional torch ops over a state dict with the wrapper's key names.
  PARITY UNPINNED for this part: torchvision is not installed here and its ImageNet weights need a download, so neither
  the architecture restatement nor real activations can be checked against the reference's module; what IS checked is
  the published structure (parameter count 21.8 M for the four blocks, the 64 / 192 / 768 / 2048-channel block outputs
  at 73 / 35 / 17 / 1 pixels the reference's docstring lists) and that the CUDA network equals this restatement.
  BasicConv2d = conv (no bias) -> BatchNorm(eps 1e-3, inference statistics) -> ReLU; the Mixed_5/6/7 `branch_pool` is
  F.avg_pool2d(x, 3, stride 1, padding 1) (count_include_pad default) -> 1x1; Mixed_6a / 7a pool with F.max_pool2d(x, 3, 2).
* `truncated_noise` restates the rejection sampling of main.py:244-253.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import linalg


# ------------------------------------------------------------------------------------------------ statistics (utils/fid.py)
def calculate_stats(act):
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


def truncated_noise(n, dim, sigma, generator=None):
    """main.py:244-253: N(0,1) samples with every |component| <= sigma, by re-drawing the offenders."""
    noise = torch.randn(n, dim, generator=generator)
    while (noise.abs() > sigma).any():
        mask = noise.abs() > sigma
        noise[mask] = torch.randn(int(mask.sum()), generator=generator)
    return noise


# ------------------------------------------------------------------------------------------------ Inception-v3 (torchvision)
def _unit(sd, name, x, stride=1, padding=0):
    x = F.conv2d(x, sd[name + ".conv.weight"], None, stride, padding)
    x = F.batch_norm(x, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], sd[name + ".bn.weight"], sd[name + ".bn.bias"],
                     False, 0.0, 0.001)
    return F.relu(x)


def _mixed_a(sd, p, x):
    b1 = _unit(sd, p + ".branch1x1", x)
    b5 = _unit(sd, p + ".branch5x5_2", _unit(sd, p + ".branch5x5_1", x), padding=2)
    b3 = _unit(sd, p + ".branch3x3dbl_1", x)
    b3 = _unit(sd, p + ".branch3x3dbl_3", _unit(sd, p + ".branch3x3dbl_2", b3, padding=1), padding=1)
    bp = _unit(sd, p + ".branch_pool", F.avg_pool2d(x, 3, 1, 1))
    return torch.cat([b1, b5, b3, bp], 1)


def _mixed_b(sd, p, x):
    b3 = _unit(sd, p + ".branch3x3", x, stride=2)
    bd = _unit(sd, p + ".branch3x3dbl_2", _unit(sd, p + ".branch3x3dbl_1", x), padding=1)
    bd = _unit(sd, p + ".branch3x3dbl_3", bd, stride=2)
    return torch.cat([b3, bd, F.max_pool2d(x, 3, 2)], 1)


def _mixed_c(sd, p, x):
    b1 = _unit(sd, p + ".branch1x1", x)
    b7 = _unit(sd, p + ".branch7x7_1", x)
    b7 = _unit(sd, p + ".branch7x7_2", b7, padding=(0, 3))
    b7 = _unit(sd, p + ".branch7x7_3", b7, padding=(3, 0))
    bd = _unit(sd, p + ".branch7x7dbl_1", x)
    bd = _unit(sd, p + ".branch7x7dbl_2", bd, padding=(3, 0))
    bd = _unit(sd, p + ".branch7x7dbl_3", bd, padding=(0, 3))
    bd = _unit(sd, p + ".branch7x7dbl_4", bd, padding=(3, 0))
    bd = _unit(sd, p + ".branch7x7dbl_5", bd, padding=(0, 3))
    bp = _unit(sd, p + ".branch_pool", F.avg_pool2d(x, 3, 1, 1))
    return torch.cat([b1, b7, bd, bp], 1)


def _mixed_d(sd, p, x):
    b3 = _unit(sd, p + ".branch3x3_2", _unit(sd, p + ".branch3x3_1", x), stride=2)
    b7 = _unit(sd, p + ".branch7x7x3_1", x)
    b7 = _unit(sd, p + ".branch7x7x3_2", b7, padding=(0, 3))
    b7 = _unit(sd, p + ".branch7x7x3_3", b7, padding=(3, 0))
    b7 = _unit(sd, p + ".branch7x7x3_4", b7, stride=2)
    return torch.cat([b3, b7, F.max_pool2d(x, 3, 2)], 1)


def _mixed_e(sd, p, x):
    b1 = _unit(sd, p + ".branch1x1", x)
    b3 = _unit(sd, p + ".branch3x3_1", x)
    b3 = torch.cat([_unit(sd, p + ".branch3x3_2a", b3, padding=(0, 1)), _unit(sd, p + ".branch3x3_2b", b3, padding=(1, 0))], 1)
    bd = _unit(sd, p + ".branch3x3dbl_2", _unit(sd, p + ".branch3x3dbl_1", x), padding=1)
    bd = torch.cat([_unit(sd, p + ".branch3x3dbl_3a", bd, padding=(0, 1)), _unit(sd, p + ".branch3x3dbl_3b", bd, padding=(1, 0))], 1)
    bp = _unit(sd, p + ".branch_pool", F.avg_pool2d(x, 3, 1, 1))
    return torch.cat([b1, b3, bd, bp], 1)


def inception_forward(sd, inp, output_blocks=(3,), resize_input=True, normalize_input=True):
    """utils/inception.py:107-141 on a state dict `sd` (keys blocks.<i>.<j>....); inp [B,3,H,W] in (0,1), any float dtype."""
    sd = {k: v.to(inp.dtype) if v.is_floating_point() else v for k, v in sd.items()}
    x = inp
    if resize_input:
        x = F.interpolate(x, size=(299, 299), mode='bilinear', align_corners=False)
    if normalize_input:
        x = 2 * x - 1
    out = []
    last = max(output_blocks)
    # block 0: Conv2d_1a_3x3 (stride 2), 2a, 2b (padding 1), max pool
    x = _unit(sd, "blocks.0.0", x, stride=2)
    x = _unit(sd, "blocks.0.1", x)
    x = _unit(sd, "blocks.0.2", x, padding=1)
    x = F.max_pool2d(x, 3, 2)
    if 0 in output_blocks:
        out.append(x)
    if last >= 1:                                  # block 1: Conv2d_3b_1x1, 4a_3x3, max pool
        x = _unit(sd, "blocks.1.1", _unit(sd, "blocks.1.0", x))
        x = F.max_pool2d(x, 3, 2)
        if 1 in output_blocks:
            out.append(x)
    if last >= 2:                                  # block 2: Mixed_5b 5c 5d 6a 6b 6c 6d 6e
        for j, fn in enumerate((_mixed_a, _mixed_a, _mixed_a, _mixed_b, _mixed_c, _mixed_c, _mixed_c, _mixed_c)):
            x = fn(sd, f"blocks.2.{j}", x)
        if 2 in output_blocks:
            out.append(x)
    if last >= 3:                                  # block 3: Mixed_7a 7b 7c, global average pool
        for j, fn in enumerate((_mixed_d, _mixed_e, _mixed_e)):
            x = fn(sd, f"blocks.3.{j}", x)
        x = F.adaptive_avg_pool2d(x, (1, 1))
        if 3 in output_blocks:
            out.append(x)
    return out
