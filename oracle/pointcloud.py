"""CPU oracle for the point-cloud "effective loss" path (pipeline B).

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import it.  The product path (2dimageto3dmodel_b200/) never falls back to it.

Parity status: PINNED for mode "R" against the reference's own Python classes
run in the authoring container (tests/golden/make_golden_pointcloud.py imports
/root/reference/code with execution patches P1-P3 of SURVEY.md App. A and
writes tests/golden/pointcloud_*.npz).  Mode "P" (paper-intended semantics)
has no executable reference; it is pinned only through the properties the
math gives (weights sum to 1, silhouette = 1 - prod(1 - o)).

Restates, in plain torch (dtype generic, float64 capable, autograd gives the
gradients):
  rotate_points            /root/reference/code/quaternions/points_quaternions.py:41-81
  hamilton / conjugate     /root/reference/code/quaternions/operations.py:68-97,120-136
  project                  /root/reference/code/camera/coordinate_system_transformation.py:20-39
  splat                    /root/reference/code/utils/trilinear_interpolation.py:17-74
  kernel_1d / smooth       /root/reference/code/utils/smooth_voxels.py:14-84
  termination / silhouette /root/reference/code/utils/effective_loss_function.py:18-56,79-81

Modes
  "R"  the reference as written (keeps D3 last-kernel-only blur = z axis, D4
       positive Gaussian exponent, D5 `1 - g - floor(g)` weights, D10 epsilon pads)
  "P"  paper-intended: chained xyz blur, exp(-x^2/2s^2), w0 = 1-(g-floor g), zero pads
"""
import torch

FOV = 1.875          # effective_loss_function.py:69
CAM_DIST = 2.0       # effective_loss_function.py:70
BORDER_EPS = 1e-6    # trilinear_interpolation.py:12
TERM_EPS = 1e-5      # effective_loss_function.py:18


def hamilton(a, b):
    """Hamilton product, components (w, x, y, z) on the last axis (operations.py:68-97)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ], dim=-1)


def rotate_points(p, q):
    """q (x) (0,p) (x) q*, q normalised first (points_quaternions.py:53-76).

    F.normalize semantics: q / max(||q||, 1e-12)."""
    n = q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    q = (q / n)[:, None, :]
    qc = q * q.new_tensor([1.0, -1.0, -1.0, -1.0])
    p4 = torch.nn.functional.pad(p, (1, 0))
    return hamilton(hamilton(q, p4), qc)[..., 1:4]


def project(points, q):
    """Columns of a point are (z, y, x); x,y get the perspective factor
    (coordinate_system_transformation.py:25-39)."""
    r = rotate_points(points, q)
    z, y, x = r.unbind(-1)
    x = x * FOV / (z + CAM_DIST)
    y = y * FOV / (z + CAM_DIST)
    return torch.stack([z, y, x], dim=-1)


def inbounds(c):
    """trilinear_interpolation.py:17-25 (strict inequalities, all three coords)."""
    return ((c < 0.5 - BORDER_EPS) & (c > -0.5 + BORDER_EPS)).all(dim=-1)


def splat(c, V, mode="R"):
    """8-corner trilinear scatter-add into [B,V,V,V] (index order b,z,y,x), clamp [0,1].

    Returns (occupancy, base_index[B,N,3] int64, inbounds[B,N] bool).  The
    reference builds eight dense grids, one per corner, and sums them
    (trilinear_interpolation.py:62-74); the sum order per cell differs only in
    float rounding, which the fp tolerance of the tests covers."""
    B, N, _ = c.shape
    g = (V - 1) * (c + 0.5)
    f = g.floor()
    r = g - f
    w0 = (1.0 - g - f) if mode == "R" else (1.0 - r)       # D5 kept in R
    w = [w0, r]
    inb = inbounds(c)
    base = f.long()
    grid = c.new_zeros(B * V * V * V)
    bidx = torch.arange(B, device=c.device)[:, None].expand(B, N)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                upd = w[i][..., 0] * w[j][..., 1] * w[k][..., 2]
                lin = ((bidx * V + base[..., 0] + i) * V + base[..., 1] + j) * V + base[..., 2] + k
                grid = grid.index_add(0, lin[inb], upd[inb])
    return grid.view(B, V, V, V).clamp(0, 1), base, inb


def kernel_1d(sigma, kernel_size=21, mode="R", dtype=torch.float32):
    """smooth_voxels.py:24-31; R keeps the positive exponent (D4)."""
    a, b = (-kernel_size // 2, kernel_size // 2)
    x = torch.arange(a + 1.0, b + 1.0, dtype=dtype)
    s = torch.as_tensor(sigma, dtype=dtype)
    e = x * x / (2 * s * s)
    k = torch.exp(e if mode == "R" else -e)
    return k / k.sum()


def _conv_axis(vox, k, axis):
    """zero-padded 1-D cross-correlation along one of the dims 1(z),2(y),3(x)."""
    pad = k.numel() // 2
    B = vox.shape[0]
    shape = [1, 1, 1, 1, 1]
    shape[axis + 1] = -1
    padding = [0, 0, 0]
    padding[axis - 1] = pad
    w = k.to(vox.dtype).view(shape)
    return torch.nn.functional.conv3d(vox.unsqueeze(1), w, padding=padding).squeeze(1)


def smooth(vox, k, scale=None, mode="R"):
    """smooth_voxels.py:44-84.  R: only the last (depth = z) kernel acts (D3)."""
    if mode == "R":
        out = _conv_axis(vox, k, 1)
    else:
        out = _conv_axis(_conv_axis(_conv_axis(vox, k, 3), k, 2), k, 1)
    if scale is not None:
        out = (out * scale.view(-1, 1, 1, 1)).clamp(0, 1)
    return out


def termination_probs(vox, mode="R"):
    """effective_loss_function.py:18-56 -> [B, V+1, V, V]."""
    o = vox.permute(1, 0, 2, 3).clamp(TERM_EPS, 1.0 - TERM_EPS)
    x = torch.log(1 - o)
    xp = torch.log(o)
    cs = torch.cumsum(x, dim=0)
    pad = torch.full_like(o[:1], TERM_EPS if mode == "R" else 0.0)   # D10 kept in R
    return torch.exp(torch.cat([pad, cs], 0) + torch.cat([xp, pad], 0)).permute(1, 0, 2, 3)


def silhouette_from_voxels(vox, mode="R"):
    """effective_loss_function.py:79-81: sum of the first V termination terms, flipped in y."""
    return termination_probs(vox, mode)[:, :-1].sum(1).flip(1)


def effective_loss_forward(points, q, scale=None, V=64, kernel_size=21, sigma=3.0, mode="R",
                           return_aux=False):
    """EffectiveLossFunction.forward (effective_loss_function.py:58-81) with P1-P3."""
    c = project(points, q)
    occ, base, inb = splat(c, V, mode)
    k = kernel_1d(sigma, kernel_size, mode, dtype=points.dtype)
    sm = smooth(occ, k, scale, mode)
    sil = silhouette_from_voxels(sm, mode)
    if return_aux:
        return sil, dict(coords=c, occupancy=occ, base=base, inbounds=inb, smoothed=sm)
    return sil


def silhouette_mse_sum(sil, mask):
    """Supervised / eval form: sum of squared error / B (supervised_part.py:68-72 uses /(2B)
    over two views; unsupervised_part.py:111 uses reduction='sum' / B)."""
    return ((sil - mask) ** 2).sum() / sil.shape[0]


def downsample_mask_half(masks):
    """unsupervised_part.py:108: bilinear x1/2, align_corners=True, on [B,H,W]."""
    return torch.nn.functional.interpolate(masks.unsqueeze(0), scale_factor=0.5, mode="bilinear",
                                           align_corners=True).squeeze(0)


def candidate_min_loss(projection, masks, K):
    """unsupervised_part.py:113-125 (with D6 fixed: num_candidates = K).

    projection [B*K,V,V], masks [B,V,V] -> (min loss summed / B, argmin idx [B])."""
    m = masks.repeat_interleave(K, dim=0)
    pl = ((projection - m) ** 2).sum((1, 2)).view(-1, K)
    idx = pl.argmin(dim=-1)
    b = torch.arange(idx.numel(), device=idx.device)
    return pl[b, idx].sum() / idx.numel(), idx
