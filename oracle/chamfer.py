"""CPU oracle for the chamfer / pairwise nearest-neighbour kernel.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference has no chamfer implementation at all
(`grep -ri chamfer /root/reference` is empty); BASELINE.json's north star asks for the
kernel, so the oracle is the brute-force definition: squared L2, nearest neighbour in both directions
(ties -> lowest index), mean over each set, summed.  The only pairwise-NN site in the reference is the
mirror-vertex search of rendering/mesh_template.py:33-39 (argmin of L2), which the same kernel serves.
Distances are formed as ((dx*dx + dy*dy) + dz*dz) in fp32 so the index outputs can be compared bit for bit.
"""
import torch


def pairwise_sqdist(a, b):
    d = a[:, :, None, :] - b[:, None, :, :]
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def nearest(a, b, chunk=1024):
    """For every a_i the squared distance to, and index of, its nearest b_j. a [B,N,3], b [B,M,3]."""
    dist, idx = [], []
    for s in range(0, a.shape[1], chunk):
        d = pairwise_sqdist(a[:, s:s + chunk], b)
        m, i = d.min(dim=2)
        dist.append(m)
        idx.append(i)
    return torch.cat(dist, 1), torch.cat(idx, 1)


def chamfer(a, b):
    """-> (loss [B], dist_ab [B,N], idx_ab [B,N], dist_ba [B,M], idx_ba [B,M])."""
    dab, iab = nearest(a, b)
    dba, iba = nearest(b, a)
    return dab.mean(1) + dba.mean(1), dab, iab, dba, iba
