"""Pose application between the mesh template and the renderer, and the silhouette IoU metric — the two helper functions
the reference keeps at script level in run_reconstruction.py (`transform_vertices` :237-252, `mean_iou` :225-231; SURVEY §8
rows a11 / a12), as importable functions.  Device-agnostic torch code (a handful of small elementwise ops in front of the
rasteriser; SURVEY §8f rank 1 lists fusing them with the template deformation as the next kernel)."""
import torch

from rendering.utils import qrot


def transform_vertices(vtx, gt_scale, gt_translation, gt_rot, gt_idx=None, dataset_params=None, optimize_deltas=False,
                       optimize_z0=False):
    """Object space -> camera space: v' = flip_yz(qrot(q, (s + ds) v) + (t + dt)), then the optional perspective
    correction x,y *= (z0 + z/2) / (z0 - z/2) with the learned z0 = 1 + exp(theta).
    vtx [B,P,3], gt_scale [B,1], gt_translation [B,3], gt_rot [B,4] (w,x,y,z); `dataset_params` is a
    models.reconstruction.DatasetParams (needed when optimize_deltas / optimize_z0), `gt_idx` its image indices.
    (The reference's guard against using z0-trained parameters without --optimize_z0, :251, inspects the module's
    __dict__, where nn.Parameters never live, and therefore never fires; it is not reproduced.)"""
    scale_delta, translation_delta = 0, 0
    if optimize_deltas:
        translation_delta, scale_delta = dataset_params(gt_idx, 'deltas')
    vtx = qrot(gt_rot, (gt_scale + scale_delta).unsqueeze(-1) * vtx) + (gt_translation + translation_delta).unsqueeze(1)
    vtx = torch.cat((vtx[..., :1], -vtx[..., 1:]), dim=-1)       # * (1, -1, -1) without a host constant (CUDA-graph safe)
    if optimize_z0:
        z0 = dataset_params(gt_idx, 'z0').unsqueeze(-1)
        z = vtx[:, :, 2:]
        vtx = torch.cat((vtx[:, :, :2] * ((z0 + z / 2) / (z0 - z / 2)), z), dim=2)
    return vtx


def mean_iou(alpha_pred, alpha_real):
    """Mean over the batch of |pred & real| / |pred | real| with both alpha maps thresholded at 0.5 ([B,H,W])."""
    p, r = alpha_pred > 0.5, alpha_real > 0.5
    return torch.mean((p & r).float().sum(dim=[1, 2]) / (p | r).float().sum(dim=[1, 2]))
