"""Chamfer loss module for point clouds (BASELINE.json north star: "silhouette/chamfer-style losses").
The reference repository ships no chamfer code; this module follows the common definition (squared L2,
both directions, means) on the shared-memory-blocked nearest-neighbour kernel of libb3d."""
import torch.nn as nn

from b3d.chamfer import chamfer_distance, nearest  # noqa: F401


class ChamferDistance(nn.Module):
    def __init__(self, reduction='mean'):
        super().__init__()
        if reduction not in ('mean', 'sum', 'none'):
            raise ValueError(reduction)
        self.reduction = reduction

    def forward(self, a, b):
        loss = chamfer_distance(a, b)
        return loss.mean() if self.reduction == 'mean' else loss.sum() if self.reduction == 'sum' else loss
