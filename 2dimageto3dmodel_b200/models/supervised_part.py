"""Drop-in for the LOSS of /root/reference/code/models/supervised_part.py (`SupervisedLoss`, :66-72):
sum of squared silhouette errors against the half-resolution masks, divided by 2B.  The network of that file is a caller
of the hot path and out of scope."""
import torch.nn as nn
import torch.nn.functional as F

from models.unsupervised_part import half_resolution_masks


class SupervisedLoss(nn.Module):
    def forward(self, projection, masks, **kwargs):
        masks = half_resolution_masks(masks)
        return dict(full_loss=F.mse_loss(input=projection, target=masks, reduction="sum") / (2 * projection.size(0)))
