"""Drop-in for the LOSS of /root/reference/code/models/unsupervised_part.py (`UnsupervisedLoss`, :91-143).
The encoder / decoder networks of that file are callers of the hot path and out of scope; the loss is the
arithmetic that consumes the silhouettes `EffectiveLossFunction` renders: a handful of tiny tensor ops, device agnostic.

As written the reference raises AttributeError in training (`self.num_candidates`, :117); this class uses
the constructor's `number_of_pose_predictor_candidates`, which is what the golden test patches into the reference.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from quaternions.operations import QuaternionOperations
from utils.batch_repetition import repeat_tensor_for_each_element_in_batch


def half_resolution_masks(masks):
    """[B,2V,2V] -> [B,V,V]: bilinear, align_corners=True, through a fake batch axis exactly as :108 does."""
    return F.interpolate(input=masks.unsqueeze(0), scale_factor=1 / 2, mode="bilinear", align_corners=True).squeeze()


class UnsupervisedLoss(nn.Module):
    """min over K pose candidates of the silhouette MSE + student-pose loss (:98-143).
    forward(predictions=(projection [B*K,V,V], ensemble_poses [B*K,4], student_poses [B,4]), masks [B,2V,2V], training)."""

    def __init__(self, number_of_pose_predictor_candidates=4, student_weight=20.00):
        super().__init__()
        self.student_weight = student_weight
        self.number_of_pose_predictor_candidates = number_of_pose_predictor_candidates
        self.minimum_indexes = None

    def forward(self, predictions, masks, training):
        projection, *poses = predictions
        masks = half_resolution_masks(masks)
        if not training:
            return dict(projection_loss=F.mse_loss(projection, masks, reduction="sum") / projection.size(0))
        K = self.number_of_pose_predictor_candidates
        ensemble_poses, student_poses = poses
        masks = repeat_tensor_for_each_element_in_batch(torch_tensor=masks, n=K)
        projection_loss = F.mse_loss(projection, masks, reduction="none").sum((1, 2)).view(-1, K)
        minimum_indexes = projection_loss.argmin(dim=-1).detach()
        rows = torch.arange(minimum_indexes.size(0), device=minimum_indexes.device)
        minimum_projection_loss = projection_loss[rows, minimum_indexes].sum() / minimum_indexes.size(0)

        # the student pose is pulled towards the best candidate: 1 - cos^2 of half the relative rotation angle
        best_poses = ensemble_poses.view(-1, K, 4)[rows, minimum_indexes, :].detach()
        ops = QuaternionOperations()
        difference = F.normalize(ops.quaternion_multiplication(q1=best_poses, q2=ops.quaternion_conjugate(q=student_poses)), dim=-1)
        student_loss = (1 - difference[:, 0] ** 2).sum() / minimum_indexes.size(0)

        self.minimum_indexes = minimum_indexes.detach()
        total_loss = minimum_projection_loss + self.student_weight * student_loss
        return dict(projection_loss=minimum_projection_loss, student_loss=student_loss, total_loss=total_loss)
