"""Drop-in for /root/reference/code/quaternions/points_quaternions.py (:11-81): points <-> pure quaternions and
rotation of a cloud by a quaternion (q (x) p (x) q*, q normalised first).  The batch-size assert of the reference
(:23, it tests len(batch) == 3) is not reproduced."""
import torch.nn.functional as F

from .operations import QuaternionOperations


class PointsQuaternionsConverter(object):
    @staticmethod
    def points_to_quaternions(xyz_triplet):
        if xyz_triplet.size(-1) != 3:
            raise AssertionError("points must have 3 components")
        return F.pad(xyz_triplet, (1, 0))


class PointsQuaternionsRotator(object):
    @staticmethod
    def rotate_points(xyz_triplet, q, inverse_rotation_direction):
        ops = QuaternionOperations()
        q = F.normalize(q, dim=-1)[:, None, :]
        qc = ops.quaternion_conjugate(q)
        p = PointsQuaternionsConverter.points_to_quaternions(xyz_triplet)
        a, b = (qc, q) if inverse_rotation_direction else (q, qc)
        out = ops.quaternion_multiplication(ops.quaternion_multiplication(a, p), b)
        if out.dim() == 2:
            out = out.unsqueeze(0)
        return out[:, :, 1:4]
