"""Drop-in for /root/reference/code/quaternions/operations.py (QuaternionOperations, :11-136): small quaternion
algebra on tensors whose last axis holds (w, x, y, z).  On the hot path these products run inside libb3d's
projection kernel; this module keeps the stand-alone call surface (e.g. the student-pose loss,
models/unsupervised_part.py:128-134)."""
import torch


class QuaternionOperations(object):
    def quaternion_addition(self, q1, q2):
        return q1 + q2

    def quaternion_subtraction(self, q1, q2):
        return q1 - q2

    def quaternion_multiplication(self, q1, q2):
        """Hamilton product q1 (x) q2."""
        aw, ax, ay, az = q1.unbind(-1)
        bw, bx, by, bz = q2.unbind(-1)
        return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                            aw * bx + ax * bw + ay * bz - az * by,
                            aw * by + ay * bw + az * bx - ax * bz,
                            aw * bz + az * bw + ax * by - ay * bx), dim=-1)

    def quaternion_square(self, q):
        w, x, y, z = q.unbind(-1)
        return torch.stack((w * w - x * x - y * y - z * z, 2 * w * x, 2 * w * y, 2 * w * z), dim=-1)

    def quaternion_conjugate(self, q):
        return q * q.new_tensor([1.0, -1.0, -1.0, -1.0])
