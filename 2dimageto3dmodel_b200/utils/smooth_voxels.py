"""Drop-in for /root/reference/code/utils/smooth_voxels.py (class VoxelsSmooth, :10-84) on libb3d's dense
blur kernel (csrc/vox_kernels.cu).  `semantics="R"` keeps the reference as written: the Gaussian has a POSITIVE
exponent (:29) and every kernel is applied to the ORIGINAL voxels so only the last (depth) one has an effect
(:72); `semantics="P"` chains the kernels and uses exp(-x^2/2s^2)."""
import torch

from b3d import mode_id
from b3d.pointcloud import blur_axis, scale_clamp, smoothing_taps


class VoxelsSmooth(object):
    def __init__(self, semantics="R"):
        self.semantics = semantics
        mode_id(semantics)

    def separate_kernels(self, std_dev, kernel_size=21):
        """Three views of the same normalised 1-D kernel, shaped to act along x (length), y (column), z (depth)."""
        k = torch.tensor(smoothing_taps(float(std_dev), kernel_size, self.semantics))
        return [k.view(1, 1, 1, 1, -1), k.view(1, 1, 1, -1, 1), k.view(1, 1, -1, 1, 1)]

    def smooth(self, voxels, kernels, scale=None):
        """voxels [B,V,V,V]; kernels as returned by separate_kernels; scale [B,1] or None."""
        out = None
        for kernel in kernels:
            axis = int(max(range(2, 5), key=lambda d: kernel.shape[d])) - 1      # 5-D view dim -> grid axis 1..3
            src = voxels if (mode_id(self.semantics) == 0 or out is None) else out
            out = blur_axis(src, kernel.flatten().tolist(), axis)
        if out is None:
            raise ValueError("smooth() needs at least one kernel (the reference passes kernels=() and crashes, "
                             "SURVEY App. A D2)")
        return scale_clamp(out, scale) if scale is not None else out
