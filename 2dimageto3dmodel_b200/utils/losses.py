"""Drop-in for /root/reference/code/utils/losses.py: `loss_flat` (:5-17) on the libb3d kernel and
`GANLoss` (:21-120, the pix2pixHD/SPADE objective with per-sample masked means and per-discriminator
weights) with the reference's constructor and call signature."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from b3d.mesh import flat_loss as _flat_loss


def loss_flat(mesh, norms):
    """Smoothness regulariser: (F/2) * sum over the 3 neighbours of mean (cos(n_f, n_g) - 1)^2."""
    ff = mesh.ff
    if ff.shape[1] != 3:
        ff = ff[:, :3]
    return _flat_loss(norms, ff)


class GANLoss(nn.Module):
    MODES = ('ls', 'original', 'w', 'hinge')

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor, opt=None):
        super().__init__()
        if gan_mode not in self.MODES:
            raise ValueError('Unexpected gan_mode {}'.format(gan_mode))
        self.gan_mode = gan_mode
        self.real_label, self.fake_label = target_real_label, target_fake_label
        self.Tensor, self.opt = tensor, opt

    # kept for API compatibility with the reference (targets are scalars broadcast over the logits)
    def get_target_tensor(self, input, target_is_real):
        return torch.full_like(input, self.real_label if target_is_real else self.fake_label)

    def get_zero_tensor(self, input):
        return torch.zeros_like(input)

    def mean(self, x, mask=None, weight=None):
        """Plain mean, or the batch mean of per-sample mask-weighted means (reference :62-71)."""
        w = 1 if weight is None else weight
        if mask is None:
            return x.mean() * w
        if x.shape != mask.shape:
            raise AssertionError((x.shape, mask.shape))
        per_sample = (x * mask).flatten(1).sum(1) / mask.flatten(1).sum(1)
        return per_sample.mean() * w

    def loss(self, input, target_is_real, for_discriminator=True, mask=None, weight=None):
        mode = self.gan_mode
        if mode == 'original':
            return F.binary_cross_entropy_with_logits(input, self.get_target_tensor(input, target_is_real))
        if mode == 'ls':
            return F.mse_loss(input, self.get_target_tensor(input, target_is_real))
        if mode == 'hinge':
            if not for_discriminator:
                if not target_is_real:
                    raise AssertionError("The generator's hinge loss must be aiming for real")
                return -self.mean(input, mask, weight)
            margin = (input if target_is_real else -input) - 1
            return -self.mean(margin.clamp(max=0), mask, weight)
        return -input.mean() if target_is_real else input.mean()      # wgan

    def __call__(self, input, target_is_real, for_discriminator=True, mask=None, weight=None):
        if not isinstance(input, list):
            return self.loss(input, target_is_real, for_discriminator, mask)
        if mask is not None and (not isinstance(mask, list) or len(mask) != len(input)):
            raise AssertionError("mask must be a list matching the predictions")
        total = 0
        for i, pred in enumerate(input):
            if isinstance(pred, list):
                pred = pred[-1]
            term = self.loss(pred, target_is_real, for_discriminator,
                             None if mask is None else mask[i], None if weight is None else weight[i])
            # a scalar term becomes shape [1] (losses.py:112-114: bs = 1, mean over view(1, -1)), so list inputs return [1] or [B]
            total = total + (term.view(1) if term.dim() == 0 else term.view(term.size(0), -1).mean(dim=1))
        return total / (len(input) if weight is None else sum(weight))
