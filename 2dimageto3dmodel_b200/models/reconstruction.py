"""Drop-in for /root/reference/code/models/reconstruction.py: the image -> (texture, displacement map) network that
run_reconstruction.py trains through the renderer, and the per-image pose offsets.

Module tree, parameter and buffer names equal the reference's (`conv1e.weight`, `bn1e.running_mean`, `blk1.conv1.weight`,
`blk4_mesh.shortcut.weight`, `conv_tex.bias`, `fc1_tex.weight`, ... — its checkpoints load with strict=True); every
nn.Conv2d is a models.gan.TCConv2d, i.e. runs on libb3d's tcgen05 / TMA implicit-GEMM kernels (fprop, dgrad and wgrad,
tf32 inputs, fp32 accumulate; the 3-channel 5x5 heads on the thin-head kernels).  The encoder's zero padding along x is
materialised, along y it is the TMA out-of-bounds fill; the decoder's replicate / circular x padding is explicit as in the
reference.  Batch norms, the three linear layers, nearest upsampling and tanh are stock torch ops on channels-last
tensors.  CUDA only: there is no CPU fallback (TCConv2d raises on CPU tensors).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from b3d.ew import CIRCULAR, REPLICATE, pad_x
from models.gan import TCConv2d
from rendering.utils import adjust_poles, symmetrize_texture


class ResBlock(nn.Module):
    """conv3x3 -> BN -> ReLU -> conv3x3 -> BN -> ReLU, plus a 1x1 (or identity) shortcut (reference :7-26)."""

    def __init__(self, ch_in, ch_out, pad_fn):
        super().__init__()
        self.conv1 = TCConv2d(ch_in, ch_in, 3, padding=(1, 0), bias=False)
        self.conv2 = TCConv2d(ch_in, ch_out, 3, padding=(1, 0), bias=False)
        self.bn1 = nn.BatchNorm2d(ch_in)
        self.bn2 = nn.BatchNorm2d(ch_out)
        self.relu = nn.ReLU(inplace=True)
        self.pad_fn = pad_fn
        self.shortcut = TCConv2d(ch_in, ch_out, 1, bias=False) if ch_in != ch_out else (lambda x: x)

    def forward(self, x):
        skip = self.shortcut(x)
        h = self.relu(self.bn1(self.conv1(self.pad_fn(x, 1))))
        h = self.relu(self.bn2(self.conv2(self.pad_fn(h, 1))))
        return h + skip


class ReconstructionNetwork(nn.Module):
    """RGBA image [B,4,128,128] -> (texture [B,3,R,R] in [-1,1], displacement map [B,3,32,32]) (reference :29-134).
    `symmetric=True` predicts the left half of the UV map and mirrors it."""

    def __init__(self, symmetric=True, texture_res=64, mesh_res=32, interpolation_mode='nearest'):
        super().__init__()
        if mesh_res < 32 or texture_res not in (64, 128, 256):
            raise ValueError("mesh_res must be >= 32 and texture_res one of 64 / 128 / 256")
        if interpolation_mode not in ('nearest', 'bilinear'):
            raise ValueError(f"interpolation_mode={interpolation_mode!r}")
        self.symmetric = symmetric
        self.texture_res = texture_res
        mode = REPLICATE if symmetric else CIRCULAR
        self.pad = lambda x, amount: pad_x(x, amount, mode)
        self.relu = nn.ReLU(inplace=True)
        if interpolation_mode == 'nearest':
            self.up = lambda x: F.interpolate(x, scale_factor=2, mode='nearest')
        else:
            self.up = lambda x: F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)

        # encoder: 128 -> 64 -> 32 -> 16 -> 8 -> 4 (stride-2 convolutions), then two linear layers
        self.conv1e = TCConv2d(4, 64, 5, stride=2, padding=2, bias=False)
        self.bn1e = nn.BatchNorm2d(64)
        self.conv2e = TCConv2d(64, 128, 3, stride=2, padding=1, bias=False)
        self.bn2e = nn.BatchNorm2d(128)
        self.conv3e = TCConv2d(128, 256, 3, stride=2, padding=1, bias=False)
        self.bn3e = nn.BatchNorm2d(256)
        self.conv4e = TCConv2d(256, 512, 3, stride=2, padding=1, bias=False)
        self.bn4e = nn.BatchNorm2d(512)
        bottleneck_dim = 256
        self.conv5e = TCConv2d(512, 64, 3, stride=2, padding=1, bias=False)
        self.bn5e = nn.BatchNorm2d(64)
        self.fc1e = nn.Linear(64 * 8 * 8, bottleneck_dim, bias=False)
        self.bnfc1e = nn.BatchNorm1d(bottleneck_dim)
        self.fc3e = nn.Linear(bottleneck_dim, 1024, bias=False)
        self.bnfc3e = nn.BatchNorm1d(1024)

        # texture decoder
        self.base_res_h = 4
        self.base_res_w = 2 if symmetric else 4
        self.fc1_tex = nn.Linear(1024, self.base_res_h * self.base_res_w * 256)
        self.blk1 = ResBlock(256, 512, self.pad)
        self.blk2 = ResBlock(512, 256, self.pad)
        self.blk3 = ResBlock(256, 256, self.pad)
        if texture_res >= 128:
            self.blk3b_tex = ResBlock(256, 256, self.pad)
        if texture_res >= 256:
            self.blk3c_tex = ResBlock(256, 256, self.pad)
        self.blk4_tex = ResBlock(256, 128, self.pad)
        self.blk5_tex = ResBlock(128, 64, self.pad)
        self.conv_tex = TCConv2d(64, 3, 5, padding=(2, 0))

        # mesh head, zero-initialised so that training starts from the undeformed template (reference :101-103)
        self.blk4_mesh = ResBlock(256, 64, self.pad)
        self.conv_mesh = TCConv2d(64, 3, 5, padding=(2, 0))
        with torch.no_grad():
            self.conv_mesh.bias.zero_()
            self.conv_mesh.weight.zero_()
        print('Model parameters: {:.2f}M'.format(sum(p.nelement() for p in self.parameters()) / 1000000))

    def forward(self, x):
        x = x.contiguous(memory_format=torch.channels_last)
        for conv, bn in ((self.conv1e, self.bn1e), (self.conv2e, self.bn2e), (self.conv3e, self.bn3e),
                         (self.conv4e, self.bn4e), (self.conv5e, self.bn5e)):
            x = self.relu(bn(conv(x)))
        x = x.reshape(x.shape[0], -1)                       # flatten in (C, H, W) order, as the reference's .view
        z = self.relu(self.bnfc1e(self.fc1e(x)))
        z = self.relu(self.bnfc3e(self.fc3e(z)))

        bb = self.fc1_tex(z).view(z.shape[0], -1, self.base_res_h, self.base_res_w)
        bb = bb.contiguous(memory_format=torch.channels_last)
        bb = self.up(self.blk1(bb))
        bb = self.up(self.blk2(bb))
        bb = self.up(self.blk3(bb))
        bb_mesh = bb
        if self.texture_res >= 128:
            bb = self.up(self.blk3b_tex(bb))
        if self.texture_res >= 256:
            bb = self.up(self.blk3c_tex(bb))

        mesh_map = self.blk4_mesh(bb_mesh)
        mesh_map = adjust_poles(self.conv_mesh(self.pad(self.relu(mesh_map), 2)))

        tex = self.up(self.blk4_tex(bb))
        tex = self.blk5_tex(tex)
        tex = torch.tanh(self.conv_tex(self.pad(self.relu(tex), 2)))
        if self.symmetric:
            tex = symmetrize_texture(tex)
            mesh_map = symmetrize_texture(mesh_map)
        return tex, mesh_map


class DatasetParams(nn.Module):
    """Learned per-image corrections of the estimated poses (reference :137-179): translation / scale deltas and the
    perspective parameter z0 = 1 + exp(theta).  Indices in [N, 2N) denote the mirrored copy of image i - N: its x
    translation changes sign.  `indices=None` returns the dataset mean (used at test time)."""

    def __init__(self, args, dataset_size):
        super().__init__()
        self.dataset_size = dataset_size
        if args.optimize_deltas:
            self.ds_translation = nn.Parameter(torch.zeros(dataset_size, 2))
            self.ds_scale = nn.Parameter(torch.zeros(dataset_size, 1))
        if args.optimize_z0:
            self.ds_z0 = nn.Parameter(torch.ones(dataset_size, 1))

    def forward(self, indices, mode):
        if mode not in ('deltas', 'z0'):
            raise ValueError(f"mode={mode!r}")
        x_sign = 1
        if indices is not None:
            x_sign = (1 - 2 * (indices // self.dataset_size).float()).unsqueeze(-1)
            indices = indices % self.dataset_size
        pick = (lambda p: p[indices]) if indices is not None else (lambda p: p.mean(dim=0, keepdim=True))
        if mode == 'z0':
            return 1 + torch.exp(pick(self.ds_z0))
        t = pick(self.ds_translation)
        translation_delta = torch.cat((t[:, :1] * x_sign, t[:, 1:2], torch.zeros_like(t[:, :1])), dim=1)
        return translation_delta, pick(self.ds_scale)
