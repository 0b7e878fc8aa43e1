"""One-pass NHWC helper ops between the convolutions (libb3d csrc/ew_kernels.cu), with autograd."""
import torch

from . import check, dev, lib, ptr, stream_ptr

REPLICATE, CIRCULAR = 0, 1


class _PadX(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_nhwc, amount, mode):
        x = dev(x_nhwc.detach(), "x")
        N, H, W, C = x.shape
        out = torch.empty(N, H, W + 2 * amount, C, device=x.device, dtype=torch.float32)
        check(lib.b3d_pad_x_fwd(ptr(x), ptr(out), N * H, W, C, amount, mode, stream_ptr(x)))
        ctx.cfg = (amount, mode, x.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        amount, mode, shape = ctx.cfg
        N, H, W, C = shape
        g = dev(g, "grad")
        gx = torch.empty(shape, device=g.device, dtype=torch.float32)
        check(lib.b3d_pad_x_bwd(ptr(g), ptr(gx), N * H, W, C, amount, mode, stream_ptr(g)))
        return gx, None, None


def pad_x(x_nchw, amount, mode):
    """Padding along x of a logically-NCHW (channels-last) tensor; returns the same kind of tensor."""
    if amount == 0:
        return x_nchw
    if x_nchw.shape[1] % 4:         # odd channel counts (raw RGBA+... inputs): plain torch
        if mode == REPLICATE:
            return torch.nn.functional.pad(x_nchw, (amount, amount, 0, 0), mode='replicate')
        return torch.cat((x_nchw[..., -amount:], x_nchw, x_nchw[..., :amount]), dim=3)
    return _PadX.apply(x_nchw.permute(0, 2, 3, 1), int(amount), int(mode)).permute(0, 3, 1, 2)
