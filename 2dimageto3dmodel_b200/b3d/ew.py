"""One-pass NHWC helper ops between the convolutions (libb3d csrc/ew_kernels.cu), with autograd."""
import os

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


class _FoldRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_nhwc, kh, pad_y, Cp):
        x = dev(x_nhwc.detach(), "x")
        N, H, W, C = x.shape
        out = torch.empty(N, H + 2 * pad_y - kh + 1, W, Cp, device=x.device, dtype=torch.float32)
        check(lib.b3d_fold_rows_fwd(ptr(x), ptr(out), N, H, W, C, kh, pad_y, Cp, stream_ptr(x)))
        ctx.cfg = (kh, pad_y, Cp, x.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        kh, pad_y, Cp, shape = ctx.cfg
        N, H, W, C = shape
        g = dev(g, "grad")
        gx = torch.empty(shape, device=g.device, dtype=torch.float32)
        check(lib.b3d_fold_rows_bwd(ptr(g), ptr(gx), N, H, W, C, kh, pad_y, Cp, stream_ptr(g)))
        return gx, None, None, None


def fold_rows(x_nhwc, kh, pad_y, Cp):
    """[N,H,W,C] -> [N,H + 2*pad_y - kh + 1,W,Cp] with out[..., r*C + c] = x[n, y + r - pad_y, x, c] (zeros elsewhere)."""
    return _FoldRows.apply(x_nhwc, int(kh), int(pad_y), int(Cp))


def pad_x(x_nchw, amount, mode):
    """Padding along x of a logically-NCHW (channels-last) tensor; returns the same kind of tensor."""
    if amount == 0:
        return x_nchw
    if x_nchw.shape[1] % 4:         # odd channel counts (raw RGBA+... inputs): plain torch
        if mode == REPLICATE:
            return torch.nn.functional.pad(x_nchw, (amount, amount, 0, 0), mode='replicate')
        return torch.cat((x_nchw[..., -amount:], x_nchw, x_nchw[..., :amount]), dim=3)
    return _PadX.apply(x_nchw.permute(0, 2, 3, 1), int(amount), int(mode)).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------------------------
# fused conditional-batch-norm -> LeakyReLU -> (+ residual) -> (LeakyReLU) -> x2 nearest upsample -> replicate pad
# ------------------------------------------------------------------------------------------------------------------
def _dist_world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class _CBNActPad(torch.autograd.Function):
    """y [N,H,W,C] (conv output, NHWC), gamma_t = 1 + gamma [N,C], beta [N,C], mean / invstd [C] (already computed)
    -> out [N, up*H, up*W + 2*pad, C].  `skip` (optional) [N,H,Ws,C] read at pixel offset skip_off.
    `batch_stats`: the statistics were computed from y (training mode), so dy carries the batch-norm coupling terms."""

    @staticmethod
    def forward(ctx, y, gamma_t, beta, mean, invstd, skip, skip_off, up, pad, post_leaky, batch_stats, sync):
        y = dev(y.detach(), "y")
        N, H, W, C = y.shape
        gt, bt = gamma_t.detach().contiguous(), beta.detach().contiguous()
        scale = (invstd[None, :] * gt).contiguous()
        shift = (bt - mean[None, :] * scale).contiguous()
        sk = dev(skip.detach(), "skip") if skip is not None else None
        pitch = sk.shape[2] if sk is not None else 0
        out = torch.empty(N, up * H, up * W + 2 * pad, C, device=y.device, dtype=torch.float32)
        check(lib.b3d_cbn_act_fwd(ptr(y), ptr(scale), ptr(shift), ptr(sk), pitch, skip_off, ptr(out), N, H, W, C, up, pad, 0.2,
                                  int(post_leaky), stream_ptr(y)))
        ctx.save_for_backward(y, gt, scale, shift, mean, invstd, sk if sk is not None else torch.empty(0))
        ctx.cfg = (skip_off, up, pad, post_leaky, batch_stats, sync, sk is not None, skip.shape if skip is not None else None)
        return out

    @staticmethod
    def backward(ctx, gout):
        y, gt, scale, shift, mean, invstd, sk = ctx.saved_tensors
        skip_off, up, pad, post_leaky, batch_stats, sync, has_skip, skip_shape = ctx.cfg
        N, H, W, C = y.shape
        gout = dev(gout, "grad")
        st = stream_ptr(y)
        ga = torch.empty_like(y)
        gskip, gpitch = None, 0
        if has_skip and ctx.needs_input_grad[5]:
            gpitch = skip_shape[2]
            gskip = torch.zeros(skip_shape, device=y.device) if gpitch != W else torch.empty(skip_shape, device=y.device)
        S1 = torch.empty(N, C, device=y.device)
        S2 = torch.empty(N, C, device=y.device)
        check(lib.b3d_cbn_act_bwd1(ptr(gout), ptr(y), ptr(scale), ptr(shift), ptr(sk) if has_skip else None,
                                   sk.shape[2] if has_skip else 0, skip_off, ptr(mean), ptr(invstd), ptr(ga), ptr(gskip), gpitch,
                                   skip_off, ptr(S1), ptr(S2), N, H, W, C, up, pad, 0.2, int(post_leaky), st))
        dbeta, dgamma = S1, S2
        if batch_stats:
            red = torch.stack(((gt * S1).sum(0), (gt * S2).sum(0)))            # [2, C]
            M = N * H * W
            if sync:
                import torch.distributed as dist
                dist.all_reduce(red)
                M *= dist.get_world_size()
            red = (red / M).contiguous()
            m1, m2 = red[0].contiguous(), red[1].contiguous()
        else:
            m1 = m2 = torch.zeros(C, device=y.device)
        check(lib.b3d_cbn_act_bwd2(ptr(ga), ptr(y), ptr(gt), ptr(mean), ptr(invstd), ptr(m1), ptr(m2), N, H, W, C, st))
        return ga, dgamma, dbeta, None, None, gskip, None, None, None, None, None, None


_BN_STATS_IMPL = os.environ.get("B3D_BN_STATS", "torch")


def bn_stats(y_nhwc, eps, impl=None):
    """(mean, invstd) per channel of an NHWC tensor = torch.batch_norm_stats on the NCHW view.  impl "b3d" = the one-pass
    libb3d kernel, "torch" = the stock op (default: B3D_BN_STATS, else "torch" — measured on B200: the stock channels-last
    kernel is as fast inside the training step)."""
    y = dev(y_nhwc, "y")
    C = y.shape[-1]
    if C % 4 or 256 % (C // 4) or (impl or _BN_STATS_IMPL) != "b3d":     # odd channel counts: always the stock op
        return torch.batch_norm_stats(y.permute(0, 3, 1, 2), eps)
    mean = torch.empty(C, device=y.device, dtype=torch.float32)
    invstd = torch.empty_like(mean)
    ws = torch.empty(2 * C, device=y.device, dtype=torch.float64)
    check(lib.b3d_bn_stats(ptr(y), y.numel() // C, C, float(eps), ptr(mean), ptr(invstd), ptr(ws), stream_ptr(y)))
    return mean, invstd


def cbn_act_pad(y_nchw, cbn, z, skip_nchw=None, skip_off=0, up=1, pad=1, post_leaky=False):
    """ConditionalBatchNorm2d(y, z) -> LeakyReLU(0.2) [-> + skip] [-> LeakyReLU] [-> x2 upsample] -> replicate pad, fused.
    `cbn` is a models.gan.ConditionalBatchNorm2d whose .norm is a (Synchronized)BatchNorm2d without affine; statistics and
    running buffers follow F.batch_norm (single process) or the reference's SyncBN formulas (torch.distributed)."""
    bn = cbn.norm
    y = y_nchw.permute(0, 2, 3, 1)
    C = y.shape[3]
    gamma_t = 1 + cbn.fc_gamma(z)
    beta = cbn.fc_beta(z)
    sync = False
    if bn.training:
        yv = y_nchw.detach()
        n = yv.numel() // C
        mean, invstd = bn_stats(y.detach(), bn.eps)
        var_b = invstd.pow(-2) - bn.eps
        if _dist_world() > 1 and bn.__class__.__name__.startswith("Synchronized"):
            import torch.distributed as dist
            sync = True
            stats = torch.stack((mean * n, (var_b + mean * mean) * n))
            dist.all_reduce(stats)
            n = n * dist.get_world_size()
            mean = stats[0] / n
            var_b = stats[1] / n - mean * mean
            invstd = var_b.clamp(min=bn.eps).pow(-0.5)                       # sync_batchnorm/batchnorm.py:150
        if bn.track_running_stats:
            with torch.no_grad():
                bn.running_mean.mul_(1 - bn.momentum).add_(mean, alpha=bn.momentum)
                bn.running_var.mul_(1 - bn.momentum).add_(var_b * (n / max(n - 1, 1)), alpha=bn.momentum)
                bn.num_batches_tracked += 1
        batch_stats = True
    else:
        mean, invstd = bn.running_mean, (bn.running_var + bn.eps).rsqrt()
        batch_stats = False
    skip = skip_nchw.permute(0, 2, 3, 1) if skip_nchw is not None else None
    out = _CBNActPad.apply(y, gamma_t, beta, mean.contiguous(), invstd.contiguous(), skip, int(skip_off), int(up), int(pad),
                           bool(post_leaky), batch_stats, sync)
    return out.permute(0, 3, 1, 2)
