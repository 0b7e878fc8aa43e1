"""Chamfer distance over libb3d (squared L2, both directions, mean over each set)."""
import torch

from . import B3DError, check, dev, lib, ptr, stream_ptr


def nearest(query, cand):
    """-> (dist [B,N] fp32, idx [B,N] int32): nearest candidate of every query point (no autograd)."""
    q, c = dev(query.detach(), "query"), dev(cand.detach(), "cand")
    B, N, _ = q.shape
    M = c.shape[1]
    if c.shape[0] != B or q.shape[2] != 3 or c.shape[2] != 3 or M == 0:
        raise B3DError(f"chamfer: bad shapes {tuple(q.shape)} vs {tuple(c.shape)}")
    dist = torch.empty(B, N, device=q.device, dtype=torch.float32)
    idx = torch.empty(B, N, device=q.device, dtype=torch.int32)
    check(lib.b3d_chamfer_nn(ptr(q), ptr(c), B, N, M, ptr(dist), ptr(idx), stream_ptr(q)))
    return dist, idx


class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        dab, iab = nearest(a, b)
        dba, iba = nearest(b, a)
        ctx.save_for_backward(a.detach().contiguous(), b.detach().contiguous(), iab, iba)
        ctx.mark_non_differentiable(iab, iba)
        return dab, iab, dba, iba

    @staticmethod
    def backward(ctx, gab, _i1, gba, _i2):
        a, b, iab, iba = ctx.saved_tensors
        B, N, _ = a.shape
        M = b.shape[1]
        da, db = torch.zeros_like(a), torch.zeros_like(b)
        st = stream_ptr(a)
        if gab is not None:
            check(lib.b3d_chamfer_bwd(ptr(a), ptr(b), ptr(iab), ptr(dev(gab, "grad")), B, N, M, ptr(da), ptr(db), st))
        if gba is not None:
            check(lib.b3d_chamfer_bwd(ptr(b), ptr(a), ptr(iba), ptr(dev(gba, "grad")), B, M, N, ptr(db), ptr(da), st))
        return da, db


def chamfer_distance(a, b, return_indices=False):
    """a [B,N,3], b [B,M,3] -> per-sample loss [B] = mean_i min_j |a_i-b_j|^2 + mean_j min_i |a_i-b_j|^2."""
    dab, iab, dba, iba = _Chamfer.apply(a, b)
    loss = dab.mean(1) + dba.mean(1)
    return (loss, iab, iba) if return_indices else loss
