"""Parity at the shapes and through the kernel VARIANTS that bench.py's cfg3 / cfg2 workloads dispatch.

Round-1 gap (VERDICT "What's weak" #1/#2): the widest-tile / stacked / halo-staged convolution kernels are only
selected at batch 32 (generator) / 64 (discriminators), sizes no other test reaches.  Every layer of the cfg3 GAN
(models/gan.py:57-65, :163-177, :294-302, :359, :364 at 256^2, nd = 2) runs here at its real shape: forward, input
gradient and weight gradient against torch fp32 convolutions with TF32 off (tolerance 4e-3 of the largest magnitude,
the tf32 product error class — see tests/test_conv_gpu.py), and the kernel template instance that ran is read back
through b3d_last_variant() so the list of exercised variants is asserted, not assumed.

Also: the point-cloud path at cfg2's size (B=16, N=8000, V=128) and the mesh path at 256^2 / 960 faces / 128^2
texture against the ORACLES on sample slices (samples are independent), not just properties."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import mesh as M
from oracle import pointcloud as O

pytestmark = pytest.mark.gpu
TOL = 4e-3
DEV = "cuda:0"
SEEN = set()

B = 32
# name, N, Cin, H, W (x-padded input), Cout, k, pad_y, stride, x_crop
G_LAYERS = [
    ("G.blk1.conv", B, 512, 8, 6, 512, 3, 1, 1, 0),
    ("G.blk2.short", B, 512, 16, 10, 256, 1, 0, 1, 1),
    ("G.blk2.conv1", B, 512, 16, 10, 256, 3, 1, 1, 0),
    ("G.blk2.conv2", B, 256, 16, 10, 256, 3, 1, 1, 0),
    ("G.blk3a.conv", B, 256, 32, 18, 256, 3, 1, 1, 0),
    ("G.blk4.short", B, 256, 64, 34, 128, 1, 0, 1, 1),
    ("G.blk4.conv1", B, 256, 64, 34, 128, 3, 1, 1, 0),
    ("G.blk4.conv2", B, 128, 64, 34, 128, 3, 1, 1, 0),
    ("G.blk5.conv", B, 128, 128, 66, 128, 3, 1, 1, 0),
    ("G.blk6.short", B, 128, 256, 130, 64, 1, 0, 1, 1),
    ("G.blk6.conv1", B, 128, 256, 130, 64, 3, 1, 1, 0),
    ("G.blk6.conv2", B, 64, 256, 130, 64, 3, 1, 1, 0),
    ("G.conv_final", B, 64, 256, 132, 3, 5, 2, 1, 0),
    ("G.blk3_mesh.conv1", B, 256, 32, 18, 64, 3, 1, 1, 0),
    ("G.conv_mesh", B, 64, 32, 20, 3, 5, 2, 1, 0),
]
D_LAYERS = [
    ("D1.conv1", 8, 256, 260, 64, 5, 2, 1),
    ("D1.conv2", 64, 256, 258, 128, 4, 1, 2),
    ("D1.conv3", 128, 128, 130, 256, 4, 1, 2),
    ("D1.conv4", 256, 64, 66, 512, 4, 1, 2),
    ("D1.conv5", 512, 32, 36, 1, 5, 2, 1),
    ("D2.conv1", 11, 32, 36, 64, 5, 2, 1),
    ("D2.conv2", 64, 32, 34, 128, 4, 1, 2),
    ("D2.conv3", 128, 16, 18, 256, 4, 1, 2),
    ("D2.conv4", 256, 8, 12, 1, 5, 2, 1),
]
CASES = G_LAYERS + [(n, nb, ci, h, w, co, k, py, st, 0) for nb in (B, 2 * B) for (n, ci, h, w, co, k, py, st) in D_LAYERS]


def ref_conv(x, w, b, pad_y, stride):
    """fp64 convolution (no TF32 / FFT / Winograd error on the reference side), returned as fp32."""
    return torch.nn.functional.conv2d(x.double(), w.double(), b.double() if b is not None else None, stride=stride,
                                      padding=(pad_y, 0)).float()


@pytest.mark.parametrize("name,N,Cin,H,W,Cout,k,pad_y,stride,x_crop", CASES, ids=[f"{c[0]}-N{c[1]}" for c in CASES])
def test_layer_at_bench_shape(name, N, Cin, H, W, Cout, k, pad_y, stride, x_crop):
    import b3d.conv as C
    g = torch.Generator().manual_seed(sum(map(ord, name)) + N)
    x0 = torch.randn(N, Cin, H, W, generator=g).to(DEV)
    w0 = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(DEV)
    b0 = torch.randn(Cout, generator=g).to(DEV) if name.startswith("D") or "final" in name or "conv_mesh" in name else None
    res = {}
    for impl in ("torch", "b3d"):
        x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True) if b0 is not None else None
        if impl == "torch":
            y = ref_conv(x[..., x_crop:W - x_crop] if x_crop else x, w, b, pad_y, stride)
        else:
            C.VARIANT_LOG = []
            y = C.conv2d(x.contiguous(memory_format=torch.channels_last), w, b, pad_y, stride, x_crop=x_crop)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
        grads = torch.autograd.grad(y, [x, w] + ([b] if b is not None else []), gy)
        res[impl] = [y.detach().float()] + [g_.float() for g_ in grads]
        del x, w, y, gy, grads
    torch.cuda.synchronize()
    SEEN.update(C.VARIANT_LOG)
    C.VARIANT_LOG = None
    for a, r, what in zip(res["b3d"], res["torch"], ("y", "dx", "dw", "db")):
        assert a.shape == r.shape
        err, ref = float((a - r).abs().max()), float(r.abs().max())
        assert err <= TOL * ref, (name, what, err, ref)


@pytest.mark.parametrize("N,Cin,H,W,Cout", [(32, 64, 256, 130, 64),      # G.blk6.conv2: row-window kernel
                                             (32, 128, 128, 66, 128),     # G.blk5: persistent, stacked 128-wide tiles
                                             (32, 256, 32, 18, 256),      # G.blk3a: 256-wide tiles
                                             (3, 512, 8, 6, 512),         # blk1: tiny maps, several images per tile
                                             (2, 128, 16, 19, 96)])       # ragged channel count / width
def test_epilogue_statistics(N, Cin, H, W, Cout):
    """BatchNorm statistics accumulated by the conv epilogue (b3d_conv2d_tf32 `stats`) == sums of the stored output."""
    import b3d.conv as C
    from b3d.bank import WeightBank
    from models.gan import TCConv2d
    torch.manual_seed(Cin + Cout)
    conv = TCConv2d(Cin, Cout, 3, padding=(1, 0), bias=False).to(DEV)
    W_ = WeightBank({"c": conv}).forward(True)
    x = torch.randn(N, Cin, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    stats = torch.zeros(2 * Cout, device=DEV, dtype=torch.float64)
    C.VARIANT_LOG = []
    with torch.no_grad():
        y = C.conv2d_banked(x, W_["c"], pad_y=1, stats=stats)
    torch.cuda.synchronize()
    SEEN.update(C.VARIANT_LOG)
    C.VARIANT_LOG = None
    yd = y.double()
    ref = torch.cat((yd.sum(dim=(0, 2, 3)), (yd * yd).sum(dim=(0, 2, 3))))
    scale = float(yd.abs().sum(dim=(0, 2, 3)).max())
    assert float((stats[:Cout] - ref[:Cout]).abs().max()) <= 2e-6 * scale
    assert float((stats[Cout:] - ref[Cout:]).abs().max()) <= 2e-6 * float(ref[Cout:].max())


@pytest.mark.parametrize("N,H,W,need_dx", [(64, 256, 260, False), (32, 256, 260, True), (32, 128, 132, True)])
def test_stem_fold_on_the_fly(N, H, W, need_dx):
    """Discriminator stem (8 -> 64 channels, 5x5, models/gan.py:163-166) through the banked path with FROZEN weights (the
    generator step): the forward kernel folds the five vertical taps into the K dimension ON THE FLY from the raw 8-channel
    input (TMA boxes of 4 rows x 8 channels, 32-byte swizzle) — forward and input gradient against an fp64 convolution."""
    import b3d.conv as C
    from b3d.bank import WeightBank
    from models.gan import TCConv2d
    torch.manual_seed(N + W)
    conv = TCConv2d(8, 64, 5, padding=(2, 0)).to(DEV)
    for p_ in conv.parameters():
        p_.requires_grad_(False)
    x0 = torch.randn(N, 8, H, W, device=DEV)
    res = {}
    for impl in ("ref", "b3d"):
        x = x0.clone().requires_grad_(need_dx)
        if impl == "ref":
            y = ref_conv(x, conv.weight, conv.bias, 2, 1)
        else:
            W_ = WeightBank({"c": conv}, fold=("c",), round_tf32=False).forward(True)
            C.VARIANT_LOG = []
            y = C.conv2d_banked(x, W_["c"], pad_y=2)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
        grads = torch.autograd.grad(y, [x], gy) if need_dx else []
        res[impl] = [y.detach().float()] + [g_.float() for g_ in grads]
    torch.cuda.synchronize()
    SEEN.update(C.VARIANT_LOG)
    C.VARIANT_LOG = None
    for a, r in zip(res["b3d"], res["ref"]):
        assert a.shape == r.shape
        err, ref = float((a - r).abs().max()), float(r.abs().max())
        assert err <= TOL * ref, (tuple(a.shape), err, ref)


def test_every_dispatched_variant_was_exercised():
    """The kernel instances cfg3 dispatches at batch 32 / 64 (profiles/r2_launches.md) all ran in the cases above."""
    need = {
        # fprop / dgrad: wide-N, stacked and small persistent tiles, the halo-staged kernel, N-major weights in the dgrad
        "conv_tf32_persistent<256,4,0,1>", "conv_tf32_persistent<128,4,0,2>", "conv_tf32_persistent<64,3,0,4>",
        "conv_tf32_persistent<128,3,0,1>", "conv_tf32_persistent<64,4,0,1>",
        # row-window kernel (tc_conv3.cu): 128-pixel row tiles of the 64-wide layers (3x3, 1x5 / 5x5) and blk6.conv1's dgrad
        "conv_rowwin_tf32<64,3,4,2>", "conv_rowwin_tf32<64,5,4,2>", "conv_rowwin_tf32<128,3,2,2>",     # <64,5,..>: conv_final dgrad
        "conv_rowwin_tf32<64,2,4,2>",                                                                  # D1.conv2 dgrad parity classes
        "conv_rowwin_tf32<16,5,4,2>",                                                                  # conv_final fprop (N = 16 tiles)
        # weight gradients: row-of-taps (T = 3, 5), stride-2 tap pairs (T = 2), single taps, both Cin tile widths
        "wgrad_tf32<128,6,3>", "wgrad_tf32<64,4,3>", "wgrad_tf32<64,8,5>", "wgrad_tf32<128,3,2>", "wgrad_tf32<64,4,2>",
        "wgrad_tf32<128,6,1>", "wgrad_tf32<64,8,1>",
        # 1-3 output channel heads on the CUDA-core kernels
        "conv_thin_fwd<3,2>", "conv_thin_fwd<1,4>", "conv_thin_wgrad_win<3,2>", "conv_thin_wgrad_win<1,4>",     # fwd<3,2>: conv_mesh
    }
    flat = {v for v in SEEN if v.startswith("conv_flat_tf32<128>")}
    assert flat, f"the halo-staged kernel never ran; seen: {sorted(SEEN)}"
    missing = need - SEEN
    assert not missing, f"variants the bench dispatches but no case reached: {sorted(missing)}; seen: {sorted(SEEN)}"


# ----------------------------------------------------------------------------------------------------------------
# point-cloud effective loss at cfg2 size against the oracle (two samples of the batch)
# ----------------------------------------------------------------------------------------------------------------
def test_pointcloud_bench_size_against_oracle():
    from utils.effective_loss_function import EffectiveLossFunction
    Bp, N, V = 16, 8000, 128
    g = torch.Generator().manual_seed(1234)
    pts = (torch.rand(Bp, N, 3, generator=g) * 2 - 1) * 0.45          # bench.py:host_inputs (half on a noisy shell)
    shell = torch.nn.functional.normalize(torch.randn(Bp, N // 2, 3, generator=g), dim=-1)
    pts[:, : N // 2] = shell * (0.33 + 0.01 * torch.randn(Bp, N // 2, 1, generator=g))
    q = torch.randn(Bp, 4, generator=g)
    s = 0.5 + 0.5 * torch.rand(Bp, 1, generator=g)
    wts = torch.rand(Bp, V, V, generator=g)
    p, r, sc = (t.to(DEV).requires_grad_(True) for t in (pts, q, s))
    sil = EffectiveLossFunction(voxel_size=V).to(DEV)(p, r, sc)
    gp, gq, gs = torch.autograd.grad((sil * wts.to(DEV)).sum(), [p, r, sc])
    for i in (0, 11):
        out = {}
        for dt in (torch.float32, torch.float64):
            po, qo, so = (t[i:i + 1].to(dt).requires_grad_(True) for t in (pts, q, s))
            so_ = O.effective_loss_forward(po, qo, so, V=V, kernel_size=21, sigma=3.0, mode="R")
            out[dt] = [so_.detach()] + list(torch.autograd.grad((so_ * wts[i:i + 1].to(dt)).sum(), [po, qo, so]))
        for v, o32, o64, name in zip((sil[i:i + 1], gp[i:i + 1], gq[i:i + 1], gs[i:i + 1]), out[torch.float32],
                                     out[torch.float64], ("sil", "d_points", "d_q", "d_scale")):
            gap = float((o32.double() - o64).abs().max())                # the reference's own fp32 noise (App. A D5)
            tol = 4.0 * gap + 1e-5 * max(1.0, float(o64.abs().max()))
            err = float((v.detach().cpu().double() - o64).abs().max())
            assert err <= tol, f"sample {i} {name}: |cuda - oracle64| = {err:.3e} > {tol:.3e}"


# ----------------------------------------------------------------------------------------------------------------
# mesh render at cfg2 size (256^2, 960 faces, 128^2 texture) against the oracle (one sample of the batch)
# ----------------------------------------------------------------------------------------------------------------
def test_mesh_bench_size_against_oracle():
    from rendering.mesh_template import MeshTemplate
    from rendering.renderer import Renderer
    path = M.write_uvsphere_obj(os.path.join(tempfile.mkdtemp(), "uvsphere_16rings.obj"), rings=16)
    mt, T = MeshTemplate(path, device=DEV), M.TemplateData(M.load_obj(path), path)
    Bm, H = 16, 256
    g = torch.Generator().manual_seed(5)
    mesh_map = torch.randn(Bm, 3, 32, 32, generator=g) * 0.05
    q = torch.nn.functional.normalize(torch.randn(Bm, 4, generator=g), dim=-1)
    s = 0.55 + 0.3 * torch.rand(Bm, 1, generator=g)
    t = (torch.rand(Bm, 3, generator=g) - 0.5) * 0.3
    tex = torch.rand(Bm, 3, 128, 128, generator=g) * 2 - 1
    wi, wa = torch.rand(Bm, H, H, 3, generator=g), torch.rand(Bm, H, H, 1, generator=g)
    vtx0 = M.transform_vertices(M.get_vertex_positions(T, mesh_map), s, t, q)
    vc, tc = vtx0.to(DEV).requires_grad_(True), tex.to(DEV).requires_grad_(True)
    r = Renderer(H, H)
    img, alpha = mt.forward_renderer(r, vc, tc)
    idx = r.last_face_index.cpu()
    gv, gt = torch.autograd.grad((img * wi.to(DEV)).sum() + (alpha * wa.to(DEV)).sum(), [vc, tc])
    for i in (3,):
        vo, to = vtx0[i:i + 1].clone().requires_grad_(True), tex[i:i + 1].clone().requires_grad_(True)
        img_o, alpha_o, idx_o = M.forward_renderer(T, vo, to, H, H)
        gvo, gto = torch.autograd.grad((img_o * wi[i:i + 1]).sum() + (alpha_o * wa[i:i + 1]).sum(), [vo, to])
        nbad = int((idx[i] != idx_o[0]).sum())
        assert nbad == 0, f"face-index buffer differs from the oracle in {nbad} of {idx_o.numel()} pixels"
        # colour: bilinear fetch from a 128-texel texture amplifies the fp32 rounding of the interpolated uv by ~T |d tex|
        assert float((img[i].cpu() - img_o[0]).abs().max()) < 1e-4
        assert float((alpha[i].cpu() - alpha_o[0]).abs().max()) < 2e-5
        assert float((gv[i].cpu() - gvo[0]).abs().max()) < 2e-3 * float(gvo.abs().max())
        assert float((gt[i].cpu() - gto[0]).abs().max()) < 1e-4 * float(gto.abs().max())
