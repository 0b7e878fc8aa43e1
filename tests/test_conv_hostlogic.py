"""Host-side index logic of the convolution wrappers (CPU, no kernel launches): the stride-2 input-gradient parity classes
and the thin-stem fold are emulated with plain torch ops and compared with autograd / the unfolded convolution."""
import pytest
import torch


@pytest.mark.parametrize("kh,kw,pad_y,H,W", [(4, 4, 1, 16, 18), (3, 3, 1, 16, 18), (5, 5, 2, 12, 16), (3, 3, 1, 9, 11),
                                             (4, 4, 1, 8, 10), (1, 1, 0, 6, 6)])
def test_stride2_parity_classes_reassemble_the_input_gradient(kh, kw, pad_y, H, W):
    from b3d.conv import stride2_classes
    g = torch.Generator().manual_seed(kh * 10 + H)
    N, Cin, Cout = 2, 3, 5
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, kh, kw, generator=g, dtype=torch.float64)
    y = torch.nn.functional.conv2d(x, w, stride=2, padding=(pad_y, 0))
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    ref, = torch.autograd.grad(y, x, gy)
    Hout, Wout = y.shape[2:]
    out = torch.full_like(ref, float('nan'))
    classes = stride2_classes(kh, kw, pad_y, H, W)
    assert len(classes) == 4 and sum(len(c[2]) for c in classes) == kh * kw            # every tap in exactly one class
    for cy, cx, rs, dy, dx, Ha, Wa in classes:
        acc = torch.zeros(N, Cin, Ha, Wa, dtype=torch.float64)
        for (r, s), oy, ox in zip(rs, dy, dx):
            for a in range(Ha):
                for b in range(Wa):
                    yo, xo = a + oy, b + ox
                    if 0 <= yo < Hout and 0 <= xo < Wout:                              # the kernel's TMA zero fill
                        acc[:, :, a, b] += gy[:, :, yo, xo] @ w[:, :, r, s]
        assert out[:, :, cy::2, cx::2].shape == acc.shape
        out[:, :, cy::2, cx::2] = acc
    assert not torch.isnan(out).any()                                                  # the classes tile the input
    assert torch.allclose(out, ref, atol=1e-10)


@pytest.mark.parametrize("Cin,kh,kw,pad_y", [(8, 5, 5, 2), (11, 5, 5, 2), (4, 3, 3, 1)])
def test_kh_fold_equals_the_unfolded_convolution(Cin, kh, kw, pad_y):
    """conv(x, w) == conv(fold_rows(x), fold_kh_weight(w)) with fold_rows restated in torch (the CUDA fold_rows kernel is
    checked against the same restatement in tests/test_conv_gpu.py)."""
    from b3d.conv import fold_kh_weight
    g = torch.Generator().manual_seed(Cin)
    N, H, W, Cout = 2, 9, 12, 6
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, kh, kw, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.conv2d(x, w, padding=(pad_y, 0))
    cpad = (-kh * Cin) % 32
    xn = torch.nn.functional.pad(x.permute(0, 2, 3, 1), (0, 0, 0, 0, pad_y, pad_y))    # NHWC, zero rows
    Hout = H + 2 * pad_y - kh + 1
    folded = torch.cat([xn[:, r:r + Hout] for r in range(kh)] + [xn.new_zeros(N, Hout, W, cpad)], dim=3)
    wf = fold_kh_weight(w, cpad)
    assert wf.shape == (Cout, kh * Cin + cpad, 1, kw)
    out = torch.nn.functional.conv2d(folded.permute(0, 3, 1, 2), wf)
    assert torch.allclose(out, ref, atol=1e-10)


def test_taps_layout_and_thin_dispatch():
    from b3d.conv import _thin, taps_layout
    w = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2)
    t = taps_layout(w)
    assert t.shape == (4, 2, 3) and torch.equal(t[1], w[:, :, 0, 1]) and torch.equal(t[2], w[:, :, 1, 0])
    assert _thin(3, 64, 5, 5, 1) and _thin(1, 512, 5, 5, 1)
    assert not _thin(3, 64, 3, 3, 1) and not _thin(3, 64, 5, 5, 2) and not _thin(8, 64, 5, 5, 1) and not _thin(3, 48, 5, 5, 1)
