"""Per-layer timing of the GAN's conv shapes (fprop / dgrad / wgrad), B=32 at 256^2 (G) and 2B for D."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
from b3d.conv import conv2d_nhwc, conv2d_dgrad_nhwc, conv2d_wgrad_nhwc
B = int(os.environ.get("B", 32))
# name, N, Cin, H, Wpadded, Cout, k, pad_y, stride
L = [("G.blk1.conv", B, 512, 8, 6, 512, 3, 1, 1), ("G.blk2.conv1", B, 512, 16, 10, 256, 3, 1, 1), ("G.blk3a.conv", B, 256, 32, 18, 256, 3, 1, 1),
     ("G.blk4.conv1", B, 256, 64, 34, 128, 3, 1, 1), ("G.blk4.conv2", B, 128, 64, 34, 128, 3, 1, 1), ("G.blk5.conv", B, 128, 128, 66, 128, 3, 1, 1),
     ("G.blk6.conv1", B, 128, 256, 130, 64, 3, 1, 1), ("G.blk6.conv2", B, 64, 256, 130, 64, 3, 1, 1), ("G.blk6.short", B, 128, 256, 128, 64, 1, 0, 1),
     ("G.conv_final", B, 64, 256, 132, 3, 5, 2, 1),
     ("D1.c1.kwfold", 2 * B, 64, 256, 256, 64, (5, 1), 2, 1), ("D1.c1.khfold", 2 * B, 64, 256, 260, 64, (1, 5), 0, 1), ("D1.conv2", 2 * B, 64, 256, 258, 128, 4, 1, 2), ("D1.conv3", 2 * B, 128, 128, 130, 256, 4, 1, 2),
     ("D1.conv4", 2 * B, 256, 64, 66, 512, 4, 1, 2), ("D1.conv5", 2 * B, 512, 32, 36, 1, 5, 2, 1)]
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tot = [0, 0, 0]
print(f"{'layer':14s} {'GF':>8s} | {'fprop ms':>9s} {'TF/s':>6s} | {'dgrad ms':>9s} {'TF/s':>6s} | {'wgrad ms':>9s} {'TF/s':>6s}")
for name, N, Cin, H, W, Cout, k, py, st in L:
    kh, kw = k if isinstance(k, tuple) else (k, k)
    x = torch.randn(N, H, W, Cin, device="cuda"); w = torch.randn(Cout, Cin, kh, kw, device="cuda") * 0.05
    y = conv2d_nhwc(x, w, pad_y=py, stride=st)
    gy = torch.randn_like(y)
    gf = 2.0 * y.numel() * Cin * kh * kw / 1e9
    cpad = (-Cout) % 32
    gyp = torch.nn.functional.pad(gy, (0, cpad)); wp = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, cpad))
    a = t(lambda: conv2d_nhwc(x, w, pad_y=py, stride=st))
    b_ = t(lambda: conv2d_dgrad_nhwc(gyp, wp, (H, W), pad_y=py, stride=st))
    c = t(lambda: conv2d_wgrad_nhwc(gyp, x, kh, kw, pad_y=py, stride=st))
    tot[0] += a; tot[1] += b_; tot[2] += c
    print(f"{name:14s} {gf:8.1f} | {a:9.3f} {gf / a:6.1f} | {b_:9.3f} {gf / b_:6.1f} | {c:9.3f} {gf / c:6.1f}")
print("totals ms: fprop %.2f dgrad %.2f wgrad %.2f" % tuple(tot))
