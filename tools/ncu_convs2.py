"""Representative launches of the round-1 final conv kernels for one `ncu --set full` capture (no warm-up loop: every
launch below is captured once, in this order):
  0 D1.conv1 fprop (kh-folded stem)   1 D1.conv2 dgrad, one parity class   2 D1.conv3 fprop   3 G.blk6.conv1 fprop
  4 G.blk6.conv2 wgrad   5 D1.conv3 wgrad   6 conv_final thin fwd   7 conv_final thin wgrad"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
from b3d.conv import conv2d_nhwc, conv2d_dgrad_nhwc, conv2d_wgrad_nhwc
B = 32
r = lambda *s: torch.randn(*s, device="cuda")
y = conv2d_nhwc(r(2 * B, 256, 260, 64), r(64, 64, 1, 5) * 0.05)
gx = conv2d_dgrad_nhwc(r(2 * B, 128, 128, 128), r(128, 64, 4, 4) * 0.05, (256, 258), pad_y=1, stride=2)   # 4 classes (+strips)
y = conv2d_nhwc(r(2 * B, 128, 130, 128), r(256, 128, 4, 4) * 0.05, pad_y=1, stride=2)
y = conv2d_nhwc(r(B, 256, 130, 128), r(64, 128, 3, 3) * 0.05, pad_y=1)
gw = conv2d_wgrad_nhwc(r(B, 256, 128, 64), r(B, 256, 130, 64), 3, 3, pad_y=1)
gw = conv2d_wgrad_nhwc(r(2 * B, 64, 64, 256), r(2 * B, 128, 130, 128), 4, 4, pad_y=1, stride=2)
y = conv2d_nhwc(r(B, 256, 132, 64), r(3, 64, 5, 5) * 0.05, pad_y=2)
gw = conv2d_wgrad_nhwc(r(B, 256, 128, 3), r(B, 256, 132, 64), 5, 5, pad_y=2)
torch.cuda.synchronize()
print("done")
