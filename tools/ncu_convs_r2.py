"""Representative launches of the round-2 conv kernels for one `ncu --set full` capture (every launch captured once, in order):
  0 G.blk6.conv1 fprop  (conv_rowwin_tf32<64,3,4,2>)         1 G.blk6.conv1 dgrad (conv_rowwin_tf32<128,3,2,2>)
  2 D1.conv1 fprop, folded stem (conv_tf32_persistent<64,3,0,4>)
  3-6+ D1.conv2 dgrad: 4 parity classes (conv_rowwin_tf32<64,2,4,2>) + strips
  then D1.conv3 fprop (persistent<256,4,0,1>), D1.conv3 dgrad classes (persistent<128,4,0,2>), G.blk5 fprop with statistics
  (persistent<128,4,0,2> + BN-statistics epilogue), conv_final fprop (conv_rowwin_tf32<16,5,4,2>), wgrads of blk6.conv2 / D1.conv3."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
from b3d.conv import conv2d_nhwc, conv2d_dgrad_nhwc, conv2d_wgrad_nhwc, conv2d_banked
from b3d.bank import WeightBank
from models.gan import TCConv2d
B = 32
r = lambda *s: torch.randn(*s, device="cuda")
y = conv2d_nhwc(r(B, 256, 130, 128), r(64, 128, 3, 3) * 0.05, pad_y=1)
gx = conv2d_dgrad_nhwc(r(B, 256, 128, 64), r(64, 128, 3, 3) * 0.05, (256, 130), pad_y=1)
y = conv2d_nhwc(r(2 * B, 256, 260, 64), r(64, 64, 1, 5) * 0.05)
gx = conv2d_dgrad_nhwc(r(2 * B, 128, 128, 128), r(128, 64, 4, 4) * 0.05, (256, 258), pad_y=1, stride=2)
y = conv2d_nhwc(r(2 * B, 128, 130, 128), r(256, 128, 4, 4) * 0.05, pad_y=1, stride=2)
gx = conv2d_dgrad_nhwc(r(2 * B, 64, 64, 256), r(256, 128, 4, 4) * 0.05, (128, 130), pad_y=1, stride=2)
conv = TCConv2d(128, 128, 3, padding=(1, 0), bias=False).cuda()
W = WeightBank({"c": conv}).forward(True)
stats = torch.zeros(256, device="cuda", dtype=torch.float64)
with torch.no_grad():
    y = conv2d_banked(r(B, 128, 128, 66).contiguous(memory_format=torch.channels_last), W["c"], pad_y=1, stats=stats)
y = conv2d_nhwc(r(B, 256, 132, 64), r(3, 64, 5, 5) * 0.05, pad_y=2)
gw = conv2d_wgrad_nhwc(r(B, 256, 128, 64), r(B, 256, 130, 64), 3, 3, pad_y=1)
gw = conv2d_wgrad_nhwc(r(2 * B, 64, 64, 256), r(2 * B, 128, 130, 128), 4, 4, pad_y=1, stride=2)
torch.cuda.synchronize()
print("done")
