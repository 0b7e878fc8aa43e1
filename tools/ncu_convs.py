"""A few representative conv launches for ncu (blk6.conv1 fprop v2 + v1, blk5 fprop, blk6.conv1 wgrad, D1.conv2 dgrad)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
from b3d.conv import conv2d_nhwc, conv2d_dgrad_nhwc, conv2d_wgrad_nhwc
B = 32
x6 = torch.randn(B, 256, 130, 128, device="cuda"); w6 = torch.randn(64, 128, 3, 3, device="cuda") * 0.05
x5 = torch.randn(B, 128, 66, 128, device="cuda"); w5 = torch.randn(128, 128, 3, 3, device="cuda") * 0.05
for _ in range(2):
    y6 = conv2d_nhwc(x6, w6, pad_y=1)                       # v2 (flat)
    os.environ["B3D_CONV_V1"] = "1"; y6b = conv2d_nhwc(x6, w6, pad_y=1); del os.environ["B3D_CONV_V1"]   # v1
    y5 = conv2d_nhwc(x5, w5, pad_y=1)
    gw = conv2d_wgrad_nhwc(torch.randn_like(y6), x6, 3, 3, pad_y=1)
    gy = torch.randn(2 * B, 128, 128, 128, device="cuda"); wd = torch.randn(128, 64, 4, 4, device="cuda") * 0.05
    gx = conv2d_dgrad_nhwc(gy, wd, (256, 258), pad_y=1, stride=2)
torch.cuda.synchronize()
print("done", float((y6 - y6b).abs().max()))
