"""Probe the MN-major B operand inside the (known-good) fprop kernel."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
from b3d.conv import conv2d_nhwc
torch.manual_seed(0)
x = torch.randn(2, 8, 16, 64, device="cuda")
w = torch.randn(128, 64, 1, 1, device="cuda") / 8
ref = conv2d_nhwc(x, w)
out = conv2d_nhwc(x, w, cin_major=True)
torch.cuda.synchronize()
print("1x1 cin_major: max|ref|", float(ref.abs().max()), "err", float((out - ref).abs().max()), "out absmax", float(out.abs().max()))
# structured: w[co, ci] = 1000*ci + co  (small ints exact-ish in tf32?) use one-hot x instead
xo = torch.zeros(1, 8, 16, 64, device="cuda"); xo[0, 0, 0, 5] = 1.0        # pixel 0 has channel 5 hot -> out[0,0,0,co] = w[co,5]
ci = torch.arange(64, device="cuda").float().view(1, 64, 1, 1); co = torch.arange(128, device="cuda").float().view(128, 1, 1, 1)
wc = (ci * 256 + co).contiguous()
o = conv2d_nhwc(xo, wc, cin_major=True); torch.cuda.synchronize()
v = o[0, 0, 0]
print("one-hot ci=5: out[co] decode (ci,co) for co=0..5,32,33,64,127:", [(int(t) // 256, int(t) % 256) for t in v[[0, 1, 2, 3, 4, 5, 32, 33, 64, 127]].tolist()])
