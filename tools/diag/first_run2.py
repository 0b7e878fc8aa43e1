"""Diagnostic: capture the operands / results of every b3d_pad_leaky_bias_bwd call of the first and second backward."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gan_common as GC          # noqa: E402
from b3d import lib              # noqa: E402
from models import gan           # noqa: E402

res, nd, B = 256, 2, 2
args = GC.make_args(res, nd)
_, D = GC.build(gan, args)
D.cuda().train()
z, c, alpha, tex, mesh = [t.cuda() for t in GC.inputs(args, B=B)]
x0 = torch.cat((tex, alpha), dim=1)
saved = {n: b.clone() for n, b in D.named_buffers()}
for d in (D.d1, D.d2):
    d.disable_act_chain = True

calls = None
real = lib.b3d_pad_leaky_bias_bwd


def view(p, n):
    arr = (ctypes.c_float * n).from_address(0)      # placeholder, replaced below
    return None


def grab(ptr_, n):
    t = torch.empty(n, device="cuda")
    torch.cuda.synchronize()
    ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ptr_, ctypes.c_size_t(4 * n), 3)
    return t


def spy(gout, ypad, gy, gb, rows, W, C, amount, mode, slope, stream):
    torch.cuda.synchronize()
    n_in = rows * (W + 2 * amount) * C
    a, b = grab(gout, n_in), grab(ypad, n_in)
    rc = real(gout, ypad, gy, gb, rows, W, C, amount, mode, slope, stream)
    torch.cuda.synchronize()
    o = grab(gy, rows * W * C)
    calls.append(dict(shape=(rows, W, C, amount), gout=a, y=b, out=o))
    return rc


lib.b3d_pad_leaky_bias_bwd = spy


def run():
    global calls
    calls = []
    with torch.no_grad():
        for n, b in D.named_buffers():
            b.copy_(saved[n])
    D.zero_grad()
    x = x0.clone().requires_grad_(True)
    mm = mesh.clone().requires_grad_(True)
    out, _ = D(x, mm, c)
    g = torch.Generator(device="cuda").manual_seed(3)
    sum((o * torch.randn(o.shape, device=o.device, generator=g)).sum() for o in out).backward()
    torch.cuda.synchronize()
    return calls, {n: p.grad.clone() for n, p in D.named_parameters() if p.grad is not None}


(c1, g1), (c2, g2) = run(), run()
for i, (a, b) in enumerate(zip(c1, c2)):
    def rel(k):
        return float((a[k].double() - b[k].double()).abs().max()) / max(float(b[k].abs().max()), 1e-12)
    rows, W, C, am = a["shape"]
    refs = []
    for r in (a, b):
        g = r["gout"].view(rows, W + 2 * am, C).double()
        y = r["y"].view(rows, W + 2 * am, C).double()
        s = g[:, am:am + W].clone()
        s[:, W - am:] += g[:, :am]
        s[:, :am] += g[:, am + W:]
        m = torch.where(y[:, am:am + W] >= 0, s, s * 0.2)
        refs.append(float((m - r["out"].view(rows, W, C).double()).abs().max()) / max(float(m.abs().max()), 1e-12))
    nflip = int(((a["y"] >= 0) != (b["y"] >= 0)).sum())
    print("call %d shape %s: run1 vs run2 gout %.2e y %.2e out %.2e | kernel vs torch: run1 %.2e run2 %.2e | sign flips in y %d"
          % (i, a["shape"], rel("gout"), rel("y"), rel("out"), refs[0], refs[1], nflip))
for n in g1:
    r = float((g1[n].double() - g2[n].double()).abs().max()) / max(float(g2[n].abs().max()), 1e-12)
    if r > 1e-6:
        print("grad %-28s rel %.2e" % (n, r))
