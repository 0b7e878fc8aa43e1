"""Diagnostic: run the MultiScaleDiscriminator backward several times (chain off / on) and print which tensors differ."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gan_common as GC          # noqa: E402
from models import gan           # noqa: E402

res, nd, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
args = GC.make_args(res, nd)
_, D = GC.build(gan, args)
D.cuda().train()
z, c, alpha, tex, mesh = [t.cuda() for t in GC.inputs(args, B=B)]
x0 = torch.cat((tex, alpha), dim=1)
saved = {n: b.clone() for n, b in D.named_buffers()}
discs = [getattr(D, n) for n in ("d1", "d2", "d3") if hasattr(D, n)]


def run(chain):
    with torch.no_grad():
        for n, b in D.named_buffers():
            b.copy_(saved[n])
    for d in discs:
        d.disable_act_chain = not chain
    D.zero_grad()
    x = x0.clone().requires_grad_(True)
    mm = mesh.clone().requires_grad_(True)
    out, _ = D(x, mm, c)
    g = torch.Generator(device="cuda").manual_seed(3)
    sum((o * torch.randn(o.shape, device=o.device, generator=g)).sum() for o in out).backward()
    torch.cuda.synchronize()
    r = {"out%d" % i: o.detach().clone() for i, o in enumerate(out)}
    r.update({n: p.grad.clone() for n, p in D.named_parameters() if p.grad is not None})
    r["x.grad"], r["mesh.grad"] = x.grad.clone(), mm.grad.clone()
    return r


def diff(a, b, tag):
    rows = []
    for n in a:
        rel = float((a[n].double() - b[n].double()).abs().max()) / max(float(b[n].abs().max()), 1e-12)
        if rel > 1e-5:
            rows.append("%s=%.2e" % (n, rel))
    print(tag, "differ:", " ".join(rows) if rows else "none (<= 1e-5)")


A, A2, C, C2 = run(False), run(False), run(True), run(True)
diff(A2, A, "off vs off")
diff(C2, C, "on  vs on ")
diff(C, A, "on  vs off")
