"""Generate tests/golden/gan_loop_reference.npz by EXECUTING the reference's GAN training iteration (authoring container only):
    python tests/golden/make_golden_gan_loop.py
Taken from main.py's syntax tree and compiled unmodified (main.py itself cannot be imported): `divide_pred` (:414-425),
`update_generator_running_avg` (:429-447), `class ModelWrapper` (:449-526; execution patch D15: its hard-coded
torch.cuda.FloatTensor) and the `for i, data in enumerate(train_loader):` loop of the training section (:672-727) — the loop
body IS the training iteration: G / D alternation (1 : d_steps_per_g), hinge losses through the wrapper, the mesh smoothness
term in the generator step, both Adams, the running-average generator.  The namespace provides the module globals the code
reads: args, the optimisers as :588-589 builds them (Adam, betas (0, 0.9) — written as floats: the int 0 raises on torch >= 2,
SURVEY App. A D14), the reference's own GANLoss / loss_flat, a log function and, standing in for the real networks and the
kaolin-based template, the tiny stand-ins of wrapper_common.py / recon_step_common.py.  `.cuda()` is the identity.
Six iterations at epoch 0 (G D D G D D).  Nothing of the reference is copied into the repository — only its outputs."""
import ast
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import recon_step_common as RS                                 # noqa: E402
import wrapper_common as WC                                    # noqa: E402
from utils.losses import GANLoss, loss_flat                    # noqa: E402  (reference)

torch.Tensor.cuda = lambda self, *a, **k: self                 # D15


def make_args():
    a = WC.make_args(2, 512)
    a.d_steps_per_g, a.mesh_regularization, a.conditional_class, a.tensorboard = 2, 0.0001, True, False
    a.lr_g, a.lr_d = 0.01, 0.04
    return a


def loader(n=6, B=4):
    out = []
    for i in range(n):
        d = WC.inputs(seed=50 + i, B=B)
        out.append({"texture": d["X_tex"], "texture_alpha": d["X_alpha"], "mesh": d["X_mesh"], "class": d["C"]})
    return out


def main():
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    defs = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and
            n.name in ("divide_pred", "update_generator_running_avg", "ModelWrapper")]
    seg = "\n\n".join(ast.get_source_segment(src, n) for n in defs)
    patched = seg.replace("GANLoss(args.loss, tensor=torch.cuda.FloatTensor).cuda()", "GANLoss(args.loss, tensor=torch.FloatTensor)")
    assert len(defs) == 3 and patched != seg
    loop = next(n for n in ast.walk(tree) if isinstance(n, ast.For) and isinstance(n.iter, ast.Call) and
                getattr(n.iter.func, "id", "") == "enumerate" and getattr(n.iter.args[0], "id", "") == "train_loader" and
                any("optimizer_g" in ast.dump(s) for s in n.body))
    args = make_args()
    ns = {"torch": torch, "nn": nn, "math": math, "GANLoss": GANLoss, "args": args}
    exec(compile(patched, "main.py", "exec"), ns)
    gi, D = WC.build()
    trainer = ns["ModelWrapper"](gi, D).train()
    logs = []
    ns.update(trainer=trainer, generator=trainer.generator, generator_running_avg=trainer.generator_running_avg,
              optimizer_g=optim.Adam(trainer.generator.parameters(), lr=args.lr_g, betas=(0.0, 0.9)),
              optimizer_d=optim.Adam(trainer.discriminator.parameters(), lr=args.lr_d, betas=(0.0, 0.9)),
              mesh_template=RS.Template(map_size=8), loss_flat=loss_flat, use_mesh=True, train_loader=loader(),
              d_fake_curve=[0], d_real_curve=[0], g_curve=[0], flat_curve=[0], total_it=0, epoch=0, time=time, log=logs.append)
    torch.manual_seed(77)                                       # the wrapper draws its noise with torch.randn
    exec(compile(ast.Module(body=[loop], type_ignores=[]), "main.py", "exec"), ns)
    assert ns["total_it"] == 6 and len(ns["g_curve"]) == 3 and len(ns["d_fake_curve"]) == 5
    out = {"g_curve": np.array(ns["g_curve"]), "d_fake_curve": np.array(ns["d_fake_curve"]), "d_real_curve": np.array(ns["d_real_curve"]),
           "flat_curve": np.array(ns["flat_curve"]), "log0": np.array(logs[0])}
    for name, mod in (("g", trainer.generator), ("avg", trainer.generator_running_avg), ("d", trainer.discriminator)):
        for k, v in mod.state_dict().items():
            out[f"{name}.{k}"] = v.numpy().copy()
    p = os.path.join(HERE, "gan_loop_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes;", {k: out[k].tolist() for k in ("g_curve", "d_fake_curve", "d_real_curve", "flat_curve")})


if __name__ == "__main__":
    main()
