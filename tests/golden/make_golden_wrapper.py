"""Generate tests/golden/wrapper_reference.npz by EXECUTING the reference's GAN step logic (authoring container only):
    python tests/golden/make_golden_wrapper.py
main.py cannot be imported (argparse, datasets and .cuda() at module scope), so `divide_pred` (:414-425),
`update_generator_running_avg` (:429-447) and `class ModelWrapper` (:449-526) are taken from its syntax tree and compiled
into a namespace that provides the module globals they read (args, torch, nn, math, the reference's utils.losses.GANLoss,
generator / generator_running_avg).  One execution patch: the source segment of ModelWrapper.__init__
hard-codes `tensor=torch.cuda.FloatTensor).cuda()`, rewritten to the CPU tensor type.  The networks are the tiny stand-ins of
wrapper_common.py (same call signatures as models/gan.py) — what is pinned is the WRAPPER: masking / concatenation of fake and
real batches, the discriminator weights [2, 1] at 512^2 with two discriminators, divide_pred, the three modes, and the
epoch-dependent running-average update.  Nothing of the reference is copied into the repository — only its outputs."""
import ast
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import wrapper_common as WC                                   # noqa: E402
from utils.losses import GANLoss                              # noqa: E402  (reference)


def reference_namespace(args):
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and
              n.name in ("divide_pred", "update_generator_running_avg", "ModelWrapper")]
    assert [n.name for n in wanted] == ["divide_pred", "update_generator_running_avg", "ModelWrapper"]
    seg = "\n\n".join(ast.get_source_segment(src, n) for n in wanted)
    patched = seg.replace("GANLoss(args.loss, tensor=torch.cuda.FloatTensor).cuda()", "GANLoss(args.loss, tensor=torch.FloatTensor)")
    assert patched != seg and ".cuda()" not in patched
    ns = {"torch": torch, "nn": nn, "math": math, "GANLoss": GANLoss, "args": args}
    exec(compile(patched, "main.py", "exec"), ns)
    return ns


def main():
    out = {}
    for tag, nd, res in (("w21", 2, 512), ("unw", 2, 256), ("nd3", 3, 512)):
        args = WC.make_args(nd, res)
        ns = reference_namespace(args)
        gi, D = WC.build()
        mw = ns["ModelWrapper"](gi, D).train()
        d = WC.inputs()
        loss, tex, mesh = mw('g', None, d["X_alpha"], None, d["C"], None, d["noise"])
        lf, lr, _, _ = mw('d', d["X_tex"], d["X_alpha"], d["X_mesh"], d["C"], None, d["noise"])
        out.update({tag + "_g_loss": loss.detach().numpy(), tag + "_g_tex": tex.detach().numpy(), tag + "_g_mesh": mesh.detach().numpy(),
                    tag + "_d_fake": lf.detach().numpy(), tag + "_d_real": lr.detach().numpy()})
        if tag == "w21":
            mw.eval()
            itex, imesh, attn = mw('inference', None, None, None, d["C"], None, d["noise"])
            assert attn is None
            out["inf_tex"], out["inf_mesh"] = itex.numpy(), imesh.numpy()
            mw.train()
            # running average: perturb the live generator, then the reference's update at three epochs
            with torch.no_grad():
                for p in mw.generator.parameters():
                    p.add_(0.05 * torch.randn(p.shape, generator=torch.Generator().manual_seed(p.numel())))
                mw.generator.bn.num_batches_tracked.fill_(7)
            ns["generator"], ns["generator_running_avg"] = mw.generator, mw.generator_running_avg
            for epoch in (5, 50, 500):
                ns["update_generator_running_avg"](epoch)
                for k, v in mw.generator_running_avg.state_dict().items():
                    out[f"avg{epoch}_{k}"] = v.numpy().copy()
    p = os.path.join(HERE, "wrapper_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes;", {k: v.tolist() for k, v in out.items() if k.endswith("loss") or "_d_" in k})


if __name__ == "__main__":
    main()
