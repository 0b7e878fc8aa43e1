"""Generate tests/golden/recon_step_reference.npz by EXECUTING the reference's reconstruction training iteration
(authoring container only):
    python tests/golden/make_golden_recon_step.py
run_reconstruction.py cannot be imported (argparse, datasets, kaolin and .cuda() at module scope).  Taken from its syntax tree
and compiled unmodified: `mean_iou` (:225-231), `transform_vertices` (:237-252) and the `for i, (X, gt_scale, ...) in
enumerate(train_loader):` loop of the training section (:409-465) — the loop body IS the training iteration.  The namespace
provides the module globals it reads: args, the optimisers exactly as :338-345 builds them (Adam, default betas), criterion =
nn.MSELoss() (:347-352), flat_warmup = 10 (:356), the reference's own loss_flat / qrot / DatasetParams, a log function, and —
standing in for what needs kaolin / CUDA / the real networks — the tiny network, template and renderer of
recon_step_common.py.  `.cuda()` is the identity.  Four iterations; stored: every iteration's total loss
(g_curve), the logged pieces, the warm-up factor afterwards, and the updated network / DatasetParams parameters.  Nothing of
the reference is copied into the repository — only its outputs."""
import ast
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
from models.reconstruction import DatasetParams               # noqa: E402  (reference)
from rendering.utils import qrot                              # noqa: E402  (reference)
from utils.losses import loss_flat                            # noqa: E402  (reference)

torch.Tensor.cuda = lambda self, *a, **k: self                # D15


def find_loop(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.For) and isinstance(node.iter, ast.Call) and getattr(node.iter.func, "id", "") == "enumerate" \
                and getattr(node.iter.args[0], "id", "") == "train_loader" and isinstance(node.target, ast.Tuple) \
                and any(isinstance(s, ast.Assign) and getattr(s.targets[0], "id", "") == "recon_loss" for s in node.body):
            return node
    raise RuntimeError("training loop not found")


def run(deltas, z0):
    tree = ast.parse(open(os.path.join(REF, "run_reconstruction.py")).read())
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("mean_iou", "transform_vertices")]
    assert len(funcs) == 2
    args = RS.make_args(deltas, z0)
    generator = RS.build_net().train()
    dataset_params = DatasetParams(args, 10) if (deltas or z0) else None
    logs = []
    ns = {"torch": torch, "qrot": qrot, "args": args, "generator": generator, "dataset_params": dataset_params,
          "optimizer": optim.Adam(generator.parameters(), lr=args.lr),
          "optimizer_dataset": optim.Adam(dataset_params.parameters(), lr=args.lr_dataset) if dataset_params is not None else None,
          "criterion": nn.MSELoss(), "loss_flat": loss_flat, "mesh_template": RS.Template(), "renderer": None,
          "train_loader": RS.batches(), "flat_warmup": 10, "g_curve": [], "total_it": 0, "epoch": 0, "time": time,
          "log": logs.append}
    exec(compile(ast.Module(body=funcs + [find_loop(tree)], type_ignores=[]), "run_reconstruction.py", "exec"), ns)
    assert ns["total_it"] == 4 and len(logs) == 1
    out = {"g_curve": np.array(ns["g_curve"]), "flat_warmup": np.float64(ns["flat_warmup"]), "log0": np.array(logs[0])}
    for k, v in generator.state_dict().items():
        out["net." + k] = v.numpy().copy()
    if dataset_params is not None:
        for k, v in dataset_params.state_dict().items():
            out["dp." + k] = v.numpy().copy()
    return out


def main():
    out = {}
    # (neither option: transform_vertices :251 dereferences dataset_params = None — the reference cannot run that setting)
    for tag, deltas, z0 in (("deltas", True, False), ("full", True, True), ("z0", False, True)):
        for k, v in run(deltas, z0).items():
            out[f"{tag}.{k}"] = v
    p = os.path.join(HERE, "recon_step_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes;", {k: v.tolist() for k, v in out.items() if k.endswith("g_curve") or k.endswith("log0")})


if __name__ == "__main__":
    main()
