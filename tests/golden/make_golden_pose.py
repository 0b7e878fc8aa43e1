"""Generate tests/golden/pose_reference.npz by EXECUTING the reference's script-level helpers (authoring container only):
    python tests/golden/make_golden_pose.py
run_reconstruction.py cannot be imported (argparse, datasets and .cuda() at module scope), so the two function definitions
`mean_iou` (:225-231) and `transform_vertices` (:237-252) are taken from its syntax tree and compiled unmodified into a
namespace that provides what they read from module globals: torch, the reference's rendering.utils.qrot, `args` and a
reference DatasetParams.  Nothing of the reference is copied into the repository — only its outputs."""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))


def inputs(seed=4, B=6, P=11, N=10):
    g = torch.Generator().manual_seed(seed)
    vtx = torch.randn(B, P, 3, generator=g) * 0.4
    scale = torch.rand(B, 1, generator=g) + 0.5
    trans = torch.randn(B, 3, generator=g) * 0.2
    rot = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    idx = torch.tensor([0, 3, 12, 19, 7, 5])                      # >= N: mirrored copies
    state = torch.cat((torch.randn(N, 2, generator=g) * 0.1, torch.randn(N, 1, generator=g) * 0.1, torch.randn(N, 1, generator=g)), 1)
    alpha_p, alpha_r = torch.rand(B, 16, 16, generator=g), torch.rand(B, 16, 16, generator=g)
    return vtx, scale, trans, rot, idx, state, alpha_p, alpha_r


def load_params(module, state, deltas=True, z0=True):
    dp = module.DatasetParams(types.SimpleNamespace(optimize_deltas=deltas, optimize_z0=z0), state.shape[0])
    with torch.no_grad():
        if deltas:
            dp.ds_translation.copy_(state[:, :2]); dp.ds_scale.copy_(state[:, 2:3])
        if z0:
            dp.ds_z0.copy_(state[:, 3:4])
    return dp


def main():
    sys.path.insert(0, REF)
    from models import reconstruction as ref_recon
    from rendering.utils import qrot
    tree = ast.parse(open(os.path.join(REF, "run_reconstruction.py")).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("mean_iou", "transform_vertices")]
    assert len(wanted) == 2
    vtx, scale, trans, rot, idx, state, alpha_p, alpha_r = inputs()
    out = {}
    for tag, deltas, z0 in (("plain", False, False), ("deltas", True, False), ("full", True, True)):
        ns = {"torch": torch, "qrot": qrot, "args": types.SimpleNamespace(optimize_deltas=deltas, optimize_z0=z0),
              "dataset_params": load_params(ref_recon, state, deltas, z0)}
        exec(compile(ast.Module(body=wanted, type_ignores=[]), "run_reconstruction.py", "exec"), ns)
        out["vtx_" + tag] = ns["transform_vertices"](vtx, scale, trans, rot, idx).detach().numpy()
        out["iou"] = np.float64(ns["mean_iou"](alpha_p, alpha_r))
    path = os.path.join(HERE, "pose_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
