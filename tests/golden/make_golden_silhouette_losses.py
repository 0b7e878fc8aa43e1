"""Generate tests/golden/silhouette_losses.npz from the REFERENCE's UnsupervisedLoss / SupervisedLoss (authoring container
only):   python tests/golden/make_golden_silhouette_losses.py
Import shim of SURVEY §8c; the only patch is the execution patch for App. A D6 (`num_candidates` is never set)."""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))


def inputs(B=5, K=4, V=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    projection = torch.rand(B * K, V, V, generator=g)
    masks = (torch.rand(B, 2 * V, 2 * V, generator=g) > 0.5).float()
    ensemble = torch.nn.functional.normalize(torch.randn(B * K, 4, generator=g), dim=-1)
    student = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    return projection, masks, ensemble, student


def main():
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [REF]
    sys.modules["refpkg"] = pkg
    for sub in ("utils", "models"):
        sys.path.insert(0, os.path.join(REF, sub))
    up = importlib.import_module("refpkg.models.unsupervised_part")
    sp = importlib.import_module("refpkg.models.supervised_part")
    projection, masks, ensemble, student = inputs()
    projection.requires_grad_(True); student.requires_grad_(True)
    loss = up.UnsupervisedLoss(4, 20.0)
    loss.num_candidates = 4                                      # D6
    out = loss((projection, ensemble, student), masks, True)
    out["total_loss"].backward()
    res = {k: np.float64(v.detach()) for k, v in out.items()}
    res["minimum_indexes"] = loss.minimum_indexes.numpy()
    res["d_projection"] = projection.grad.numpy().copy()
    res["d_student"] = student.grad.numpy().copy()
    res["eval_projection_loss"] = np.float64(loss((projection.detach()[:5],), masks, False)["projection_loss"])
    res["supervised_full_loss"] = np.float64(sp.SupervisedLoss()(projection.detach()[:5], masks)["full_loss"])
    path = os.path.join(HERE, "silhouette_losses.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, {k: (float(v) if np.ndim(v) == 0 else v.shape) for k, v in res.items()})


if __name__ == "__main__":
    main()
