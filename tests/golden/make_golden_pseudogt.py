"""Generate tests/golden/pseudogt_reference.npz by EXECUTING the reference's pseudo-ground-truth export
(authoring container only):
    python tests/golden/make_golden_pseudogt.py
Taken from run_reconstruction.py's syntax tree and compiled unmodified: `transform_vertices` (:237-252), the nested
`class InverseRenderer` (:506-527) and the export loop `for net_image, inception_image, hd_image, ... in tqdm(train_loader):`
(:542-604: network -> vertices -> pose -> render with a differentiable texture -> d(render)/d(texture) as the visibility mask
-> inverse render of the photograph into UV space -> masking -> fp16 records written with np.savez_compressed).
Executed with the reference's OWN MeshTemplate and Renderer classes (stand-ins for their kaolin calls as in
make_golden_template.py / make_golden_renderer.py: OBJ loading, `.cuda()` = identity, and the ORACLE's rasteriser in place of
kaolin's — the one unpinned piece), the reference's qrot / DatasetParams / utils.fid.forward_inception_batch, a tiny stand-in
network, and small sizes (render 64^2, pseudo-GT 16^2, photographs 24^2).  The records the loop wrote are the golden.
Nothing of the reference is copied into the repository — only its outputs."""
import ast
import os
import pathlib
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
from oracle import mesh as M                                   # noqa: E402

# ---- stand-ins for kaolin / torchvision (see make_golden_template.py, make_golden_renderer.py, make_golden_fid.py) ------
kal = types.ModuleType("kaolin")
kal.rep = types.ModuleType("kaolin.rep")


class Mesh:
    compute_adjacency_info = None


class TriangleMesh(Mesh):
    @classmethod
    def from_obj(cls, path, enable_adjacency=False):
        d = M.load_obj(path)
        m = cls()
        m.vertices, m.faces, m.uvs, m.face_textures = d["vertices"], d["faces"], d["uvs"], d["face_textures"]
        if enable_adjacency:
            m.ff = Mesh.compute_adjacency_info(m.vertices, m.faces)[8]
        return m

    def cuda(self):
        return self


def linear_rasterizer(width, height, p3d, p2d, normalz, attr):
    imfeat, improb, _, _ = M.rasterize(p3d, p2d, normalz, attr, height, width)
    return imfeat, improb


kal.rep.Mesh, kal.rep.TriangleMesh = Mesh, TriangleMesh
mods = {"kaolin": kal, "kaolin.rep": kal.rep}
for n in ("kaolin.graphics", "kaolin.graphics.dib_renderer", "kaolin.graphics.dib_renderer.rasterizer", "kaolin.graphics.dib_renderer.utils",
          "torchvision", "torchvision.models"):
    mods[n] = types.ModuleType(n)
mods["kaolin.graphics.dib_renderer.rasterizer"].linear_rasterizer = linear_rasterizer
mods["kaolin.graphics.dib_renderer.utils"].datanormalize = M.datanormalize
mods["torchvision"].models = mods["torchvision.models"]
sys.modules.update(mods)
torch.Tensor.cuda = lambda self, *a, **k: self                 # D15

from models.reconstruction import DatasetParams               # noqa: E402  (reference)
from rendering.mesh_template import MeshTemplate               # noqa: E402  (reference)
from rendering.renderer import Renderer                        # noqa: E402  (reference)
from rendering.utils import qrot                               # noqa: E402  (reference)
from utils.fid import forward_inception_batch                  # noqa: E402  (reference)
import pseudogt_common as PC                                   # noqa: E402


def main():
    src = open(os.path.join(REF, "run_reconstruction.py")).read()
    tree = ast.parse(src)
    tv = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "transform_vertices")
    inv = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "InverseRenderer")
    loop = next(n for n in ast.walk(tree) if isinstance(n, ast.For) and isinstance(n.target, ast.Tuple) and
                [getattr(e, "id", "") for e in n.target.elts][:3] == ["net_image", "inception_image", "hd_image"])
    args = types.SimpleNamespace(optimize_deltas=True, optimize_z0=False, pseudogt_resolution=PC.PSEUDO, dataset="toy")
    tmp = tempfile.mkdtemp()
    path = M.write_uvsphere_obj(os.path.join(tmp, "uvsphere_16rings.obj"), rings=16)
    mesh_template = MeshTemplate(path, is_symmetric=True)
    generator = PC.build_net().eval()
    dataset_params = DatasetParams(args, 10)
    with torch.no_grad():
        dataset_params.ds_translation.copy_(0.02 * torch.randn(10, 2, generator=torch.Generator().manual_seed(1)))
        dataset_params.ds_scale.copy_(0.02 * torch.randn(10, 1, generator=torch.Generator().manual_seed(2)))
    pseudogt_dir = os.path.join(tmp, "out")
    pathlib.Path(pseudogt_dir).mkdir()
    batches = PC.batches()
    loader = types.SimpleNamespace(dataset=types.SimpleNamespace(paths=[f"img_{i % 10}.jpg" for i in range(20)]))
    loader_iter = list(batches)
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "os": os, "qrot": qrot, "args": args, "generator": generator,
          "dataset_params": dataset_params, "mesh_template": mesh_template, "Renderer": Renderer, "renderer": Renderer(PC.RENDER, PC.RENDER),
          "renderer_res": PC.RENDER, "tqdm": lambda x: loader_iter, "train_loader": loader, "pseudogt_dir": pseudogt_dir,
          "forward_inception_batch": forward_inception_batch, "inception_model": PC.Extractor(),
          "all_path": [], "all_gt_scale": [], "all_gt_translation": [], "all_gt_rotation": [], "all_inception_activation": []}
    exec(compile(ast.Module(body=[tv, inv], type_ignores=[]), "run_reconstruction.py", "exec"), ns)
    ns["inverse_renderer"] = ns["InverseRenderer"](mesh_template.mesh, args.pseudogt_resolution, args.pseudogt_resolution)
    exec(compile(ast.Module(body=[loop], type_ignores=[]), "run_reconstruction.py", "exec"), ns)
    out = {}
    for b in batches:
        for idx in b[6].view(-1).tolist():
            rec = np.load(os.path.join(pseudogt_dir, f"{idx}.npz"), allow_pickle=True)["data"].item()
            for k, v in rec.items():
                out[f"{idx}.{k}"] = v.numpy()
    out["paths"] = np.array(ns["all_path"])
    out["ds_translation"], out["ds_scale"] = dataset_params.ds_translation.detach().numpy(), dataset_params.ds_scale.detach().numpy()
    cov = np.mean([float((out[k] != 0).mean()) for k in out if k.endswith("texture_alpha")])
    p = os.path.join(HERE, "pseudogt_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes; records", sorted(k for k in out if k.endswith(".mesh")), "mean alpha coverage", cov)


if __name__ == "__main__":
    main()
