"""Generate tests/golden/template_reference.npz by running the REFERENCE's own rendering/mesh_template.py:MeshTemplate,
unmodified, on the CPU (authoring container only):
    python tests/golden/make_golden_template.py
The class needs two things this image cannot give it, so the script provides them and nothing else:
  * `import kaolin` (mesh_template.py:1) — used for exactly one call, kal.rep.TriangleMesh.from_obj(path,
    enable_adjacency=True) (:18), plus the hook kal.rep.Mesh.compute_adjacency_info that the class itself overwrites with the
    reference's rendering/monkey_patches.py (:235-237).  The stand-in parses the OBJ (v / vt / f v/vt lines) and calls that
    hook for `ff`; kaolin's own OBJ parser is therefore NOT what is pinned here — everything after it is the reference's code.
  * `.cuda()` on tensors / the mesh (:19,44-47,74,92 — SURVEY App. A D15): made the identity.
Inputs: the procedural UV sphere of tools/uvsphere.py (travels with the repo; same construction as the shipped
uvsphere_16rings.obj) and, for structure facts + probes only, the shipped mesh_templates/uvsphere_{16,31}rings.obj.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/code")

from oracle import mesh as M                      # noqa: E402  (only its OBJ reader / sphere writer, for the stand-in)

# ---- stand-in for `kaolin` ------------------------------------------------------------------------------------------
kal = types.ModuleType("kaolin")
kal.rep = types.ModuleType("kaolin.rep")


class Mesh:
    compute_adjacency_info = None                 # overwritten by MeshTemplate._monkey_patch_dependencies


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


kal.rep.Mesh, kal.rep.TriangleMesh = Mesh, TriangleMesh
sys.modules["kaolin"], sys.modules["kaolin.rep"] = kal, kal.rep
torch.Tensor.cuda = lambda self, *a, **k: self    # D15: the class hard-codes .cuda()

from rendering.mesh_template import MeshTemplate  # noqa: E402  (reference, unmodified)


def record(out, tag, path, symmetric, full):
    t = MeshTemplate(path, is_symmetric=symmetric)
    g = torch.Generator().manual_seed(17 + int(symmetric))
    dmap = torch.randn(2, 3, 32, 32, generator=g) * 0.05
    tex = torch.rand(2, 3, 16, 16, generator=g) * 2 - 1
    deltas = torch.randn(2, t.nonneg_topo_map.shape[0] if symmetric else t.topo_map.shape[0], 3, generator=g)
    vtx = t.get_vertex_positions(dmap)
    uvs, padded = t.adjust_uv_and_texture(tex)
    out[tag + "_counts"] = np.array([len(t.pos_indices), len(t.neg_indices), len(t.nonneg_indices) - len(t.pos_indices),
                                     t.mesh.vertices.shape[0], t.mesh.faces.shape[0], t.mesh.uvs.shape[0]])
    out[tag + "_vtx_sum"] = vtx.double().sum(dim=(1,)).numpy()
    out[tag + "_vtx_probe"] = vtx[:, ::37].numpy()
    out[tag + "_normals_probe"] = t.compute_normals(vtx)[:, ::53].numpy()
    if full:
        out[tag + "_dmap"], out[tag + "_tex"], out[tag + "_deltas"] = dmap.numpy(), tex.numpy(), deltas.numpy()
        out[tag + "_vtx"], out[tag + "_normals"] = vtx.numpy(), t.compute_normals(vtx).numpy()
        out[tag + "_uvs"], out[tag + "_padded"] = uvs[0].numpy(), padded.numpy()
        out[tag + "_deform"] = t.deform(deltas).numpy()
        out[tag + "_topo"], out[tag + "_tangent"] = t.topo_map.numpy(), t.tangent_map.numpy()
        out[tag + "_pos"], out[tag + "_neg"] = t.pos_indices.numpy(), t.neg_indices.numpy()
        out[tag + "_nonneg"], out[tag + "_symmask"] = t.nonneg_indices.numpy(), t.symmetry_mask.numpy()
        out[tag + "_ff"] = t.mesh.ff.numpy().astype(np.int16)
    else:
        out[tag + "_dmap_seed"] = np.array([17 + int(symmetric)])


def main():
    out = {}
    tmp = tempfile.mkdtemp()
    for rings in (16, 31):
        p = M.write_uvsphere_obj(os.path.join(tmp, f"uvsphere_{rings}rings.obj"), rings=rings)
        for sym in (True, False):
            record(out, f"proc{rings}_{'sym' if sym else 'asym'}", p, sym, full=True)
    for rings in (16, 31):                        # the shipped templates: probes only (the OBJ files do not travel)
        p = f"/root/reference/code/mesh_templates/uvsphere_{rings}rings.obj"
        record(out, f"ship{rings}_sym", p, True, full=False)
    path = os.path.join(HERE, "template_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.tolist() for k, v in out.items() if k.endswith("_counts")})


if __name__ == "__main__":
    main()
