"""Shared by make_golden_recon_step.py (reference side) and tests/test_recon_hostlogic.py (drop-in side): CPU stand-ins for the
network, the mesh template and the renderer with the call signatures the training iteration uses, and seeded batches.
They are smooth, differentiable functions of their inputs — enough for the ITERATION logic (what feeds which loss, the flat-loss
warm-up, the pose deltas / z0, the two optimisers) to show in every number."""
import types

import torch
import torch.nn as nn

V, F_, H = 12, 8, 10          # vertices, faces, image size of the stand-in scene


class TinyNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = nn.Conv2d(4, 6, 3, padding=1, bias=False)     # (a bias in front of a batch norm has a zero gradient: Adam would amplify rounding noise)
        self.bn = nn.BatchNorm2d(6)
        self.tex = nn.Conv2d(6, 3, 1)
        self.msh = nn.Linear(6, 3 * 4 * 4)

    def forward(self, x):
        h = torch.relu(self.bn(self.enc(x)))
        return torch.tanh(self.tex(h)), 0.1 * self.msh(h.mean(dim=(2, 3))).view(-1, 3, 4, 4)


class Template:
    """get_vertex_positions / compute_normals / forward_renderer of rendering/mesh_template.py, on a fixed toy topology."""

    def __init__(self, map_size=4):
        g = torch.Generator().manual_seed(2)
        self.base = torch.randn(V, 3, generator=g) * 0.3
        self.mix = torch.randn(3 * map_size * map_size, V * 3, generator=g) * 0.2
        faces = torch.stack([torch.randperm(V, generator=g)[:3] for _ in range(F_)])
        ff = torch.stack([torch.tensor([(i + 1) % F_, (i + 3) % F_, (i + 5) % F_]) for i in range(F_)])
        self.mesh = types.SimpleNamespace(faces=faces, ff=ff)
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing="ij")
        self.grid = torch.stack((xs, ys), dim=-1)                      # [H,H,2]

    def get_vertex_positions(self, mesh_map):
        return self.base.unsqueeze(0) + (mesh_map.flatten(1) @ self.mix).view(-1, V, 3)

    def compute_normals(self, vtx):
        f = self.mesh.faces
        a, b, c = vtx[:, f[:, 0]], vtx[:, f[:, 1]], vtx[:, f[:, 2]]
        return torch.nn.functional.normalize(torch.cross(b - a, c - a, dim=2), dim=2)

    def forward_renderer(self, renderer, vtx, tex, num_gpus=1, **kw):
        # soft blobs at the projected vertices, coloured by the mean texture: differentiable in vtx and tex
        d2 = ((self.grid.view(1, H, H, 1, 2) - vtx[:, None, None, :, :2]) ** 2).sum(-1)          # [B,H,H,V]
        w = torch.exp(-8.0 * d2) * torch.sigmoid(4.0 * vtx[:, None, None, :, 2])
        alpha = 1 - torch.prod(1 - 0.9 * w, dim=-1, keepdim=True)
        image = alpha * tex.mean(dim=(2, 3)).view(-1, 1, 1, 3)
        return image, alpha


def make_args(deltas=True, z0=False):
    return types.SimpleNamespace(optimize_deltas=deltas, optimize_z0=z0, mesh_regularization=0.00005, loss='mse', tensorboard=False,
                                 lr=0.01, lr_dataset=0.02)


def batches(n=4, B=3, N=10, seed=13):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        X = torch.rand(B, 4, H, H, generator=g) * 2 - 1
        X[:, 3] = (X[:, 3] > 0).float()
        out.append((X, 0.8 + 0.4 * torch.rand(B, 1, generator=g), (torch.rand(B, 3, generator=g) - 0.5) * 0.4,
                    torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1), torch.randint(0, 2 * N, (B, 1), generator=g)))
    return out


def build_net(seed=6):
    torch.manual_seed(seed)
    return TinyNet()
