"""Shared by make_golden_wrapper.py (reference side) and tests/test_gan_hostlogic.py (drop-in side): tiny CPU stand-ins with
the call signatures of models/gan.py's Generator / MultiScaleDiscriminator, and seeded inputs."""
import types

import torch
import torch.nn as nn


class TinyG(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(8, 3 * 8 * 8)
        self.bn = nn.BatchNorm2d(3)               # a float buffer AND an integer one (num_batches_tracked) in the state dict
        self.emb = nn.Embedding(10, 8)
        self.mesh = nn.Conv2d(3, 3, 1)

    def forward(self, z, c=None, caption=None, return_attention=False):
        h = z + (self.emb(c.view(-1)) if c is not None else 0)
        tex = torch.tanh(self.bn(self.fc(h).view(-1, 3, 8, 8)))
        mesh = 0.1 * self.mesh(tex)
        return (tex, mesh, None) if return_attention else (tex, mesh)


class TinyD(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Conv2d(4, 1, 3, padding=1)
        self.b = nn.Conv2d(4, 1, 4, stride=2, padding=1)
        self.m = nn.Conv2d(3, 1, 1)

    def forward(self, x, mesh_map=None, c=None, caption=None):
        alpha = x[:, 3:4]
        out = [self.a(x) + (c.float().view(-1, 1, 1, 1) * 0.01 if c is not None else 0), self.b(x) + self.m(mesh_map).mean()]
        return out, [alpha, torch.nn.functional.avg_pool2d(alpha, 2)]


def make_args(nd=2, res=512):
    return types.SimpleNamespace(loss='hinge', num_discriminators=nd, texture_resolution=res, latent_dim=8, conditional_text=False,
                                 g_running_average_alpha=0.999, evaluate=False, text_train_encoder=False)


def build(seed=3):
    torch.manual_seed(seed)
    return (lambda: TinyG()), TinyD()


def inputs(seed=9, B=4):
    g = torch.Generator().manual_seed(seed)
    return dict(X_tex=torch.rand(B, 3, 8, 8, generator=g) * 2 - 1, X_alpha=(torch.rand(B, 1, 8, 8, generator=g) > 0.4).float(),
                X_mesh=torch.randn(B, 3, 8, 8, generator=g) * 0.05, C=torch.randint(0, 10, (B, 1), generator=g),
                noise=torch.randn(B, 8, generator=g))
