"""Shared by make_golden_fid_loop.py (reference side) and tests/test_fid_hostlogic.py (drop-in side): a 64-feature extractor with
the reference InceptionV3's call convention (list of block outputs) and a seeded evaluation set."""
import torch
import torch.nn as nn
import torch.nn.functional as F

RES = 10            # evaluation resolution of the stand-in renderer (recon_step_common.H)


class Extractor(nn.Module):
    output_blocks = [0]

    def __init__(self):
        super().__init__()
        self.proj = nn.Parameter(torch.randn(64, 3 * 5 * 5, generator=torch.Generator().manual_seed(4)), requires_grad=False)

    def forward(self, images):
        p = F.adaptive_avg_pool2d(images, 5).flatten(1)
        return [torch.tanh(p @ self.proj.t()).view(-1, 64, 1, 1)]


def eval_set(n_batches=6, B=3, pseudo=True):
    g = torch.Generator().manual_seed(61)
    out = []
    for i in range(n_batches):
        d = {"idx": torch.arange(i * B, (i + 1) * B), "class": torch.randint(0, 10, (B, 1), generator=g),
             "rotation": F.normalize(torch.randn(B, 4, generator=g), dim=-1), "scale": 0.8 + 0.4 * torch.rand(B, 1, generator=g),
             "translation": (torch.rand(B, 3, generator=g) - 0.5) * 0.4, "image": torch.rand(B, 3, RES, RES, generator=g)}
        tex, mesh = torch.rand(B, 3, 8, 8, generator=g) * 2 - 1, torch.randn(B, 3, 8, 8, generator=g) * 0.05
        if pseudo:
            d["texture"], d["mesh"] = tex, mesh
        out.append(d)
    return out
