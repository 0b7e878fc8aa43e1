"""Multi-process FID evaluation on CPU (gloo, world size 2): every rank scores its shard of the evaluation set and the
feature sums are all-reduced (utils.fid.FIDStatistics.all_reduce) — the scores equal the single-process evaluation of the
whole set.  (The reference splits every batch over gpu_ids inside one process instead, main.py:163, :254-279.)
Stand-ins replace what needs a GPU: generator, renderer, a 64-feature extractor, and the feature sums in torch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import PKG, ROOT


class Gen(torch.nn.Module):
    """Deterministic in the class label only, so a shard sees the same samples as the full run."""

    def forward(self, z, c, caption, return_attention=False):
        t = (c.float().view(-1, 1, 1, 1) + 1) / 200.0
        base = torch.linspace(-1, 1, 3 * 16 * 16).view(1, 3, 16, 16)
        return torch.tanh(3 * base * t), 0.05 * t.expand(-1, 3, 32, 32), None


class Template:
    def vertices_and_pose(self, m, s, t, r):
        return None, m.mean(dim=(2, 3)) + t * s

    def forward_renderer(self, renderer, vtx, tex):
        img = F.interpolate(tex, size=(24, 24), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        return (img * (1 + vtx.view(-1, 1, 1, 3))).clamp(-1, 1), None


class Extractor(torch.nn.Module):
    output_blocks = [0]                          # FIDEvaluator reads the feature width from the block index: 64

    def __init__(self):
        super().__init__()
        self.proj = torch.nn.Parameter(torch.randn(64, 3 * 6 * 6, generator=torch.Generator().manual_seed(4)), requires_grad=False)

    def forward(self, images):
        p = F.adaptive_avg_pool2d(images, 6).flatten(1)
        return [torch.tanh(p @ self.proj.t()).view(-1, 64, 1, 1)]


def batches(lo, hi, B=3):
    g = torch.Generator().manual_seed(33)
    all_ = []
    for i in range(4):
        all_.append({"idx": torch.arange(i * B, (i + 1) * B), "class": torch.randint(0, 200, (B, 1), generator=g),
                     "rotation": F.normalize(torch.randn(B, 4, generator=g), dim=-1), "scale": 0.5 + 0.3 * torch.rand(B, generator=g),
                     "translation": (torch.rand(B, 3, generator=g) - 0.5) * 0.2, "image": torch.rand(B, 3, 24, 24, generator=g),
                     "texture": torch.rand(B, 3, 16, 16, generator=g) * 2 - 1, "mesh": torch.randn(B, 3, 32, 32, generator=g) * 0.05})
    return all_[lo:hi]


def run(lo, hi, distributed):
    from utils import fid as UF

    def update(self, feat):                      # b3d_fid_accumulate's arithmetic (the kernel needs a GPU)
        f = feat.detach().double()
        self.sum += f.sum(0)
        self.outer += f.t() @ f
        self.n += f.shape[0]
    UF.FIDStatistics.update = update
    import fid_evaluation as FE
    ev = FE.FIDEvaluator(Gen(), Template(), inception=Extractor(), evaluation_res=24, truncation_sigma=2.0, device="cpu")
    out = ev.evaluate(batches(lo, hi), seed=5, distributed=distributed)
    return {k: float(v) for k, v in out.items()}, ev.m_real, ev.s_real


def _worker(rank, world, port, q):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, m, s = run(2 * rank, 2 * rank + 2, True)
    q.put((rank, out, m.copy(), s.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, m_ref, s_ref = run(0, 4, False)
    assert ref["num_generated"] == 12
    for rank, out, m, s in res:
        assert out["num_generated"] == 12                                  # the whole set, on every rank
        np.testing.assert_allclose(m, m_ref, rtol=0, atol=1e-12)            # real-image statistics from both shards
        np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-12)
        for k in ("fid", "fid_texture_only", "fid_mesh_only"):
            assert abs(out[k] - ref[k]) <= 1e-7 * max(1.0, abs(ref[k])), (k, out[k], ref[k])
    assert res[0][1] == res[1][1] or all(abs(res[0][1][k] - res[1][1][k]) < 1e-9 for k in res[0][1])
