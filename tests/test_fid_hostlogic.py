"""FID evaluation path on the CPU: the statistics / distance maths against goldens produced by the
reference's own utils/fid.py, and the host logic of the Inception drop-in — batch-norm folding, tap lists, channel
padding, average-pool folding, branch plans and concat offsets — against the oracle's restatement of the network, with
the four device primitives replaced by torch emulations of their C-ABI semantics (the kernels themselves are checked on
the GPU by tests/test_fid_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from fid_common import randomize_inception
from oracle import fid as OF


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "fid_reference.npz"))


def test_statistics_and_distance_match_the_reference(gold):
    from utils.fid import calculate_frechet_distance, calculate_stats
    for mod_stats, mod_fd in ((calculate_stats, calculate_frechet_distance), (OF.calculate_stats, OF.calculate_frechet_distance)):
        m1, s1 = mod_stats(gold["act_a"])
        m2, s2 = mod_stats(gold["act_b"])
        np.testing.assert_allclose(m1, gold["mu_a"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(s1, gold["sigma_a"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(s2, gold["sigma_b"], rtol=1e-6, atol=1e-7)
        assert abs(mod_fd(m1, s1, m2, s2) - float(gold["fid_ab"])) < 1e-6 * float(gold["fid_ab"])
        assert abs(mod_fd(m1, s1, m1, s1)) < 1e-6
        m3, s3 = mod_stats(gold["act_c"])                       # 24 samples in 48 dimensions: singular covariance
        assert abs(mod_fd(m3, s3, m1, s1) - float(gold["fid_ca"])) < 1e-5 * float(gold["fid_ca"])


def test_distance_on_the_shipped_cub_statistics(gold):
    """Principal 128-d block of the reference's cached real-image statistics (train vs test+val split)."""
    from utils.fid import calculate_frechet_distance
    d = calculate_frechet_distance(gold["cub_mu_train"], gold["cub_sigma_train"], gold["cub_mu_val"], gold["cub_sigma_val"])
    assert abs(d - float(gold["cub_fid_sub"])) < 1e-6 + 1e-5 * float(gold["cub_fid_sub"])
    d2 = OF.calculate_frechet_distance(gold["cub_mu_train"], gold["cub_sigma_train"], gold["cub_mu_val"], gold["cub_sigma_val"])
    assert abs(d2 - float(gold["cub_fid_sub"])) < 1e-9


def test_truncated_noise():
    z = OF.truncated_noise(64, 64, 1.0, torch.Generator().manual_seed(3))
    assert z.shape == (64, 64) and float(z.abs().max()) <= 1.0 and float(z.std()) > 0.4
    from fid_evaluation import truncated_noise
    z2 = truncated_noise(64, 64, 1.0, torch.Generator().manual_seed(3))
    assert torch.equal(z, z2)


def test_inception_structure_and_state_dict_names():
    from utils.inception import InceptionV3
    m = InceptionV3([0, 1, 2, 3], weights=None)
    sd = m.state_dict()
    assert sum(v.numel() for k, v in sd.items() if "num_batches" not in k and "running" not in k) == 21_785_568
    assert sd["blocks.0.0.conv.weight"].shape == (32, 3, 3, 3) and sd["blocks.1.0.conv.weight"].shape == (80, 64, 1, 1)
    assert sd["blocks.2.3.branch3x3dbl_3.conv.weight"].shape == (96, 96, 3, 3)
    assert sd["blocks.2.4.branch7x7_2.conv.weight"].shape == (128, 128, 1, 7)
    assert sd["blocks.3.2.branch3x3dbl_3b.conv.weight"].shape == (384, 384, 3, 1)
    # torchvision's own key names (plus the heads the extractor drops) load through the name map
    tv = {}
    for k, v in randomize_inception(InceptionV3([3], weights=None), 5).state_dict().items():
        for tvn, mine in InceptionV3._TV_NAMES:
            if k.startswith(mine + "."):
                tv[tvn + k[len(mine):]] = v
    tv["fc.weight"], tv["AuxLogits.conv0.conv.weight"] = torch.zeros(1000, 2048), torch.zeros(128, 768, 1, 1)
    m2 = InceptionV3([3], weights=None)
    m2.load_torchvision_state_dict(tv)
    ref = randomize_inception(InceptionV3([3], weights=None), 5).state_dict()
    assert all(torch.equal(v, ref[k]) for k, v in m2.state_dict().items())
    with pytest.raises(Exception):
        m2.load_torchvision_state_dict({k: v for k, v in tv.items() if not k.startswith("Mixed_7c")})
    with pytest.raises(FileNotFoundError):
        os.environ.pop("B3D_INCEPTION_WEIGHTS", None)
        InceptionV3([3])                       # weights='pretrained' never downloads
    with pytest.raises(Exception):
        m2.train()


def test_oracle_block_shapes_at_299():
    from utils.inception import InceptionV3
    m = randomize_inception(InceptionV3([0, 1, 2, 3], weights=None), 1)
    outs = OF.inception_forward(m.state_dict(), torch.rand(1, 3, 64, 64), (0, 1, 2, 3))
    assert [tuple(o.shape[1:]) for o in outs] == [(64, 73, 73), (192, 35, 35), (768, 17, 17), (2048, 1, 1)]
    assert all(torch.isfinite(o).all() for o in outs) and 1e-3 < float(outs[3].abs().mean()) < 1e3


def _emulated(cls):
    """The drop-in with its four device primitives emulated in torch (same NHWC layouts, tap lists, output slices)."""
    class Emu(cls):
        def _input(self, inp):
            x = inp
            if self.resize_input:
                x = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
            if self.normalize_input:
                x = 2 * x - 1
            return F.pad(x.permute(0, 2, 3, 1), (0, 29)).contiguous()

        def _conv(self, x, m, avg_fold=False, out=None, coff=0):
            f = self._rec(m, avg_fold, x.device)
            N, H, W, C = x.shape
            assert C == f.cinp
            Ho, Wo = (H + 2 * f.ph - f.kh) // f.stride + 1, (W + 2 * f.pw - f.kw) // f.stride + 1
            cout = f.coutp if out is None else f.cout
            P = 8
            xp = F.pad(x, (0, 0, P, P, P, P))
            acc = f.bias[:cout].view(1, 1, 1, -1).expand(N, Ho, Wo, cout).clone()
            s = f.stride
            for t in range(f.ntaps):
                y0, x0 = P + f.dy[t], P + f.dx[t]
                patch = xp[:, y0:y0 + s * (Ho - 1) + 1:s, x0:x0 + s * (Wo - 1) + 1:s, :]
                acc += patch @ f.wt[t, :cout].t()
            acc = torch.relu(acc)
            if out is None:
                return acc
            assert tuple(out.shape[:3]) == (N, Ho, Wo)
            out[..., coff:coff + cout] = acc
            return out

        def _maxpool(self, x, out=None, coff=0):
            y = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
            if out is None:
                return y.contiguous()
            out[..., coff:coff + y.shape[3]] = y
            return out

        def _meanpool(self, h):
            return h.mean(dim=(1, 2), keepdim=True)
    return Emu


@pytest.mark.parametrize("size,resize", [(107, False), (40, True)])
def test_inception_host_logic_equals_the_oracle(size, resize):
    from utils import inception as I
    torch.manual_seed(11)
    m = randomize_inception(_emulated(I.InceptionV3)([0, 1, 2, 3], resize_input=resize, weights=None), 2)
    x = torch.rand(2 if not resize else 1, 3, size, size)
    got = m(x)
    ref = OF.inception_forward(m.state_dict(), x.double(), (0, 1, 2, 3), resize_input=resize)
    for g, r in zip(got, ref):
        assert g.shape == r.shape
        err = float((g.double() - r).abs().max()) / float(r.abs().max())
        assert err < 5e-3, err                              # tf32-rounded weights, fp32 accumulation vs the fp64 oracle
    assert torch.count_nonzero(got[3]) > 100


def test_tf32_rounding():
    from utils.inception import round_tf32
    w = torch.randn(4096)
    r = round_tf32(w)
    assert int((r.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float(((r - w).abs() / w.abs()).max()) <= 2.0 ** -11 + 1e-9


def test_real_statistics_cache_format(tmp_path):
    """main.py:170-172 reads stats_s as ONE triangle and mirrors it with `s + np.triu(s.T, 1)`: the file holds the LOWER one."""
    from fid_evaluation import load_real_statistics, save_real_statistics
    a = np.random.default_rng(0).standard_normal((40, 12))
    mu, sigma = a.mean(0), np.cov(a, rowvar=False)
    p = str(tmp_path / "precomputed_fid_299x299_train.npz")
    save_real_statistics(p, mu, sigma, 40)
    raw = np.load(p)["stats_s"]
    assert np.abs(np.triu(raw, 1)).max() == 0.0
    np.testing.assert_allclose(raw + np.triu(raw.T, 1), sigma, rtol=0, atol=0)         # the reference's own reconstruction
    m2, s2, n = load_real_statistics(p, 299, expect_images=40)
    np.testing.assert_allclose(s2, sigma, rtol=0, atol=0)
    assert n == 40 and np.array_equal(m2, mu)
    with pytest.raises(ValueError):
        load_real_statistics(p, 512)
    with pytest.raises(ValueError):
        load_real_statistics(p, 299, expect_images=41)


def test_evaluation_loop_host_logic(monkeypatch):
    """FIDEvaluator's control flow (main.py:188-376) with stand-ins for everything that needs the GPU: a generator and a
    'renderer' that are deterministic functions of their inputs, the emulated Inception, and the feature sums in torch."""
    import fid_evaluation as FE
    from utils import fid as UF
    from utils import inception as I

    def update(self, feat):
        f = feat.detach().double()
        self.sum += f.sum(0)
        self.outer += f.t() @ f
        self.n += f.shape[0]
    monkeypatch.setattr(UF.FIDStatistics, "update", update)

    class Gen(torch.nn.Module):
        def forward(self, z, c, caption, return_attention=False):
            assert return_attention and float(z.abs().max()) <= 1.5
            base = torch.linspace(-1, 1, 64 * 64 * 3).view(1, 3, 64, 64)
            return torch.tanh(base * z[:, :1, None, None] * 3), 0.05 * z[:, :3, None, None].expand(-1, -1, 32, 32), None

    class Template:
        def vertices_and_pose(self, m, s, t, r):
            assert s.shape == (m.shape[0], 1) and t.shape[1] == 3 and r.shape[1] == 4
            return None, m.mean(dim=(2, 3)) + t * s

        def forward_renderer(self, renderer, vtx, tex):
            img = F.interpolate(tex, size=(75, 75), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
            return (img * (1 + vtx.view(-1, 1, 1, 3))).clamp(-1, 1), None

    inc = randomize_inception(_emulated(I.InceptionV3)([3], resize_input=False, weights=None), 6)
    g = torch.Generator()

    def batches(n, B=3, pseudo=True, image=True):
        g.manual_seed(21)
        for i in range(n):
            d = {"idx": torch.arange(i * B, (i + 1) * B), "class": torch.randint(0, 200, (B, 1), generator=g),
                 "rotation": F.normalize(torch.randn(B, 4, generator=g), dim=-1), "scale": 0.5 + 0.3 * torch.rand(B, generator=g),
                 "translation": (torch.rand(B, 3, generator=g) - 0.5) * 0.2}
            img = torch.rand(B, 3, 75, 75, generator=g)            # always drawn: the same poses with and without real images
            if image:
                d["image"] = img
            if pseudo:
                d["texture"], d["mesh"] = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1, torch.randn(B, 3, 32, 32, generator=g) * 0.05
            yield d

    ev = FE.FIDEvaluator(Gen(), Template(), inception=inc, evaluation_res=75, truncation_sigma=1.5, device="cpu")
    with pytest.raises(ValueError):
        ev.evaluate(batches(1, image=False))                       # no real statistics and no real images
    out = ev.evaluate(batches(2), seed=1234, keep_features=True)
    assert set(out) == {"fid", "fid_texture_only", "fid_mesh_only", "features", "num_generated"} and out["num_generated"] == 6
    f = out["features"]["combined"].double().numpy()
    ref = OF.calculate_frechet_distance(*OF.calculate_stats(f), ev.m_real, ev.s_real)
    assert abs(ref - out["fid"]) < 1e-4 * abs(ref)
    # cached statistics: no real images needed, same seed -> same score; `fast` drops the pseudo-ground-truth renders
    ev2 = FE.FIDEvaluator(Gen(), Template(), inception=inc, evaluation_res=75, truncation_sigma=1.5, device="cpu")
    ev2.set_real_statistics(ev.m_real, ev.s_real)
    out2 = ev2.evaluate(batches(2, image=False), fast=True, seed=1234)
    assert set(out2) == {"fid", "num_generated"} and abs(out2["fid"] - out["fid"]) < 1e-9 * abs(out["fid"])
    # validation statistics: scores on a subset of the generated images of the validation split's size (main.py:359-376)
    ev2.set_real_statistics(ev.m_real, ev.s_real, validation=True, num_images=4)
    out3 = ev2.evaluate(batches(2, image=False), seed=1234)
    assert {"fid_val", "fid_texture_only_val", "fid_mesh_only_val"} <= set(out3) and abs(out3["fid"] - out["fid"]) < 1e-9 * abs(out["fid"])
    ev2.set_real_statistics(ev.m_real, ev.s_real, validation=True, num_images=7)
    with pytest.raises(ValueError):
        ev2.evaluate(batches(2, image=False), seed=1234)           # 'Not supported': more validation images than generated


def test_evaluator_matches_the_reference_evaluate_fid(monkeypatch):
    """tests/golden/fid_loop_reference.npz: main.py's `evaluate_fid` (:188-412) EXECUTED from the script's syntax tree on stand-in
    generator / template / renderer / extractor, with the reference's own utils/fid.py (make_golden_fid_loop.py).  FIDEvaluator on
    the same stand-ins (feature sums in torch instead of the CUDA kernel) must give the same scores: the seeded truncated-noise
    stream, the three renders per batch, real statistics from the images, the seeded validation subset, the fast mode.
    Tolerance 1e-5 relative (observed 2e-8): 18 samples in 64 dimensions make every covariance singular, where the reference's
    scipy.linalg.sqrtm itself warns about its accuracy."""
    import sys
    sys.path.insert(0, GOLDEN)
    import fid_loop_common as FL
    import recon_step_common as RS
    import wrapper_common as WC
    import fid_evaluation as FE
    from oracle import mesh as OM
    from utils import fid as UF

    def update(self, feat):
        f = feat.detach().double()
        self.sum += f.sum(0)
        self.outer += f.t() @ f
        self.n += f.shape[0]
    monkeypatch.setattr(UF.FIDStatistics, "update", update)
    d = np.load(os.path.join(GOLDEN, "fid_loop_reference.npz"))
    gi, _ = WC.build()
    G = gi()
    G2 = gi()
    G2.load_state_dict(G.state_dict())          # ModelWrapper's running-average copy: the SECOND instance, same weights
    tpl = RS.Template(map_size=8)
    tpl.vertices_and_pose = lambda m, s, t, r: (lambda raw: (raw, OM.transform_vertices(raw, s, t, r)))(tpl.get_vertex_positions(m))
    ev = FE.FIDEvaluator(G2, tpl, inception=FL.Extractor(), evaluation_res=FL.RES, latent_dim=8, truncation_sigma=1.0, device="cpu")

    def close(a, b):
        assert abs(a - b) <= 1e-5 * abs(b), (a, b)
    out = ev.evaluate(FL.eval_set(), seed=1234)
    for k, ref in zip(("fid", "fid_texture_only", "fid_mesh_only"), d["run1"]):
        close(out[k], ref)
    np.testing.assert_allclose(ev.m_real, d["m_real"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(ev.s_real, d["s_real"], rtol=0, atol=1e-6)
    ev.set_real_statistics(d["m_val"], d["s_val"], validation=True, num_images=11)
    out = ev.evaluate(FL.eval_set(), seed=1234)
    for k, ref in zip(("fid", "fid_texture_only", "fid_mesh_only", "fid_val", "fid_texture_only_val", "fid_mesh_only_val"), d["run2"]):
        close(out[k], ref)
    out = ev.evaluate(FL.eval_set(), seed=1234, fast=True)
    close(out["fid"], d["run3"][0])
    assert set(out) == {"fid", "num_generated"}
