"""b3d.bank.WeightBank (csrc/sn_kernels.cu: spectral norm + kernel weight layouts of all layers in four launches) against
oracle/gan.py:sn_weight — the restatement of torch.nn.utils.spectral_norm that tests/test_gan_oracle.py pins to golden
vectors produced by the reference's own modules.  fp32 tolerances: 2e-6 relative on u / v / sigma-normalised weights
(different summation order of the mat-vecs), 1e-5 relative on gradients."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle import gan as OG

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_layers():
    from models.gan import TCConv2d
    torch.manual_seed(7)
    sn = nn.utils.spectral_norm
    layers = {
        "a": sn(TCConv2d(64, 128, 3, padding=(1, 0), bias=False)),      # ResBlockUp conv
        "stem": sn(TCConv2d(8, 64, 5, padding=(2, 0))),                 # discriminator stem, folded (kh into channels)
        "s2": sn(TCConv2d(64, 128, 4, padding=(1, 0), stride=2)),       # 4x4 / stride 2
        "head": sn(TCConv2d(256, 1, 5, padding=(2, 0))),                # 1-channel head (Cout' = 32 in the D layout)
        "short": sn(TCConv2d(128, 64, 1, bias=False)),                  # 1x1 shortcut
        "plain": TCConv2d(64, 3, 5, padding=(2, 0)),                    # conv_final: no spectral norm
        "stem11": sn(TCConv2d(11, 64, 5, padding=(2, 0))),              # mesh discriminator stem: 55 -> 64 channels
        "thin32": sn(TCConv2d(8, 64, 4, padding=(1, 0), stride=2)),     # 512^2 stem: 8 -> 32 zero-padded channels, no fold
    }
    return nn.ModuleDict(layers).to(DEV)


def oracle_state(mods):
    sd = {}
    for n, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{n}.{k}"] = v.detach().cpu().clone()
    return sd


def f_layout(w, fold):
    """[Cout,Cin,kh,kw] -> the bank's F layout (plain torch)."""
    co, ci, kh, kw = w.shape
    if fold:
        f = w.permute(3, 0, 2, 1).reshape(kw, co, kh * ci)                 # [s][co][r*Cin + c]
    else:
        f = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    pad = (-f.shape[2]) % 32
    return torch.nn.functional.pad(f, (0, pad))


@pytest.mark.parametrize("training", [True, False])
def test_bank_matches_spectral_norm_oracle(training):
    from b3d.bank import WeightBank
    mods = make_layers()
    mods.train(training)
    sd = oracle_state(mods)
    fold = ("stem", "stem11")
    bank = WeightBank(dict(mods.items()), fold=fold, round_tf32=False)       # exact W / sigma for the comparison
    W = bank.forward(training)
    gs = {n: torch.randn(W[n].wf.shape, generator=torch.Generator().manual_seed(i)).to(DEV) for i, n in enumerate(W)}
    loss = sum((W[n].wf * gs[n]).sum() for n in W)
    params = {n: (m.weight_orig if hasattr(m, "weight_orig") else m.weight) for n, m in mods.items()}
    grads = torch.autograd.grad(loss, list(params.values()))
    torch.cuda.synchronize()
    for (n, m), g in zip(mods.items(), grads):
        is_sn = hasattr(m, "weight_orig")
        w0 = sd[f"{n}.weight_orig" if is_sn else f"{n}.weight"].clone().requires_grad_(True)
        if is_sn:
            sdo = {f"{n}.weight_orig": w0, f"{n}.weight_u": sd[f"{n}.weight_u"].clone(), f"{n}.weight_v": sd[f"{n}.weight_v"].clone()}
            wn = OG.sn_weight(sdo, n, training)
            assert torch.allclose(m.weight_u.cpu(), sdo[f"{n}.weight_u"], rtol=2e-5, atol=2e-6), n
            assert torch.allclose(m.weight_v.cpu(), sdo[f"{n}.weight_v"], rtol=2e-5, atol=2e-6), n
            if not training:        # eval mode leaves the buffers untouched
                assert torch.equal(m.weight_u.cpu(), sd[f"{n}.weight_u"]) and torch.equal(m.weight_v.cpu(), sd[f"{n}.weight_v"])
        else:
            wn = w0
        ref_f = f_layout(wn, n in fold)
        got = W[n].wf.detach().cpu()
        assert got.shape == ref_f.shape, (n, got.shape, ref_f.shape)
        assert float((got - ref_f.detach()).abs().max()) <= 2e-6 * float(ref_f.abs().max()) + 1e-7, n
        if W[n].wd is not None:
            d = W[n].wd.cpu()
            co = got.shape[1]
            assert d.shape == (got.shape[0], got.shape[2], (co + 31) // 32 * 32)
            assert torch.equal(d[:, :, :co], got.transpose(1, 2)) and float(d[:, :, co:].abs().max() if d.shape[2] > co else 0) == 0
        # gradient through W / sigma (u, v constants): same upstream gradient in the F layout
        gref, = torch.autograd.grad((ref_f * gs[n].cpu()).sum(), w0)
        err = float((g.cpu() - gref).abs().max())
        assert err <= 2e-5 * float(gref.abs().max()) + 1e-6, (n, err, float(gref.abs().max()))


def test_bank_rounds_weights_to_tf32_by_default():
    """The tensor cores read fp32 words as tf32 by dropping 13 mantissa bits (truncation: a bias towards zero).  The bank
    therefore rounds the emitted weights to the nearest tf32 value (cvt.rna), which is what cuDNN's TF32 path does."""
    from b3d.bank import WeightBank
    mods = make_layers()
    mods.train(False)
    exact = WeightBank({"a": mods["a"]}, round_tf32=False).forward(False)["a"].wf
    rnd = WeightBank({"a": mods["a"]}).forward(False)["a"].wf
    assert int((rnd.view(torch.int32) & 0x1FFF).abs().max()) == 0               # low 13 mantissa bits cleared
    assert float(((rnd - exact).abs() / exact.abs().clamp_min(1e-30)).max()) <= 2.0 ** -11 * 1.001  # round to nearest
    assert float((rnd - exact).mean().abs()) < 1e-3 * float((rnd - exact).abs().mean()) + 1e-9      # unbiased


def test_two_forwards_before_backward_do_not_alias():
    """torch's spectral_norm clones u / v for the graph; the bank keeps per-call copies in its output buffer."""
    from b3d.bank import WeightBank
    mods = make_layers()
    mods.train(True)
    bank = WeightBank({"a": mods["a"]}, round_tf32=False)
    sd = oracle_state(nn.ModuleDict({"a": mods["a"]}))
    W1 = bank.forward(True)
    W2 = bank.forward(True)                       # second power iteration: u, v advance again
    g = torch.randn(W1["a"].wf.shape, device=DEV)
    g1, = torch.autograd.grad((W1["a"].wf * g).sum(), mods["a"].weight_orig)
    w0 = sd["a.weight_orig"].clone().requires_grad_(True)
    sdo = {"a.weight_orig": w0, "a.weight_u": sd["a.weight_u"].clone(), "a.weight_v": sd["a.weight_v"].clone()}
    w1 = OG.sn_weight(sdo, "a", True)             # first call's graph uses the first iteration's u, v
    gref, = torch.autograd.grad((f_layout(w1, False) * g.cpu()).sum(), w0)
    assert float((g1.cpu() - gref).abs().max()) <= 2e-5 * float(gref.abs().max())
    assert not torch.equal(W1["a"].wf, W2["a"].wf)


def test_banked_conv_matches_module_path():
    """models.gan discriminators / generator: bank path == module path (torch's spectral-norm hook + per-call layouts)
    on the same weights: outputs and parameter gradients (tf32 on both sides -> tight tolerance)."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import gan_common as GC
    from models import gan
    from utils.losses import GANLoss
    args = GC.make_args(256, 2)
    res = []
    for disable in (True, False):
        G, D = GC.build(gan, args)
        G.cuda().train(); D.cuda().train()
        G.disable_bank, D.disable_bank = disable, disable
        crit = GANLoss('hinge', tensor=torch.cuda.FloatTensor)
        z, c, alpha, tex, mesh = [t.cuda() for t in GC.inputs(args, B=2)]
        loss, pred_tex, pred_mesh, dout, mask = GC.g_step(G, D, crit, z, c, alpha)
        loss.mean().backward()
        res.append((pred_tex.detach(), pred_mesh.detach(), dout[0].detach(), {n: p.grad.clone() for n, p in G.named_parameters() if p.grad is not None},
                    {n: b.clone() for n, b in G.named_buffers()}))
    a, b = res
    for i in range(3):
        err, ref = float((a[i] - b[i]).abs().max()), float(a[i].abs().max())
        # ~25 stacked tf32 convs; the bank path feeds weights rounded to the nearest tf32, the module path lets the tensor
        # core truncate them: the two differ by up to 2^-11 per weight
        assert err <= 1e-2 * ref, ("output", i, err, ref)
    assert a[3].keys() == b[3].keys()
    # both paths run the same tf32 kernels on weights that agree to 1 ulp; what differs is the order of the split-K
    # atomics in the weight gradients.  Cancelling sums (scalar biases) get the golden test's absolute floor.
    # per-parameter gradient NORMS, the golden test's criterion (elementwise the B = 2 batch statistics amplify the 2^-11
    # weight differences too irregularly for a max-norm bound)
    floor = 2e-3 * max(float(g.norm()) for g in a[3].values())
    bad = []
    for n in a[3]:
        na, nb = float(a[3][n].norm()), float(b[3][n].norm())
        if abs(na - nb) > 8e-2 * na + floor:
            bad.append((n, na, nb))
    assert not bad, bad[:10]
    for n in a[4]:
        if a[4][n].dtype.is_floating_point:
            err = float((a[4][n] - b[4][n]).abs().max())
            assert err <= 1e-3 * float(a[4][n].abs().max()) + 1e-5, ("buffer", n, err)
        else:
            assert torch.equal(a[4][n], b[4][n]), n
