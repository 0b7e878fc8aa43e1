"""Host-side logic of the GAN drop-ins (CPU): state-dict layout, same-seed initial values and positional
encoding equal the reference's when /root/reference is present (authoring container); GANLoss vs golden."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, PKG

sys.path.insert(0, GOLDEN)
import gan_common as GC          # noqa: E402

REF = "/root/reference/code"


def test_ganloss_matches_reference_golden():
    from utils.losses import GANLoss
    g = np.load(os.path.join(GOLDEN, "mesh_reference_pieces.npz"))
    t = lambda k: torch.tensor(g[k])
    crit = GANLoss('hinge')
    preds, masks = [t("gl_p0"), t("gl_p1")], [t("gl_m0"), t("gl_m1")]
    assert abs(float(crit(preds, True, for_discriminator=False, mask=masks, weight=[2, 1])) - float(g["gl_g"])) < 1e-7
    assert abs(float(crit(preds, False, for_discriminator=True, mask=masks)) - float(g["gl_d_fake"])) < 1e-7
    assert abs(float(crit(preds, True, for_discriminator=True, mask=masks, weight=[2, 1])) - float(g["gl_d_real"])) < 1e-7
    assert abs(float(crit(preds, True, for_discriminator=False)) - float(g["gl_g_nomask"])) < 1e-7
    with pytest.raises(ValueError):
        GANLoss('nope')


def test_state_dict_layout():
    from models import gan
    args = GC.make_args(512, 3)
    G, D = GC.build(gan, args)
    sd = G.state_dict()
    # names and shapes SURVEY §8b lists for the shipped checkpoints
    assert sd["blk1.conv1.weight_orig"].shape == (512, 512, 3, 3)
    assert sd["blk1.conv1.weight_u"].shape == (512,) and sd["blk1.conv1.weight_v"].shape == (4608,)
    assert sd["blk1.norm1.fc_gamma.weight"].shape == (512, 128) and "blk1.norm1.norm.running_mean" in sd
    assert sd["emb_class.weight"].shape == (200, 64) and sd["fc.weight"].shape == (16384, 128)
    assert len(sd) == 179                                   # 179 entries in generator_running_avg
    assert sum(p.numel() for p in G.parameters()) > 13.0e6
    assert {"d1.conv1.weight_orig", "d2.conv4.bias", "d3.conv5.weight_orig", "d1.projector.weight"} <= set(D.state_dict())
    with pytest.raises(Exception, match="CUDA only|no CPU fallback"):
        G(torch.zeros(1, 64), torch.zeros(1, 1, dtype=torch.long))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_equal_to_reference_modules():
    # import the reference's models.gan next to ours under a private name
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split('.')[0] in ("models", "rendering", "utils", "sync_batchnorm")}
    sys.path.remove(PKG)
    sys.path.insert(0, REF)
    try:
        ref = importlib.import_module("models.gan")
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split('.')[0] in ("models", "rendering", "utils", "sync_batchnorm")]:
            sys.modules.pop(k)
        sys.path.insert(0, PKG)
        sys.modules.update(saved)
    from models import gan
    args = GC.make_args(256, 2)
    Gr, Dr = GC.build(ref, args)
    Gm, Dm = GC.build(gan, args)
    for a, b in ((Gr, Gm), (Dr, Dm)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
    for ny, nx in ((32, 32), (32, 16), (256, 128)):
        assert np.abs(ref.positional_encoding(ny, nx) - gan.positional_encoding(ny, nx)).max() < 1e-12
    ck = os.path.join(REF, "gan_weights/pretrained_weights_cub/checkpoint_latest.pth")
    G5 = gan.Generator(GC.make_args(512, 3), 64)
    G5.load_state_dict(torch.load(ck, map_location="cpu")["generator_running_avg"], strict=True)
