"""b3d.ew.stem_input (csrc/ew_kernels.cu stem_input_*_kernel): the discriminator stem's input assembly
pad_x(cat(x, positions), amount, mode) in one pass, against the torch composition it replaces (models/gan.py
`_with_positions` + the wrap-around padding, reference models/gan.py:102-111,95-96).  Pure data movement (and two-term sums
in the adjoint): bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,C1,C2,H,W,amount", [(3, 4, 4, 16, 24, 2), (2, 4, 4, 32, 32, 1), (2, 3, 1, 8, 10, 2), (1, 4, 0, 8, 8, 3)])
def test_stem_input_equals_cat_and_pad(mode, N, C1, C2, H, W, amount):
    from b3d.ew import pad_x, stem_input
    torch.manual_seed(0)
    x = torch.randn(N, C1, H, W, device=DEV, requires_grad=True)
    pos = torch.randn(C2, H, W, device=DEV)
    got = stem_input(x, pos, amount, mode)
    ref = pad_x(torch.cat((x, pos.unsqueeze(0).expand(N, -1, -1, -1)), dim=1), amount, mode)
    assert got.shape == ref.shape and torch.equal(got, ref)
    g = torch.randn_like(ref)
    ga, = torch.autograd.grad(got, x, g)
    gb, = torch.autograd.grad(ref, x, g)
    assert torch.equal(ga, gb)


def test_texture_discriminator_uses_it_and_matches_the_torch_path():
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import gan_common as GC
    from models import gan
    args = GC.make_args(256, 2)
    _, D = GC.build(gan, args)
    D.cuda().train()
    z, c, alpha, tex, mesh = [t.cuda() for t in GC.inputs(args, B=2)]
    x0 = torch.cat((tex, alpha), dim=1)
    saved = {n: b.clone() for n, b in D.named_buffers()}
    res = []
    for off in (True, False):
        D.d1.disable_stem_input = off
        with torch.no_grad():
            for n, b in D.named_buffers():
                b.copy_(saved[n])
        x = x0.clone().requires_grad_(True)
        out, _ = D(x, mesh, c)
        sum(o.sum() for o in out).backward()
        res.append((out[0].detach().clone(), x.grad.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
