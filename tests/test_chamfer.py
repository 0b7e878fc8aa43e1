"""Chamfer kernel vs the brute-force oracle (parity unpinned: the reference has no chamfer)."""
import pytest
import torch

from oracle import chamfer as C


def test_oracle_definition():
    torch.manual_seed(0)
    a, b = torch.rand(2, 7, 3), torch.rand(2, 5, 3)
    loss, dab, iab, dba, iba = C.chamfer(a, b)
    d = torch.cdist(a.double(), b.double()) ** 2
    assert torch.equal(iab, d.argmin(2)) and torch.equal(iba, d.argmin(1))
    assert torch.allclose(loss.double(), d.min(2).values.mean(1) + d.min(1).values.mean(1), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,M", [(2, 1000, 777), (1, 8000, 8000), (3, 5, 1), (2, 0, 9)])
def test_chamfer_matches_oracle(B, N, M):
    from b3d.chamfer import chamfer_distance, nearest
    g = torch.Generator().manual_seed(N + M)
    a, b = torch.rand(B, N, 3, generator=g), torch.rand(B, M, 3, generator=g)
    if N > 10:       # exact duplicates: ties must go to the lowest index
        b[:, 10] = b[:, 3]
    ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    if N == 0:
        d, i = nearest(ac, bc)
        assert d.shape == (B, 0)
        return
    loss, iab, iba = chamfer_distance(ac, bc, return_indices=True)
    lo, dab, oab, dba, oba = C.chamfer(a, b)
    assert torch.equal(iab.cpu().long(), oab) and torch.equal(iba.cpu().long(), oba)      # bit exact indices
    d, _ = nearest(ac, bc)
    assert torch.equal(d.cpu(), dab)                                                     # same fp32 op order
    assert torch.allclose(loss.cpu(), lo, rtol=1e-5, atol=1e-7)
    ao, bo = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    C.chamfer(ao, bo)[0].sum().backward()
    loss.sum().backward()
    assert torch.allclose(ac.grad.cpu(), ao.grad, atol=1e-6) and torch.allclose(bc.grad.cpu(), bo.grad, atol=1e-5)
