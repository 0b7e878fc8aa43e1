"""The pseudo-ground-truth record format (data/pseudo_gt.py) on the CPU: round trip, dtypes, and — when /root/reference is
present (authoring container) — that the reference's own dataset class reads our files and mirrors textures identically."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference/code"


def _record(seed=0, R=32):
    from data.pseudo_gt import make_record
    g = torch.Generator().manual_seed(seed)
    return make_record(torch.randn(3, 32, 32, generator=g) * 0.05, torch.rand(3, R, R, generator=g) * 2 - 1,
                       (torch.rand(1, R, R, generator=g) > 0.3).float(), torch.rand(4, 16, 16, generator=g) * 2 - 1)


def test_round_trip_and_dtypes(tmp_path):
    from data.pseudo_gt import load_pseudo_ground_truth, pseudo_gt_dir, save_pseudo_gt
    rec = _record()
    assert rec['texture'].dtype == torch.float16 and rec['texture_alpha'].dtype == torch.float16
    assert rec['image'].dtype == torch.float16 and rec['mesh'].dtype == torch.float32
    save_pseudo_gt(pseudo_gt_dir(str(tmp_path), 32), 7, rec)
    assert os.path.exists(os.path.join(str(tmp_path), "pseudogt_32x32", "7.npz"))
    out = load_pseudo_ground_truth(str(tmp_path), 32, 7)
    assert set(out) == {'image', 'texture', 'texture_alpha', 'mesh'}
    assert out['image'].shape == (3, 16, 16) and out['image'].dtype == torch.float32
    assert torch.equal(out['image'], rec['image'][:3].float() / 2 + 0.5)
    assert torch.equal(out['texture'], rec['texture'].float()) and torch.equal(out['mesh'], rec['mesh'])
    assert float(out['image'].min()) >= -1e-3 and float(out['image'].max()) <= 1 + 1e-3


def test_mirror_tex_is_an_involution_with_half_turn_shift():
    from data.pseudo_gt import mirror_tex
    t = torch.arange(2 * 3 * 8, dtype=torch.float32).reshape(2, 3, 8)
    m = mirror_tex(t)
    assert m.shape == t.shape and torch.equal(mirror_tex(m), t)
    # flip, then rotate u by half a turn: each half of the map is reversed in place
    assert torch.equal(m[..., :4], t[..., :4].flip(2)) and torch.equal(m[..., 4:], t[..., 4:].flip(2))


def test_visibility_to_mask():
    from data.pseudo_gt import visibility_to_mask
    v = torch.zeros(1, 3, 8, 8)
    v[0, 1, 2:4, 2:4] = 0.5
    m = visibility_to_mask(v, 16)
    assert m.shape == (1, 16, 16, 1) and set(m.unique().tolist()) <= {0.0, 1.0}
    assert m[0, 5, 5, 0] == 1 and m[0, 0, 0, 0] == 0 and m[0, 15, 15, 0] == 0


def test_poses_metadata_round_trip(tmp_path):
    from data.pseudo_gt import load_poses_metadata, save_poses_metadata
    s, t, r = torch.rand(5, 1), torch.rand(5, 2), torch.rand(5, 4)
    save_poses_metadata(str(tmp_path), s, t, r, [f"img{i}.jpg" for i in range(5)])
    d = load_poses_metadata(str(tmp_path))
    assert torch.equal(d['scale'], s) and torch.equal(d['rotation'], r) and d['path'][3] == "img3.jpg"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_dataset_reads_our_records(tmp_path, monkeypatch):
    import importlib.util
    from data.pseudo_gt import load_pseudo_ground_truth, mirror_tex, pseudo_gt_dir, save_pseudo_gt
    spec = importlib.util.spec_from_file_location("ref_abstract_dataset", os.path.join(REF, "data", "abstract_dataset.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rec = _record(seed=3)
    cache = os.path.join(str(tmp_path), "cache", "cub")
    save_pseudo_gt(pseudo_gt_dir(cache, 32), 0, rec)
    ds = object.__new__(ref.AbstractDataset)                     # bypass __init__ (it globs the real dataset)
    ds.args = types.SimpleNamespace(texture_resolution=32)
    ds.cache_dir = cache
    theirs = ds.load_pseudo_ground_truth(0)
    ours = load_pseudo_ground_truth(cache, 32, 0)
    assert set(theirs) == set(ours)
    for k in ours:
        assert torch.equal(theirs[k], ours[k]), k
    assert torch.equal(ref.AbstractDataset.mirror_tex(ours['texture']), mirror_tex(ours['texture']))
