"""The run decomposition behind the TMA staging of the point-cloud bin records (csrc/pc_kernels.cu: chunk_runs, shared by
the device code that issues the cp.async.bulk copies and by b3d_pc_stream_plan): for random bin tables and patch ranges,
the runs of every chunk tile the chunk exactly and, concatenated over the chunks, reproduce the record sequence "bins
[bx_lo, bx_hi] of bin row by_lo, then of by_lo + 1, ..." the plain-load path walks."""
import ctypes

import numpy as np
import pytest


def plan(lib, bs, nbx, by_lo, by_hi, bx_lo, bx_hi, chunk, cap=64):
    d, s, c = ((ctypes.c_int * cap)() for _ in range(3))
    n = lib.b3d_pc_stream_plan(bs.ctypes.data_as(ctypes.c_void_p), nbx, by_lo, by_hi, bx_lo, bx_hi, chunk, d, s, c, cap)
    assert 0 <= n <= cap
    return [(d[i], s[i], c[i]) for i in range(n)]


@pytest.mark.parametrize("seed", range(6))
def test_runs_tile_every_chunk(seed):
    import b3d
    lib = b3d.lib
    S = lib.b3d_pc_stage_records()
    assert S >= 32 and S % 32 == 0
    rng = np.random.default_rng(seed)
    nbx, nby = int(rng.integers(1, 9)), int(rng.integers(1, 17))
    # bin populations from empty to several stages' worth
    counts = rng.choice([0, 0, 1, 7, 60, S - 1, S, S + 3, 3 * S], size=nbx * nby)
    bs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    for _ in range(20):
        by_lo = int(rng.integers(0, nby)); by_hi = int(rng.integers(by_lo, nby))
        bx_lo = int(rng.integers(0, nbx)); bx_hi = int(rng.integers(bx_lo, nbx))
        want = np.concatenate([np.arange(bs[by * nbx + bx_lo], bs[by * nbx + bx_hi + 1]) for by in range(by_lo, by_hi + 1)])
        got = []
        nchunks = (len(want) + S - 1) // S
        for ch in range(nchunks):
            runs = plan(lib, bs, nbx, by_lo, by_hi, bx_lo, bx_hi, ch)
            size = min(S, len(want) - ch * S)
            off = 0
            for d, s, c in runs:                       # the runs fill the stage front to back without gaps or overlap
                assert d == off and c > 0
                got.append(np.arange(s, s + c))
                off += c
            assert off == size
            assert len(runs) <= by_hi - by_lo + 1      # at most one bulk copy per bin row
        got = np.concatenate(got) if got else np.zeros(0, dtype=np.int64)
        assert np.array_equal(got, want)
        assert plan(lib, bs, nbx, by_lo, by_hi, bx_lo, bx_hi, nchunks) == []      # past the end: nothing to copy


def test_bad_arguments_are_rejected():
    import b3d
    bs = np.zeros(5, dtype=np.int32)
    assert b3d.lib.b3d_pc_stream_plan(bs.ctypes.data_as(ctypes.c_void_p), 2, 0, 0, 1, 2, 0, None, None, None, 0) < 0   # bx_hi >= nbx
    assert b3d.lib.b3d_pc_stream_plan(None, 2, 0, 0, 0, 1, 0, None, None, None, 0) < 0
