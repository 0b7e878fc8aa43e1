"""The C-ABI library loads on a CPU-only box and exports every symbol include/b3d.h declares."""
import ctypes
import os
import re

from conftest import PKG, ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b3d.h")).read()
    return sorted(set(re.findall(r"B3D_API[^;]*?\b(b3d_\w+)\s*\(", text)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "b3d_pc_project" in syms and "b3d_last_error" in syms
    assert len(syms) >= 10


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(PKG, "b3d", "libb3d.so"))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in b3d.h but not exported: {missing}"


def test_error_reporting_without_gpu():
    import b3d
    # argument validation happens before any CUDA call
    rc = b3d.lib.b3d_pc_project(None, None, 1, 1, 1, 1.875, 2.0, None, None, None, None, None, None, None)
    assert rc == -1
    assert b"bad sizes" in b3d.lib.b3d_last_error()
    assert b3d.lib.b3d_version() >= 100


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch
    import b3d
    from utils.effective_loss_function import EffectiveLossFunction, PointCloudRender
    assert PointCloudRender is EffectiveLossFunction
    m = EffectiveLossFunction(voxel_size=32)
    with pytest.raises(b3d.B3DError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 3), torch.ones(1, 4))
