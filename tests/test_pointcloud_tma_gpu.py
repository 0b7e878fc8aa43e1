"""The silhouette kernels read the bin records either with plain loads or through a cp.async.bulk (TMA 1-D bulk copy)
shared-memory ring (csrc/pc_kernels.cu: BinStream; B3D_PC_TMA).  The switch is read once per process, so the
whole point-cloud parity suite (goldens from the reference's classes, oracle parity at cfg2 size, bit-exact index buffers)
is re-run in a child process with the path that is NOT this process's default."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_pointcloud_suite_on_the_other_record_path():
    import b3d
    other = "0" if b3d.lib.b3d_pc_tma_staging() else "1"
    env = dict(os.environ, B3D_PC_TMA=other)
    files = [os.path.join(ROOT, "tests", f) for f in ("test_pointcloud_gpu.py", "test_silhouette_losses.py")]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout", "100"] + files,
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, f"B3D_PC_TMA={other}:\n{r.stdout[-3000:]}\n{r.stderr[-2000:]}"
    assert " passed" in r.stdout
