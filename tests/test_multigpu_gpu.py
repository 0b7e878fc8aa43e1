"""Hardware check of the N-GPU path: the SAME global batch of 64 through GANTrainer
on 1 GPU and on 2 GPUs (torchrun, NCCL: SyncBN statistic all-reduces + gradient all-reduce) must agree — losses, per-parameter
gradient norms, BatchNorm running statistics, spectral-norm vectors and the parameters after the Adam step.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`); skipped on a 1-GPU box.
Tolerances: different reduction order of tf32 convolutions / fp32 sums over 32 vs 64 samples per rank: 2e-3 relative;
parameters after the first Adam(0, 0.9) step move by +-lr (= 1e-4) whatever the gradient's size, so entries whose gradient
is noise-level may differ by 2 lr: absolute 2.5e-4 there."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
WORKER = os.path.join(ROOT, "tests", "mgpu", "equiv_worker.py")


def run(world, out, batch=64):
    env = dict(os.environ, TORCH_NCCL_ASYNC_ERROR_HANDLING="0")
    if world == 1:
        cmd = [sys.executable, WORKER, out, str(batch)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 400), WORKER, out, str(batch)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_equal_one_gpu_on_the_same_global_batch():
    tmp = tempfile.mkdtemp()
    a, b = os.path.join(tmp, "w1.npz"), os.path.join(tmp, "w2.npz")
    run(1, a)
    run(2, b)
    A, B = np.load(a), np.load(b)
    assert np.allclose(A["losses"], B["losses"], rtol=2e-3, atol=1e-5), (A["losses"], B["losses"])
    for key in ("g", "d"):
        assert list(A[f"{key}_names"]) == list(B[f"{key}_names"])
        ga, gb = A[f"{key}_grad_norms"], B[f"{key}_grad_norms"]
        bad = [(str(n), x, y) for n, x, y in zip(A[f"{key}_names"], ga, gb) if abs(x - y) > 2e-3 * max(x, y) + 1e-4 * ga.max()]
        assert not bad, bad[:8]
    for k, tol in (("bn_mean", 1e-4), ("bn_var", 1e-4), ("sn_u", 1e-4), ("w_probe", 2.5e-4), ("fc_probe", 2.5e-4)):
        assert np.allclose(A[k], B[k], rtol=2e-3, atol=tol), (k, float(np.abs(A[k] - B[k]).max()))
    print("1-GPU vs 2-GPU: losses", A["losses"], B["losses"], "max rel grad-norm diff",
          max(float(np.abs(A[f"{k}_grad_norms"] - B[f"{k}_grad_norms"]).max() / A[f"{k}_grad_norms"].max()) for k in ("g", "d")))
