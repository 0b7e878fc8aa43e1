"""Generate tests/golden/fid_reference.npz from the REFERENCE's own Python (authoring container only):
    python tests/golden/make_golden_fid.py
Imports, unmodified, /root/reference/code/utils/fid.py (calculate_stats, calculate_frechet_distance).  Its sibling
utils/inception.py imports torchvision at module scope (absent in this image, and its pretrained weights need a download),
so a placeholder `torchvision` module is registered first: nothing of it is executed by the two functions used here.
Also distils the reference's shipped statistics of the REAL CUB images (cache/cub/precomputed_fid_299x299_{train,testval}.npz,
2048-d Inception pool features, 7.9 MB each: data, never copied) into the principal 128 x 128 sub-blocks — a principal
sub-block of a covariance is the covariance of those coordinates — and records the reference's FID between them.
One execution patch (SciPy API drift), see below.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/code")
tv = types.ModuleType("torchvision")
tv.models = types.ModuleType("torchvision.models")
sys.modules["torchvision"], sys.modules["torchvision.models"] = tv, tv.models

import scipy.linalg                                                       # noqa: E402
import utils.fid as ref_fid                                               # noqa: E402  (reference)
from utils.fid import calculate_frechet_distance, calculate_stats        # noqa: E402  (reference)

# Execution patch (like P1-P3 of SURVEY App. A): utils/fid.py:67 calls linalg.sqrtm(..., disp=False) and unpacks
# (sqrt, error estimate); SciPy >= 1.16 dropped that keyword (TypeError as written).  Same function, old calling convention:
_sqrtm = scipy.linalg.sqrtm
ref_fid.linalg = types.SimpleNamespace(sqrtm=lambda a, disp=True: _sqrtm(a) if disp else (_sqrtm(a), 0.0))


def main():
    rng = np.random.default_rng(7)
    out = {}
    # 1. activations -> statistics -> distance (n > D: full-rank covariances)
    a = rng.standard_normal((300, 48)).astype(np.float32) * rng.uniform(0.2, 2.0, 48).astype(np.float32) + 0.3
    mix = rng.standard_normal((48, 48)).astype(np.float32) * 0.2 + np.eye(48, dtype=np.float32)
    b = (rng.standard_normal((260, 48)).astype(np.float32) @ mix) - 0.1
    m1, s1 = calculate_stats(a)
    m2, s2 = calculate_stats(b)
    out.update(act_a=a, act_b=b, mu_a=m1, sigma_a=s1, mu_b=m2, sigma_b=s2, fid_ab=calculate_frechet_distance(m1, s1, m2, s2),
               fid_aa=calculate_frechet_distance(m1, s1, m1, s1))
    # 2. fewer samples than dimensions (rank-deficient covariances, the `fast` evaluations of small splits)
    c = rng.standard_normal((24, 48)).astype(np.float32)
    m3, s3 = calculate_stats(c)
    out.update(act_c=c, fid_ca=calculate_frechet_distance(m3, s3, m1, s1))
    # 3. the real CUB statistics shipped with the reference, principal 128-d sub-block
    cache = "/root/reference/code/cache/cub"
    tr = np.load(os.path.join(cache, "precomputed_fid_299x299_train.npz"), allow_pickle=True)
    va = np.load(os.path.join(cache, "precomputed_fid_299x299_testval.npz"), allow_pickle=True)
    K = 128
    mt, st = tr["stats_m"], tr["stats_s"] + np.triu(tr["stats_s"].T, 1)           # main.py:171-172
    mv, sv = va["stats_m"], va["stats_s"] + np.triu(va["stats_s"].T, 1)
    out.update(cub_mu_train=mt[:K].astype(np.float64), cub_sigma_train=st[:K, :K].astype(np.float64),
               cub_mu_val=mv[:K].astype(np.float64), cub_sigma_val=sv[:K, :K].astype(np.float64),
               cub_fid_sub=calculate_frechet_distance(mt[:K].astype(np.float64), st[:K, :K].astype(np.float64),
                                                      mv[:K].astype(np.float64), sv[:K, :K].astype(np.float64)),
               cub_num_images=np.array([int(tr["num_images"]), int(va["num_images"])]))
    print("dtype of shipped stats:", tr["stats_m"].dtype, tr["stats_s"].dtype, "n =", out["cub_num_images"])
    print({k: float(v) for k, v in out.items() if k.startswith("fid") or k == "cub_fid_sub"})
    np.savez_compressed(os.path.join(HERE, "fid_reference.npz"), **out)


if __name__ == "__main__":
    main()
