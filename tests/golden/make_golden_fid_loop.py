"""Generate tests/golden/fid_loop_reference.npz by EXECUTING the reference's FID evaluation (authoring container only):
    python tests/golden/make_golden_fid_loop.py
`evaluate_fid` (main.py:188-412) is compiled unmodified from the script's syntax tree, together with `ModelWrapper` (:449-526,
execution patch D15) for its 'inference' mode, into a namespace holding the module globals it reads: args (evaluate = True: the
seeded branch :223-227 / :361-362), the reference's own utils/fid.py functions (forward_inception_batch, calculate_stats,
calculate_frechet_distance — SciPy API patch as in make_golden_fid.py), rendering.utils.qrot, tqdm = identity, a log function,
and stand-ins for what needs the GPU / kaolin / ImageNet weights: tiny generator (wrapper_common.TinyG), template + renderer
(recon_step_common.Template) and a 64-feature extractor (fid_loop_common.Extractor).  Run 1: real statistics unknown (computed
from the images), pseudo-ground-truth present -> combined / texture-only / mesh-only scores.  Run 2: validation statistics set
-> the *_val scores on the seeded subset.  Run 3: fast.  Every value `calculate_frechet_distance` returned is recorded in call
order (the log lines round to 0.01).  Nothing of the reference is copied into the repository — only its outputs."""
import ast
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
tv = types.ModuleType("torchvision")
tv.models = types.ModuleType("torchvision.models")
sys.modules["torchvision"], sys.modules["torchvision.models"] = tv, tv.models
import scipy.linalg                                            # noqa: E402
import fid_loop_common as FL                                   # noqa: E402
import recon_step_common as RS                                 # noqa: E402
import wrapper_common as WC                                    # noqa: E402
import utils.fid as ref_fid                                    # noqa: E402  (reference)
from rendering.utils import qrot                               # noqa: E402  (reference)
from utils.losses import GANLoss                               # noqa: E402  (reference)

_sqrtm = scipy.linalg.sqrtm
ref_fid.linalg = types.SimpleNamespace(sqrtm=lambda a, disp=True: _sqrtm(a) if disp else (_sqrtm(a), 0.0))
torch.Tensor.cuda = lambda self, *a, **k: self                 # D15


def main():
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    defs = {n.name: n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
    wrapper = ast.get_source_segment(src, defs["ModelWrapper"]).replace(
        "GANLoss(args.loss, tensor=torch.cuda.FloatTensor).cuda()", "GANLoss(args.loss, tensor=torch.FloatTensor)")
    args = WC.make_args(2, 512)
    args.conditional_class, args.evaluate, args.tensorboard, args.truncation_sigma = True, True, False, 1.0
    ns = {"torch": torch, "nn": nn, "math": math, "np": np, "GANLoss": GANLoss, "args": args}
    exec(compile(wrapper, "main.py", "exec"), ns)
    exec(compile(ast.Module(body=[defs["evaluate_fid"]], type_ignores=[]), "main.py", "exec"), ns)
    gi, _ = WC.build()
    trainer = ns["ModelWrapper"](gi, None)
    fds, logs = [], []

    def recording_fd(*a, **k):
        v = ref_fid.calculate_frechet_distance(*a, **k)
        fds.append(float(v))
        return v
    data = FL.eval_set()
    ns.update(trainer=trainer, generator_running_avg=trainer.generator_running_avg, gpu_ids=[0], mesh_template=RS.Template(map_size=8),
              renderer=None, inception_model=FL.Extractor(), forward_inception_batch=ref_fid.forward_inception_batch,
              calculate_stats=ref_fid.calculate_stats, calculate_frechet_distance=recording_fd, qrot=qrot, tqdm=lambda x: x,
              log=logs.append, eval_loader=data, train_ds=list(range(18)), evaluation_res=FL.RES,
              m_real_train=None, s_real_train=None, m_real_val=None, s_real_val=None, n_images_val=None)
    out = {}
    ret = ns["evaluate_fid"](None, 0)
    out["run1"] = np.array(fds)                                # combined, texture-only, mesh-only
    assert len(fds) == 3 and abs(ret - fds[0]) < 1e-12
    out["m_real"], out["s_real"] = ns["m_real_train"], ns["s_real_train"]
    # validation statistics: any fixed Gaussian does; the subset of 11 generated images is drawn with np.random.seed(1234)
    rng = np.random.default_rng(3)
    a = rng.standard_normal((40, 64)) * 0.3
    ns["m_real_val"], ns["s_real_val"], ns["n_images_val"] = a.mean(0), np.cov(a, rowvar=False), 11
    out["m_val"], out["s_val"] = ns["m_real_val"], ns["s_real_val"]
    del fds[:]
    ns["evaluate_fid"](None, 0)
    out["run2"] = np.array(fds)                                # combined, texture-only, mesh-only, then the three *_val scores
    assert len(fds) == 6
    del fds[:]
    ns["evaluate_fid"](None, 0, fast=True)
    out["run3"] = np.array(fds)
    assert len(fds) == 1
    out["logs"] = np.array(logs)
    p = os.path.join(HERE, "fid_loop_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes;", out["run1"], out["run2"], out["run3"])


if __name__ == "__main__":
    main()
