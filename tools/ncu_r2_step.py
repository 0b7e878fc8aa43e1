"""One eager cfg3 iteration set (render+loss, chamfer, G step, D step) for an `ncu --set full` capture of the NON-conv kernels
(point cloud, mesh raster, losses, chamfer, GAN glue) at bench sizes.  Run as
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'pc_|mesh_|chamfer|cbn_act|pad_leaky|fold_rows|wrap_x|rgba|flat_loss|bn_stats|sn_|vertex' -c 90 \
      -o gpurun_out/r2_nonconv python tools/ncu_r2_step.py
(the profiler starts after one un-profiled warm-up pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts the package on sys.path)
import torch  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 32))
wl = bench.CudaWorkload(dev, bench.WORKLOADS["cfg3"])
d = {k: v.to(dev) for k, v in bench.host_inputs(B, 1234, False).items()}
d.update({k: v.to(dev) for k, v in bench.gan_host_inputs(B, 1234, False).items()})
from b3d.chamfer import nearest  # noqa: E402
g = torch.Generator().manual_seed(5)
a = (torch.rand(B, 8000, 3, generator=g) - 0.5).to(dev)
b = (torch.rand(B, 8000, 3, generator=g) - 0.5).to(dev)


def one():
    wl.render_step(d)
    nearest(a, b)
    wl.gan.g_step(d["X_alpha"], d["C"])
    wl.gan.d_step(d["X_tex"], d["X_alpha"], d["X_mesh"], d["C"])


one()
torch.cuda.synchronize()
torch.cuda.profiler.start()
one()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
