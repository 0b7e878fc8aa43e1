"""Device timing of the FID evaluation path on one GPU (development aid): the Inception-v3 extractor alone
(batch 32 at 299 x 299; dense-conv FLOPs counted from the layer table) and one evaluation batch end to end (running-average
generator -> fused vertex pipeline -> 299 x 299 render -> Inception -> feature sums).  CUDA events, median of 10 after 3 warm-ups."""
import os
import statistics
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import b3d  # noqa: E402
import gan_common as GC  # noqa: E402
from fid_common import randomize_inception  # noqa: E402
from fid_evaluation import FIDEvaluator  # noqa: E402
from models import gan  # noqa: E402
from rendering.mesh_template import MeshTemplate  # noqa: E402
from tools.uvsphere import write_uvsphere_obj  # noqa: E402
from utils.inception import InceptionV3  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 32))
inc = randomize_inception(InceptionV3([3], weights=None), 0).to(dev)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


# dense-conv FLOPs per image: hook the conv entry point once and read the geometry from its arguments
flops = [0]
orig = b3d.lib.b3d_conv2d_tf32


def counting(*a):
    n, hout, wout, cout, ntaps = a[4], a[8], a[9], a[10], a[11]
    flops[0] += 2 * n * hout * wout * cout * a[7] * ntaps          # as executed (padded channels, 9-tap pooled branches)
    return orig(*a)


x = torch.rand(B, 3, 299, 299, device=dev)
b3d.lib.b3d_conv2d_tf32 = counting
n0 = b3d.launch_count()
inc(x)
launches = b3d.launch_count() - n0
b3d.lib.b3d_conv2d_tf32 = orig
ms = timed(lambda: inc(x))
print(f"inception B={B}: {ms:.2f} ms per batch = {B / ms * 1e3:.0f} img/s, {flops[0] / B / 1e9:.2f} GF/img executed, "
      f"{flops[0] / ms / 1e9:.1f} TF/s, {launches} libb3d launches per forward")

args = GC.make_args(256, 2)
G, _ = GC.build(gan, args)
G.to(dev).eval()
mt = MeshTemplate(write_uvsphere_obj(os.path.join(tempfile.mkdtemp(), "uvsphere_16rings.obj"), rings=16), device=dev)
ev = FIDEvaluator(G, mt, inception=inc, device=dev)
g = torch.Generator().manual_seed(0)
batch = {"idx": torch.arange(B), "class": torch.randint(0, 200, (B, 1), generator=g),
         "rotation": torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1), "scale": 0.5 + 0.3 * torch.rand(B, generator=g),
         "translation": (torch.rand(B, 3, generator=g) - 0.5) * 0.2, "image": torch.rand(B, 3, 299, 299, generator=g)}
batch = {k: v.to(dev) for k, v in batch.items()}
from fid_evaluation import truncated_noise  # noqa: E402
from utils.fid import FIDStatistics  # noqa: E402
st = FIDStatistics(2048, dev)


@torch.no_grad()
def one_batch():
    noise = truncated_noise(B, 64, 1.0).to(dev)
    pred_tex, pred_mesh_map, _ = G(noise, batch["class"], None, return_attention=True)
    ev.render_and_score(pred_mesh_map, pred_tex, batch, st)


ms = timed(one_batch)
print(f"evaluate_fid, one batch of {B} on the device (truncated noise -> generator 256^2 -> vertex pipeline -> render 299^2 -> "
      f"inception -> feature sums): {ms:.1f} ms = {B / ms * 1e3:.0f} img/s")
