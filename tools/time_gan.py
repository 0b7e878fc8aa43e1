"""Eager timing of the GAN iteration on one GPU (development aid): per-call libb3d times + wall per step."""
import os, sys, time, types, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
import b3d
from gan_training import GANTrainer
B = int(os.environ.get("B", 32)); R = int(os.environ.get("R", 256))
args = types.SimpleNamespace(texture_resolution=R, conditional_class=True, conditional_color=False, conditional_text=False,
                             norm_g='syncbatch', norm_d='none', n_classes=(200,), mask_output=True, texture_only=False,
                             num_discriminators=2 if R < 512 else 3, text_embedding_dim=256, latent_dim=64, loss='hinge', lr_g=1e-4,
                             lr_d=4e-4, d_steps_per_g=2, mesh_regularization=1e-4, g_running_average_alpha=0.999)
tr = GANTrainer(args)
g = torch.Generator().manual_seed(0)
X_tex = (torch.rand(B, 3, R, R, generator=g) * 2 - 1).cuda(); X_alpha = (torch.rand(B, 1, R, R, generator=g) > 0.4).float().cuda()
X_mesh = (torch.randn(B, 3, 32, 32, generator=g) * 0.05).cuda(); C = torch.randint(0, 200, (B, 1), generator=g).cuda()
for _ in range(3): tr.step(X_tex, X_alpha, X_mesh, C)
torch.cuda.synchronize()
for name, fn in (("g", lambda: tr.g_step(X_alpha, C)), ("d", lambda: tr.d_step(X_tex, X_alpha, X_mesh, C))):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    b3d.prof_enable(); fn(); pr = b3d.prof_disable()
    tot = {k: sum(v) for k, v in pr.items()}
    print(f"{name}-step B={B} R={R}: wall {statistics.median(ts):.2f} ms; libb3d conv time {sum(tot.values()):.2f} ms in {sum(len(v) for v in pr.values())} calls:",
          {k.replace('b3d_conv2d_', ''): round(v, 2) for k, v in tot.items()})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.g_step(X_alpha, C); tr.d_step(X_tex, X_alpha, X_mesh, C); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
