"""Which aten ops / kernels make up the non-conv time of one G step + two D steps (eager, B=32, 256^2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dimageto3dmodel_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
sys.argv = [sys.argv[0]]
import bench
from gan_training import GANTrainer
B = 32
tr = GANTrainer(bench.gan_args())
d = {k: v.cuda() for k, v in bench.gan_host_inputs(B, 0, False).items()}
for _ in range(2):
    tr.g_step(d["X_alpha"], d["C"]); tr.d_step(d["X_tex"], d["X_alpha"], d["X_mesh"], d["C"])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.g_step(d["X_alpha"], d["C"]); tr.d_step(d["X_tex"], d["X_alpha"], d["X_mesh"], d["C"]); tr.d_step(d["X_tex"], d["X_alpha"], d["X_mesh"], d["C"])
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.self_device_time_total > 0]
ev.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ev)
print(f"total device time {tot/1e3:.1f} ms")
for e in ev[:34]:
    print(f"{e.self_device_time_total/1e3:8.2f} ms {100*e.self_device_time_total/tot:5.1f}%  x{e.count:4d}  {e.key[:110]}")

print("---- aten ops by input shape (self device time)")
ev = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0 and e.key.startswith("aten::")]
ev.sort(key=lambda e: -e.self_device_time_total)
print(f"aten total {sum(e.self_device_time_total for e in ev)/1e3:.1f} ms")
for e in ev[:45]:
    print(f"{e.self_device_time_total/1e3:8.3f} ms x{e.count:4d}  {e.key:28s} {str(e.input_shapes)[:120]}")
