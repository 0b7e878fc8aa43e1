#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu captures.  Everything lands in gpurun_out/ (kept < 64 MiB).
T=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'pc_|mesh_|chamfer|cbn_|pad_leaky|fold_rows|wrap_x|rgba|flat_loss|bn_stats|bank_' -c 100 -f \
    -o gpurun_out/${T}_nonconv python tools/ncu_r2_step.py > gpurun_out/${T}_ncu.log 2>&1
  B3D_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_launches_bench.log 2>&1
fi
echo "==== pytest"; tail -n 12 gpurun_out/${T}_pytest.log
echo "==== bench"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["roofline"], d.get("chamfer"))
    ks = d["kernel_ms_per_step"]; print("libb3d ms/step", round(sum(ks.values()), 2), dict(list(ks.items())[:12]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${T}_bench.err").read()[-1500:])
PY
du -sh gpurun_out
