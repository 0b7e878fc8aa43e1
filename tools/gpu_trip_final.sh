#!/bin/bash
# End-of-round visit: smoke, the whole -m gpu suite, the default bench line, the ncu launch list of the same command and an
# `ncu --set full` capture of the row-window / weight-gradient kernels inside the real step.  Outputs: gpurun_out/${T}_*.
T=${1:-r2z}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rf --tb=short > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
B3D_BENCH_NO_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv \
  --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"conv_rowwin_tf32_kernel|wgrad_tf32_kernel" -c 36 -f \
  -o gpurun_out/${T}_tc python tools/ncu_r2_step.py > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_tc.ncu-rep --page raw --csv > gpurun_out/${T}_tc_raw.csv 2>/dev/null
rm -f gpurun_out/${T}_tc.ncu-rep
tail -n 2 gpurun_out/${T}_smoke.log
echo "==== pytest"; grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest.log | tail -n 20
echo "==== bench"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"],
          d["clocks"], d.get("cpu_baseline"))
    ks = d["kernel_ms_per_step"]; print("libb3d ms/step", round(sum(ks.values()), 2))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${T}_bench.err").read()[-1500:])
PY
du -sh gpurun_out
