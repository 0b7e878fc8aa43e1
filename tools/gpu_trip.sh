#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu captures.  Everything lands in gpurun_out/ (kept well below 64 MiB:
# ncu reports are exported to CSV on the box and deleted).
T=${1:-r2c}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rf --tb=short ${PYTEST_ARGS} > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --profile-from-start off \
    -k regex:'pc_|mesh_|chamfer|cbn_|pad_leaky|fold_rows|wrap_x|rgba|flat_loss|bn_stats|bank_|vertex_' -c 70 -f \
    -o gpurun_out/${T}_nonconv python tools/ncu_r2_step.py > gpurun_out/${T}_ncu.log 2>&1
  ncu -i gpurun_out/${T}_nonconv.ncu-rep --page raw --csv > gpurun_out/${T}_nonconv_raw.csv 2>/dev/null
  rm -f gpurun_out/${T}_nonconv.ncu-rep
  [ "${LAUNCHES:-1}" = "1" ] && B3D_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_launches_bench.log 2>&1
fi
if [ "${EXTRA:-0}" = "1" ]; then
  B3D_BENCH_NO_CPU=1 python bench.py --workload cfg4 --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err
  B3D_BENCH_NO_CPU=1 python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err
  for w in cfg4 cfg5; do echo "== $w"; cut -c1-400 gpurun_out/${T}_bench_$w.json; tail -n 3 gpurun_out/${T}_bench_$w.err | cut -c1-300; done
fi
if [ "${NCUCONV:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none -k regex:'conv_|wgrad_' -c 40 -f -o gpurun_out/${T}_convs python tools/ncu_convs_r2.py > gpurun_out/${T}_ncu_convs.log 2>&1
  ncu -i gpurun_out/${T}_convs.ncu-rep --page raw --csv > gpurun_out/${T}_convs_raw.csv 2>/dev/null; rm -f gpurun_out/${T}_convs.ncu-rep
  timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'mesh_raster_bwd' -c 1 -f -o gpurun_out/${T}_meshbwd python tools/ncu_r2_step.py > gpurun_out/${T}_ncu_meshbwd.log 2>&1
  ncu -i gpurun_out/${T}_meshbwd.ncu-rep --page source --csv > gpurun_out/${T}_meshbwd_source.csv 2>/dev/null; rm -f gpurun_out/${T}_meshbwd.ncu-rep
fi
if [ "${DIAG:-0}" = "1" ]; then
  python tools/time_convs.py > gpurun_out/${T}_convs_rowwin1.txt 2>&1
  ${DIAG_ENV:-B3D_CONV_ROWWIN=0} python tools/time_convs.py > gpurun_out/${T}_convs_rowwin0.txt 2>&1
  echo "== conv layer times, default | ${DIAG_ENV:-B3D_CONV_ROWWIN=0}"; paste -d'|' <(cut -c1-82 gpurun_out/${T}_convs_rowwin1.txt) <(cut -c24-82 gpurun_out/${T}_convs_rowwin0.txt) | head -24
fi
if [ "${REFARM:-0}" = "1" ]; then
  SECONDS=0
  python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
  echo "== reference arm: ${SECONDS} s wall"; cut -c1-260 gpurun_out/${T}_bench_reference.json; tail -n 2 gpurun_out/${T}_bench_reference.err
fi
echo "==== pytest"; grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest.log | tail -n 30
grep -E "^E  " gpurun_out/${T}_pytest.log | head -n 30
echo "==== bench"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"])
    ks = d["kernel_ms_per_step"]; print("libb3d ms/step", round(sum(ks.values()), 2), dict(list(ks.items())[:14]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${T}_bench.err").read()[-1500:])
PY
du -sh gpurun_out
