#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): 2-GPU equivalence test, weak-scaling bench with the fused SyncBN path and with the
# NCCL fallback, cfg5 / cfg4 where N allows.
N=${1:-2}; T=${2:-r2mg}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
if [ "$N" = "2" ]; then
  python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -p no:cacheprovider --timeout 1200 > gpurun_out/${T}_pytest_n$N.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${T}_pytest_n$N.log
fi
B3D_BENCH_NO_CPU=1 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n${N}_fused.json 2> gpurun_out/${T}_bench_n${N}_fused.err
B3D_SYNC_FUSED=0 timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n${N}_nccl.json 2> gpurun_out/${T}_bench_n${N}_nccl.err
if [ "$N" = "8" ]; then
  B3D_BENCH_NO_CPU=1 python bench.py --gpus 1 --workload cfg5 --steps 10 --warmup 3 > gpurun_out/${T}_cfg5_n1.json 2> gpurun_out/${T}_cfg5_n1.err
  timeout 600 $RUN bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 > gpurun_out/${T}_cfg5_n${N}.json 2> gpurun_out/${T}_cfg5_n${N}.err
fi
if [ "$N" = "4" ]; then
  B3D_BENCH_NO_CPU=1 python bench.py --gpus 1 --workload cfg4 --steps 10 --warmup 3 > gpurun_out/${T}_cfg4_n1.json 2> gpurun_out/${T}_cfg4_n1.err
  timeout 600 $RUN bench.py --gpus $N --workload cfg4 --steps 10 --warmup 3 > gpurun_out/${T}_cfg4_n${N}.json 2> gpurun_out/${T}_cfg4_n${N}.err
fi
for f in gpurun_out/${T}_*n*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, d["e2e"]["value"], d["config"].get("cuda_graph"))
except Exception as e:
    print("parse failed:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
done
[ -f gpurun_out/${T}_pytest_n$N.log ] && grep -E "passed|failed|rc=|^E  |1-GPU vs" gpurun_out/${T}_pytest_n$N.log | head -20
grep -h "b3d.sync" gpurun_out/${T}_*.err | head -3
du -sh gpurun_out
