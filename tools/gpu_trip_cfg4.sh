#!/bin/bash
# cfg4 (reconstruction loop, global batch 50) on 1 GPU and as DDP x N (`gpurun --gpus N -- bash tools/gpu_trip_cfg4.sh N tag`).
N=${1:-4}; T=${2:-r2cfg4}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613"
B3D_BENCH_NO_CPU=1 timeout 400 python bench.py --gpus 1 --workload cfg4 --steps 20 --warmup 5 > gpurun_out/${T}_cfg4_n1.json 2> gpurun_out/${T}_cfg4_n1.err
timeout 400 $RUN bench.py --gpus $N --workload cfg4 --steps 20 --warmup 5 > gpurun_out/${T}_cfg4_n${N}.json 2> gpurun_out/${T}_cfg4_n${N}.err
timeout 400 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_cfg3_n${N}.json 2> gpurun_out/${T}_cfg3_n${N}.err
for f in gpurun_out/${T}_*n*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, d["e2e"]["value"], d["config"], d.get("collectives"))
except Exception as e:
    print("parse failed:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
du -sh gpurun_out
