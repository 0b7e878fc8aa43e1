#!/bin/bash
# Last GPU visit of round 2 (4.8 GPU-minutes left): the whole -m gpu suite at HEAD (FID path, refactored point kernels, the
# child-process run of the TMA-staged record path), the point-kernel A/B, the default bench line.  Most important first.
T=${1:-r2y}
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 100 -rf --tb=short --durations=8 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest.log | tail -n 25
rm -f /tmp/pc_ab.pt
B3D_PC_TMA=0 timeout 40 python tools/time_pc.py /tmp/pc_ab.pt > gpurun_out/${T}_pc.log 2>&1
B3D_PC_TMA=1 timeout 40 python tools/time_pc.py /tmp/pc_ab.pt >> gpurun_out/${T}_pc.log 2>&1
tail -n 4 gpurun_out/${T}_pc.log | cut -c1-600
B3D_BENCH_NO_CPU=1 timeout 60 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-330 gpurun_out/${T}_bench.json; tail -n 2 gpurun_out/${T}_bench.err | cut -c1-300
