#!/bin/bash
# Last GPU visit of round 2 (4.8 GPU-minutes left).  Most important first: the tests of everything that changed since the last
# full green run (r2z: only pc_kernels.cu was edited and fid_kernels.cu added; the other objects are byte-identical), the
# point-kernel A/B, the default bench line, the FID timing, then the rest of the suite and an ncu capture with whatever is left.
T=${1:-r2y}
mkdir -p gpurun_out
NEW="tests/test_fid_gpu.py tests/test_pointcloud_gpu.py tests/test_pointcloud_dense_gpu.py tests/test_pointcloud_tma_gpu.py tests/test_silhouette_losses.py tests/test_abi.py"
timeout 150 python -m pytest $NEW -m gpu -q -p no:cacheprovider --timeout 100 -rf --tb=short --durations=6 > gpurun_out/${T}_pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest_new.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest_new.log | tail -n 25
rm -f /tmp/pc_ab.pt
B3D_PC_TMA=0 timeout 40 python tools/time_pc.py /tmp/pc_ab.pt > gpurun_out/${T}_pc.log 2>&1
B3D_PC_TMA=1 timeout 40 python tools/time_pc.py /tmp/pc_ab.pt >> gpurun_out/${T}_pc.log 2>&1
tail -n 4 gpurun_out/${T}_pc.log | cut -c1-700
B3D_BENCH_NO_CPU=1 timeout 60 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-330 gpurun_out/${T}_bench.json; tail -n 2 gpurun_out/${T}_bench.err | cut -c1-300
timeout 40 python tools/time_fid.py > gpurun_out/${T}_fid.log 2>&1; tail -n 3 gpurun_out/${T}_fid.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${T}_smoke.log | cut -c1-300
IGN=""; for f in $NEW; do IGN="$IGN --ignore=$f"; done
timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 100 -rf --tb=short $IGN > gpurun_out/${T}_pytest_rest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest_rest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest_rest.log | tail -n 10
B3D_PC_TMA=1 ITERS=7 timeout 80 ncu --set full --clock-control none -k regex:"pc_sil" -s 8 -c 2 -f -o gpurun_out/${T}_pc python tools/time_pc.py > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_pc.ncu-rep --page raw --csv > gpurun_out/${T}_pc_raw.csv 2>/dev/null; rm -f gpurun_out/${T}_pc.ncu-rep
ls -la gpurun_out/${T}_* | head -20
