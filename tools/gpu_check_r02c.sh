#!/bin/bash
# round 2, call C: the probe kernel now derives the row indices itself (no locs kernel, no locs buffer): parity + speed
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -5 gpurun_out/r02c_gpu_tests.log
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench exit $?"; python - <<'P'
import json
a=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step','stage_ms_per_step')}, a['roofline']['frac'], a['roofline']['avg_launch_ms'], a['e2e']['value'])
P
( time NG=85205 GL=100000 NR=100000 NCHK=200 timeout 300 python tools/c4_shape.py ) > gpurun_out/r02c_c4_shape.json 2> gpurun_out/r02c_c4_shape.err
echo "c4 exit $?"; cat gpurun_out/r02c_c4_shape.json; tail -3 gpurun_out/r02c_c4_shape.err
