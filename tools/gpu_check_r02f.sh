#!/bin/bash
# round 2, call F (KMCPG_DEV build): row indices by the locs kernel on the query-preparation stream (double-buffered per block) vs derived
# in the probe kernel; whole GPU suite incl. the demo-profiling golden
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02f_tests.log 2>&1
echo "gpu tests exit $?"; tail -15 gpurun_out/r02f_tests.log
show() { python -c "
import json,sys
a=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step','stage_ms_per_step')}, 'frac', round(a['roofline']['frac'],4), 'step_frac', round(a['roofline']['whole_step_frac'],4), 'launch_ms', round(a['roofline']['avg_launch_ms'],3), 'e2e', round(a['e2e']['value']), a['e2e']['breakdown_ms_per_step'])
"; }
for cfg in "X=1" "KMCPG_HASH_STREAM=0" "KMCPG_PROBE_LOCS=kernel" "KMCPG_HASH_PRIO=high"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gtdb 2>> gpurun_out/r02f_bench.err | show
  env $cfg NG=85205 GL=100000 NR=100000 NCHK=0 timeout 200 python tools/c4_shape.py 2>> gpurun_out/r02f_bench.err
done 2>&1 | tee gpurun_out/r02f_ab.log
tail -5 gpurun_out/r02f_bench.err
