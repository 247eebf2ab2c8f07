#!/bin/bash
# round 2, call D (KMCPG_DEV build): executor thread + hash stream + submit/wait; row indices derived in the probe kernel (Barrett)
# vs precomputed by locs_kernel, C2 and C4 shape; the new round-2 tests
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02d_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -25 gpurun_out/r02d_gpu_tests.log
for cfg in "X=1" "KMCPG_PROBE_LOCS=buffer"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02d_bench.err | python -c "
import json,sys
a=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step','stage_ms_per_step')}, 'frac', a['roofline']['frac'], 'launch_ms', a['roofline']['avg_launch_ms'], 'e2e', a['e2e']['value'])"
  env $cfg NG=85205 GL=100000 NR=100000 NCHK=0 timeout 200 python tools/c4_shape.py 2>> gpurun_out/r02d_bench.err
done 2>&1 | tee gpurun_out/r02d_ab.log
tail -5 gpurun_out/r02d_bench.err
