#!/bin/bash
# round 2, call E (KMCPG_DEV build): new bench.py (two jobs in flight, GTDB-scale block) under three placements of the query preparation
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q ) > gpurun_out/r02e_tests.log 2>&1
echo "round2 tests exit $?"; tail -15 gpurun_out/r02e_tests.log
show() { python -c "
import json,sys
a=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step','stage_ms_per_step')}, 'frac', round(a['roofline']['frac'],4), 'step_frac', round(a['roofline']['whole_step_frac'],4), 'launch_ms', round(a['roofline']['avg_launch_ms'],3), 'e2e', round(a['e2e']['value']), a['e2e']['breakdown_ms_per_step'])
g=a.get('gtdb_scale')
if g: print('gtdb', {k:g[k] for k in ('job_reads_per_s','ms_per_step','frac_of_roofline','hits_per_step','hit_list_digest')}, [(r['build_s'], round(r['probe_GBps'])) for r in g['per_rank']])
"; }
for cfg in "KMCPG_HASH_STREAM=0" "KMCPG_HASH_PRIO=low" "KMCPG_HASH_PRIO=high"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gtdb 2>> gpurun_out/r02e_bench.err | show
done 2>&1 | tee gpurun_out/r02e_ab.log
echo "== full default run"
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02e_bench.json 2>> gpurun_out/r02e_bench.err
show < gpurun_out/r02e_bench.json
tail -5 gpurun_out/r02e_bench.err
