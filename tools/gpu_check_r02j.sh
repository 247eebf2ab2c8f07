#!/bin/bash
# round 2, call J (KMCPG_DEV build): TMA form with whole rows per bulk copy; long-query probe kernels with dynamic task scheduling (tests + full bench at N=1)
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "long_reads or count_codes or four_hash or sketch or full_size_c2 or c4_shape or dedup or degenerate" ) > gpurun_out/r02j_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r02j_tests.log
for b in 1; do
  echo "== KMCPG_PROBE_BULK=$b (whole rows per copy): C4 shape, 3 GB index (oracle sample), then 26 GB index"
  KMCPG_PROBE_BULK=$b NG=85205 GL=100000 NR=100000 NCHK=200 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02j.err
  KMCPG_PROBE_BULK=$b NG=85205 GL=875000 NR=100000 NCHK=0 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02j.err
  echo "== KMCPG_PROBE_BULK=$b: C2"
  KMCPG_PROBE_BULK=$b timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gtdb 2>> gpurun_out/r02j.err | python -c "
import json,sys
a=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, 'frac', round(a['roofline']['frac'],4), 'launch_ms', round(a['roofline']['avg_launch_ms'],3), 'hits', a['config']['hits_per_step'])"
done 2>&1 | tee gpurun_out/r02j_tma_ab.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02j_bench.json 2>> gpurun_out/r02j.err
python - <<'P'
import json
a=json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['e2e']['value'], a['cpu_baseline'])
for k in ('gtdb_scale','c5_hifi','c3_fracminhash'):
    g=a[k]; print(k, {x:g[x] for x in g if x not in ('per_rank','workload','digest_note')}); 
    if 'per_rank' in g: print([(round(r['probe_GBps']), round(r.get('prep_ms',0),1)) for r in g['per_rank']])
P
tail -5 gpurun_out/r02j.err
