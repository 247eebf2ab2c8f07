#!/bin/bash
# round 2, call K (product build): ncu launch list and one --set full capture of the probe kernel (same command), then the long-read checks
# after the per-warp task draw, and the C5 / C3 blocks of bench.py
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gtdb"
( time timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kmcpg|DeviceRadixSort|DeviceScan|DeviceSegmented' -s 2500 -c 400 \
    --csv --log-file gpurun_out/launches_r02.csv $CMD ) > gpurun_out/r02k_ncu_list.log 2>&1
echo "launch list exit $?"; tail -3 gpurun_out/r02k_ncu_list.log | cut -c1-300; wc -l gpurun_out/launches_r02.csv
( time timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -s 16 -c 2 -o gpurun_out/probe_r02 $CMD ) > gpurun_out/r02k_ncu_full.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/r02k_ncu_full.log | cut -c1-300; ls -la gpurun_out/probe_r02.ncu-rep
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "long_reads or count_codes or four_hash or sketch or dedup or degenerate or refcounts" ) > gpurun_out/r02k_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r02k_tests.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02k_bench.json 2> gpurun_out/r02k.err
python - <<'P'
import json
a=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['e2e']['value'])
for k in ('gtdb_scale','c5_hifi','c3_fracminhash'):
    g=a[k]; print(k, {x:g[x] for x in g if x not in ('per_rank','workload','digest_note','roofline')})
    if 'per_rank' in g: print([(round(r['probe_GBps']), round(r.get('prep_ms',0),1)) for r in g['per_rank']])
P
tail -3 gpurun_out/r02k.err
