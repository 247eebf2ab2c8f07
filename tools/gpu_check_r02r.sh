#!/bin/bash
# round 2, call R (product build): ncu --set full of the h=3 probe kernel on the GTDB-shape index (26 GB), and the final whole-suite run of the round
mkdir -p gpurun_out
( time NG=85205 GL=875000 NR=100000 NCHK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -s 40 -c 2 -o gpurun_out/probe_gtdb_r02 \
    python tools/c4_shape.py ) > gpurun_out/r02r_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r02r_ncu.log | cut -c1-400; ls -la gpurun_out/probe_gtdb_r02.ncu-rep
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02r_tests.log 2>&1
echo "gpu tests exit $?"; grep -E "passed|failed|error" gpurun_out/r02r_tests.log | tail -3
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -2
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02r_bench.json 2> gpurun_out/r02r.err
python - <<'P'
import json
a=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step','stage_ms_per_step')}, a['roofline']['frac'], a['roofline']['whole_step_frac'], a['e2e']['value'], a['cpu_baseline'])
for k in ('gtdb_scale','c5_hifi','c3_fracminhash'):
    g=a[k]; print(k, {x:g[x] for x in g if x not in ('per_rank','digest_note','roofline','workload')})
P
tail -3 gpurun_out/r02r.err
