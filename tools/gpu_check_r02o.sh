#!/bin/bash
# round 2, call O (8 GPUs): bench.py under torchrun at N=8 with the GTDB-shape index at FULL size (genome length 3.5 Mb: ~105 GB, 13 GB per GPU)
mkdir -p gpurun_out
( time KMCP_GTDB_GL=3500000 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 ) \
    > gpurun_out/r02o_bench_n8_fullgtdb.json 2> gpurun_out/r02o_bench_n8.err
echo "bench n8 exit $?"; python - <<'P'
import json
a=json.loads(open('gpurun_out/r02o_bench_n8_fullgtdb.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['e2e']['value'], a['e2e']['ms_per_step'])
for k in ('gtdb_scale','c5_hifi'):
    g=a[k]; print(k, {x:g[x] for x in g if x not in ('per_rank','digest_note','roofline')})
    print([(round(r['probe_GBps']), round(r.get('prep_ms',0),1), r.get('build_s')) for r in g['per_rank']])
P
tail -5 gpurun_out/r02o_bench_n8.err
