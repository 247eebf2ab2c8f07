#!/bin/bash
# round 2, call Q (4 GPUs): bench.py under torchrun at N=4 (the one rank count not run by hand yet), in-process shard / replica split at 2 and 4 devices
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 ) \
    > gpurun_out/r02q_bench_n4.json 2> gpurun_out/r02q_bench_n4.err
echo "bench n4 exit $?"; python - <<'P'
import json
a=json.loads(open('gpurun_out/r02q_bench_n4.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['roofline']['whole_step_frac'], a['e2e']['value'], a['e2e']['ms_per_step'])
for k in ('gtdb_scale','c5_hifi'):
    g=a[k]; print(k, {x:g[x] for x in g if x not in ('per_rank','digest_note','roofline','workload')})
P
tail -3 gpurun_out/r02q_bench_n4.err
( time GPUS=0,1,2,3 WORLDS=2,4 MODES=shard,replicas NR=4000000 REPS=3 timeout 150 python tools/sharded_scale.py ) > gpurun_out/r02q_scale_4M.json 2> gpurun_out/r02q_scale_4M.err
echo "scale exit $?"; python - <<'P'
import json
a=json.loads(open('gpurun_out/r02q_scale_4M.json').read().strip().splitlines()[-1])
print(a['reads'], a['one_context'])
for s in a['sharded']: print('  ', s['mode'], s['world'], 'ms', s['ms'], 'speedup', s['speedup_vs_one_context'], 'gpu_ms', s['gpu_ms_max_shard'], 'post', s['post_ms'], s['matches_equal'])
P
