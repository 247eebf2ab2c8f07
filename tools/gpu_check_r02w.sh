#!/bin/bash
# round 2, call W (2 GPUs): bench.py under torchrun after the run-id broadcast, and the reference arm under torchrun (rank 0 works, rank 1 exits)
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 ) \
    > gpurun_out/r02w_bench_n2.json 2> gpurun_out/r02w_bench_n2.err
echo "bench n2 exit $?"; python - <<'P'
import json
a=json.loads(open('gpurun_out/r02w_bench_n2.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['e2e']['value'], a['gtdb_scale']['job_reads_per_s'], a['gtdb_scale']['hit_list_digest'], a['c5_hifi']['hit_list_digest'])
P
tail -3 gpurun_out/r02w_bench_n2.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) \
    > gpurun_out/r02w_ref_n2.json 2> gpurun_out/r02w_ref_n2.err
echo "reference n2 exit $?"; cut -c1-400 gpurun_out/r02w_ref_n2.json; tail -3 gpurun_out/r02w_ref_n2.err
