#!/bin/bash
# Multi-GPU call of round 2 (gpurun --gpus 8, charged 8x: keep it short): the replica split that round 1 could not time any more,
# next to the column-range shards, and bench.py under torchrun at N=8.
# usage: /usr/local/graft/bin/gpurun --gpus 8 --timeout 240 -- 'bash tools/gpu_check_r02b.sh'
mkdir -p gpurun_out
( time GPUS=0,1,2,3,4,5,6,7 WORLDS=2,4,8 MODES=replicas,shard timeout 90 python tools/sharded_scale.py ) > gpurun_out/r02b_scale_8gpu.json 2> gpurun_out/r02b_scale_8gpu.err
echo "scale exit $?"; cat gpurun_out/r02b_scale_8gpu.json; tail -5 gpurun_out/r02b_scale_8gpu.err
( time timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 ) \
    > gpurun_out/r02b_bench_n8.json 2> gpurun_out/r02b_bench_n8.err
echo "bench n8 exit $?"; tail -c 700 gpurun_out/r02b_bench_n8.json; tail -5 gpurun_out/r02b_bench_n8.err
