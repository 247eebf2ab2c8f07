#!/bin/bash
# round 2, call H (8 GPUs, charged 8x: short): bench.py under torchrun at N=8 (weak C2 + GTDB-scale strong + C5 HiFi), then the in-process
# sharded engine / replica split at 2, 4, 8 devices
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 ) \
    > gpurun_out/r02h_bench_n8.json 2> gpurun_out/r02h_bench_n8.err
echo "bench n8 exit $?"; tail -c 1500 gpurun_out/r02h_bench_n8.json; tail -5 gpurun_out/r02h_bench_n8.err
( time GPUS=0,1,2,3,4,5,6,7 WORLDS=4,8 MODES=shard,replicas NR=4000000 REPS=2 timeout 150 python tools/sharded_scale.py ) > gpurun_out/r02h_scale_8gpu.json 2> gpurun_out/r02h_scale_8gpu.err
echo "scale exit $?"; cat gpurun_out/r02h_scale_8gpu.json; tail -5 gpurun_out/r02h_scale_8gpu.err
