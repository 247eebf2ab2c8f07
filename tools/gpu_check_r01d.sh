#!/bin/bash
# 2-GPU call: bench.py at N=2 under torchrun (NCCL broadcast + padded gather path), then the in-process sharded engine on two devices
mkdir -p gpurun_out
( time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r01d_bench_n2.json 2> gpurun_out/r01d_bench_n2.err
echo "bench n2 exit $?"; tail -c 600 gpurun_out/r01d_bench_n2.json; tail -5 gpurun_out/r01d_bench_n2.err
( time GPUS=0,1 timeout 120 python tools/sharded_scale.py ) > gpurun_out/r01d_sharded_scale_2gpu.json 2> gpurun_out/r01d_sharded_scale_2gpu.err
echo "sharded scale exit $?"; cat gpurun_out/r01d_sharded_scale_2gpu.json; tail -5 gpurun_out/r01d_sharded_scale_2gpu.err
