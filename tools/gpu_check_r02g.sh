#!/bin/bash
# round 2, call G (2 GPUs): bench.py under torchrun at N=2 (broadcast + hit return inside the timed region, GTDB-scale block sharded over 2 ranks),
# the in-process sharded engine (one H2D + peer copies) and the replica split
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 ) \
    > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
echo "bench n2 exit $?"; tail -c 3000 gpurun_out/r02g_bench_n2.json; tail -5 gpurun_out/r02g_bench_n2.err
( time timeout 300 python -m pytest tests/test_zz_gpu_sharded.py -m gpu -x -q ) > gpurun_out/r02g_sharded_tests.log 2>&1
echo "sharded tests exit $?"; tail -5 gpurun_out/r02g_sharded_tests.log
( time GPUS=0,1 WORLDS=2 MODES=shard,replicas NR=2000000 timeout 200 python tools/sharded_scale.py ) > gpurun_out/r02g_scale_2gpu.json 2> gpurun_out/r02g_scale_2gpu.err
echo "scale exit $?"; cat gpurun_out/r02g_scale_2gpu.json; tail -5 gpurun_out/r02g_scale_2gpu.err
