#!/bin/bash
# round 2, call L (8 GPUs, charged 8x: short): the in-process forms — one database sharded over the contexts of one process (batch staged once,
# peer copies) and the replica split — at 2, 4, 8 devices, 1 M and 4 M reads per call
mkdir -p gpurun_out
( time GPUS=0,1,2,3,4,5,6,7 WORLDS=2,4,8 MODES=shard,replicas NR=1000000 REPS=3 timeout 120 python tools/sharded_scale.py ) > gpurun_out/r02l_scale_1M.json 2> gpurun_out/r02l_scale_1M.err
echo "scale 1M exit $?"; cat gpurun_out/r02l_scale_1M.json; tail -3 gpurun_out/r02l_scale_1M.err | cut -c1-300
( time GPUS=0,1,2,3,4,5,6,7 WORLDS=4,8 MODES=shard,replicas NR=4000000 REPS=3 timeout 150 python tools/sharded_scale.py ) > gpurun_out/r02l_scale_4M.json 2> gpurun_out/r02l_scale_4M.err
echo "scale 4M exit $?"; cat gpurun_out/r02l_scale_4M.json; tail -3 gpurun_out/r02l_scale_4M.err | cut -c1-300
