#!/bin/bash
# 8-GPU call (charged 8x): in-process sharded engine, C2 index cut by column range over 2, 4 and 8 devices
mkdir -p gpurun_out
( time GPUS=0,1,2,3,4,5,6,7 WORLDS=2,4,8 timeout 28 python tools/sharded_scale.py ) > gpurun_out/r01e_sharded_scale_8gpu.json 2> gpurun_out/r01e_sharded_scale_8gpu.err
echo "exit $?"; cat gpurun_out/r01e_sharded_scale_8gpu.json; tail -8 gpurun_out/r01e_sharded_scale_8gpu.err
