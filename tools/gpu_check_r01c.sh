#!/bin/bash
# one short GPU call: the sharded-engine tests first (new code), then the rest of the GPU suite while time remains
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r01c_gpu.txt 2>&1
( time timeout 200 python -m pytest tests/test_gpu_sharded.py -x -q -p no:cacheprovider ) > gpurun_out/r01c_sharded.log 2>&1
echo "sharded exit $?" >> gpurun_out/r01c_sharded.log
tail -5 gpurun_out/r01c_sharded.log
( time timeout 90 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r01c_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r01c_smoke.log
tail -3 gpurun_out/r01c_smoke.log
( time timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_refcounts.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r01c_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/r01c_parity.log
tail -5 gpurun_out/r01c_parity.log
