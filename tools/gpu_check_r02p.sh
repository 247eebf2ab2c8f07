#!/bin/bash
# round 2, call P (product build): eight-lane hash kernel for short reads, compact hit sort keys, G3/G4 golden tables through the CUDA path:
# the whole GPU suite, then bench at N=1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02p_tests.log 2>&1
echo "gpu tests exit $?"; grep -E "passed|failed|error" gpurun_out/r02p_tests.log | tail -3; grep -B30 "short test summary" gpurun_out/r02p_tests.log | head -60
( time timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gtdb ) > gpurun_out/r02p_bench.json 2> gpurun_out/r02p.err
python - <<'P'
import json
a=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step','stage_ms_per_step')}, a['roofline']['frac'], a['roofline']['whole_step_frac'], a['e2e']['value'], a['config']['hits_per_step'])
P
tail -3 gpurun_out/r02p.err
