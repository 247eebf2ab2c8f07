#!/bin/bash
# round 2, call U (product build, last of the round): the whole GPU suite and smoke() on the library as committed
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02u_tests.log 2>&1
echo "gpu tests exit $?"; grep -E "passed|failed|error" gpurun_out/r02u_tests.log | tail -3
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -2
( time timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gtdb ) 2>/dev/null | python -c "
import json,sys
a=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['roofline']['whole_step_frac'], a['e2e']['value'])"
