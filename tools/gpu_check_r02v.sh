#!/bin/bash
# round 2, call V (product build): hand-written TSV number formatting in the writer — the CLI parity tests, then tools/cli_e2e.py on the box
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_zz_gpu_sharded.py -m gpu -x -q -k "cli or golden or demo" ) > gpurun_out/r02v_tests.log 2>&1
echo "cli tests exit $?"; grep -E "passed|failed|error" gpurun_out/r02v_tests.log | tail -3
( time timeout 400 python tools/cli_e2e.py 4000000 ) > gpurun_out/r02v_cli_e2e.json 2> gpurun_out/r02v_cli_e2e.err
echo "cli_e2e exit $?"; python - <<'P'
import json
a=json.load(open('gpurun_out/r02v_cli_e2e.json'))
for c in a['cases']: print(c['case'], c['reads'], c['wall_s'], c['search_reads_per_s'])
P
tail -3 gpurun_out/r02v_cli_e2e.err
