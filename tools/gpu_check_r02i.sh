#!/bin/bash
# round 2, call I (KMCPG_DEV build): the TMA form of the probe (cp.async.bulk + mbarrier shared-memory ring) against the register kernel on
# wide rows; compute-sanitizer racecheck + memcheck on the small parity tests
mkdir -p gpurun_out
for b in 0 1; do
  echo "== KMCPG_PROBE_BULK=$b: C4 shape, 3 GB index (oracle sample), then 26 GB index"
  KMCPG_PROBE_BULK=$b NG=85205 GL=100000 NR=100000 NCHK=200 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02i.err
  KMCPG_PROBE_BULK=$b NG=85205 GL=875000 NR=100000 NCHK=0 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02i.err
  echo "== KMCPG_PROBE_BULK=$b: C2"
  KMCPG_PROBE_BULK=$b timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gtdb 2>> gpurun_out/r02i.err | python -c "
import json,sys
a=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, 'frac', round(a['roofline']['frac'],4), 'launch_ms', round(a['roofline']['avg_launch_ms'],3), 'hits', a['config']['hits_per_step'])"
done 2>&1 | tee gpurun_out/r02i_tma_ab.log
( time KMCPG_PROBE_BULK=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_c2 or c4_shape or engine_matches or search_batch_hits" ) > gpurun_out/r02i_bulk_tests.log 2>&1
echo "bulk parity tests exit $?"; tail -4 gpurun_out/r02i_bulk_tests.log
( time timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "search_batch_hits or degenerate" ) > gpurun_out/r02i_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/r02i_racecheck.log | tail -4
( time timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q \
    -k "four_hash or multi_k or jobs_in_flight or two_host_threads or row_index or generate_kmers or count_codes or long_reads or sketch_selection" ) > gpurun_out/r02i_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02i_memcheck.log | tail -4
tail -5 gpurun_out/r02i.err
