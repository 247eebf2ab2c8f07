#!/bin/bash
# First GPU call of round 2 (one GPU): everything written after round 1's GPU budget was spent, in one go.
#   1. the whole GPU test-suite (replica engine, --gpu-mode, CLI through the new reader / inflate / writer)
#   2. compute-sanitizer memcheck over the small parity tests (never run in round 1)
#   3. bench.py at N=1 (unchanged kernels: must reproduce profiles/bench_r01.json)
#   4. tools/cli_e2e.py: FASTQ(.gz) -> TSV(.gz) with 1 / 4 / 8 inflate threads, compression levels, reader stage rates
#   5. tools/ingest_bench.py: the host stages alone on the box's CPUs (inflate, parse, writer, the command in --dry-run)
# usage: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check_r02a.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 gpurun_out/r02a_gpu_tests.log
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "generate_kmers or count_codes or search_batch_hits or engine_matches or long_reads or sketch_selection or degenerate" ) > gpurun_out/r02a_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02a_memcheck.log | tail -3
( time timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?"; cut -c1-400 gpurun_out/r02a_bench.json
( time timeout 400 python tools/cli_e2e.py 4000000 ) > gpurun_out/r02a_cli_e2e.json 2> gpurun_out/r02a_cli_e2e.err
echo "cli_e2e exit $?"; cat gpurun_out/r02a_cli_e2e.json
( time timeout 300 python tools/ingest_bench.py 1000000 ) > gpurun_out/r02a_ingest_box.json 2> gpurun_out/r02a_ingest_box.err
echo "ingest bench exit $?"; python -c "import json; a=json.load(open('gpurun_out/r02a_ingest_box.json')); print({k: a[k] for k in ('cpus', 'inflate', 'parse', 'search --dry-run')})"
