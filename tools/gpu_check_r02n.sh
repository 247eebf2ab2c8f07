#!/bin/bash
# round 2, call N (product build): longest-first task order for long queries (tests + C5 block), the ShardedSearch single-rank test, and the
# GTDB-shape block at FULL size (genome length 3.5 Mb: ~100 GB index on one GPU) for the record
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "long_reads or count_codes or four_hash or sketch or dedup or degenerate or sharded_search_pipeline or full_size_c2" ) > gpurun_out/r02n_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r02n_tests.log
show() { python - "$1" <<'P'
import json,sys
a=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:a[k] for k in ('value','ms_per_step')}, a['roofline']['frac'], a['e2e']['value'])
for k in ('gtdb_scale','c5_hifi','c3_fracminhash'):
    g=a[k]
    if not g: continue
    print(k, {x:g[x] for x in g if x not in ('per_rank','digest_note','roofline')})
    if 'per_rank' in g: print([(round(r['probe_GBps']), round(r.get('prep_ms',0),1), r.get('build_s')) for r in g['per_rank']])
P
}
( time timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02n_bench.json 2> gpurun_out/r02n.err
show gpurun_out/r02n_bench.json
( time KMCP_GTDB_GL=3500000 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02n_bench_fullgtdb.json 2>> gpurun_out/r02n.err
show gpurun_out/r02n_bench_fullgtdb.json
tail -3 gpurun_out/r02n.err
