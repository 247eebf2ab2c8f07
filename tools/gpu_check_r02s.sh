#!/bin/bash
# round 2, call S (KMCPG_DEV build): h > 1 probe with raw double buffering + shared row indices (KMCPG_PROBE_VARH=3) against the shipped VAR 1
mkdir -p gpurun_out
for cfg in "X=1" "KMCPG_PROBE_VARH=3"; do
  echo "== $cfg"
  env $cfg NG=85205 GL=100000 NR=100000 NCHK=200 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02s.err
  env $cfg NG=85205 GL=875000 NR=100000 NCHK=0 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02s.err
done 2>&1 | tee gpurun_out/r02s_varh3_ab.log
tail -3 gpurun_out/r02s.err
