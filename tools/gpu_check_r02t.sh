#!/bin/bash
# round 2, call T (KMCPG_DEV build): h > 1 probe with shared row indices, single buffer (KMCPG_PROBE_VARH=4) at 2 and 3 CTAs per SM against the shipped VAR 1
mkdir -p gpurun_out
for cfg in "KMCPG_PROBE_VARH=4 KMCPG_PROBE_MINBH=2" "KMCPG_PROBE_VARH=4 KMCPG_PROBE_MINBH=3" "X=1"; do
  echo "== $cfg"
  env $cfg NG=85205 GL=875000 NR=100000 NCHK=0 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02t.err
done 2>&1 | tee gpurun_out/r02t_varh4_ab.log
echo "== KMCPG_PROBE_VARH=4 KMCPG_PROBE_MINBH=2: oracle sample"
KMCPG_PROBE_VARH=4 KMCPG_PROBE_MINBH=2 NG=85205 GL=100000 NR=100000 NCHK=200 timeout 300 python tools/c4_shape.py 2>> gpurun_out/r02t.err | tee -a gpurun_out/r02t_varh4_ab.log
tail -3 gpurun_out/r02t.err
