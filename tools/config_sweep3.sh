#!/bin/bash
# NOTE: the KMCPG_PROBE_* knobs exist only in development builds of the library: make -C kmcp_b200/csrc clean all DEV=1
cd "$(dirname "$0")/.."
for cfg in "KMCPG_PROBE_VARH=1 KMCPG_PROBE_MINBH=3" "KMCPG_PROBE_VARH=0 KMCPG_PROBE_MINBH=3" "KMCPG_PROBE_VARH=1 KMCPG_PROBE_MINBH=2 KMCPG_PROBE_G=16" "KMCPG_PROBE_VARH=1 KMCPG_PROBE_MINBH=2 KMCPG_PROBE_CAP=32"; do
  echo "# h=3 one block: $cfg"; env $cfg H=3 NG=1000 GL=4000000 NR=500000 python tools/probe_one.py 2>&1 | tail -1
done
echo "# 300 bp reads (n=280 > 256 → warp sort+unique, 16 planes)"; NG=1000 GL=4000000 NR=400000 RL=300 python tools/probe_one.py 2>&1 | tail -1
echo "# 1500 bp reads"; NG=1000 GL=4000000 NR=60000 RL=1500 python tools/probe_one.py 2>&1 | tail -1
