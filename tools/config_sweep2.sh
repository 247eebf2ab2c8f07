#!/bin/bash
# NOTE: the KMCPG_PROBE_* knobs exist only in development builds of the library: make -C kmcp_b200/csrc clean all DEV=1
cd "$(dirname "$0")/.."
for cfg in "KMCPG_PROBE_MINBH=1" "KMCPG_PROBE_MINBH=2" "KMCPG_PROBE_VARH=1 KMCPG_PROBE_MINBH=2" "KMCPG_PROBE_VARH=1 KMCPG_PROBE_MINBH=1" "KMCPG_PROBE_VARH=0"; do
  echo "# h=3 one block: $cfg"; env $cfg H=3 NG=1000 GL=4000000 NR=500000 python tools/probe_one.py 2>&1 | tail -1
done
echo "# C5-like: 10 kb reads vs the C2 index (16 planes: 8 in registers + totals in smem)"; NG=1000 GL=4000000 NR=10000 RL=10000 python tools/probe_one.py 2>&1 | tail -1
echo "# 300 bp reads (n=280 > 255 → 16 planes + dedup sort)"; NG=1000 GL=4000000 NR=400000 RL=300 python tools/probe_one.py 2>&1 | tail -1
