#!/bin/bash
# NOTE: the KMCPG_PROBE_* knobs exist only in development builds of the library: make -C kmcp_b200/csrc clean all DEV=1
# other BASELINE.json shapes through the same path (dev tool): h=3, many narrow blocks, long reads
cd "$(dirname "$0")/.."
echo "# C4-like: h=3, 8 blocks x 10k targets (fpr 0.3), 150 bp";  H=3 NG=8000 GL=400000 BS=10000 NR=200000 python tools/probe_one.py 2>&1 | tail -1
echo "# C2 with h=3, one block";                                    H=3 NG=1000 GL=4000000 NR=500000 python tools/probe_one.py 2>&1 | tail -1
echo "# C1-like: 150 targets in 10 blocks of 16 (2-byte rows)";     NG=15 GL=4000000 BS=16 NR=1000000 python tools/probe_one.py 2>&1 | tail -1
echo "# medium rows: 1000 targets/block (125 B rows), 10 blocks";   NG=1000 GL=1000000 BS=1000 NR=500000 python tools/probe_one.py 2>&1 | tail -1
echo "# C5-like: 10 kb reads vs the C2 index";                      NG=1000 GL=4000000 NR=10000 RL=10000 python tools/probe_one.py 2>&1 | tail -1
