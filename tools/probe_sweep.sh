#!/bin/bash
# NOTE: the KMCPG_PROBE_* knobs exist only in development builds of the library: make -C kmcp_b200/csrc clean all DEV=1
# sweeps the probe-kernel knobs on the GPU box; prints one JSON line per configuration
cd "$(dirname "$0")/.."
for cfg in "KMCPG_PROBE_VAR=0" "KMCPG_PROBE_VAR=1" "KMCPG_PROBE_VAR=1 KMCPG_PROBE_MINB=3" "KMCPG_PROBE_VAR=2" "KMCPG_PROBE_VAR=2 KMCPG_PROBE_CAP=8" \
           "KMCPG_PROBE_VAR=2 KMCPG_PROBE_CAP=32" "KMCPG_PROBE_VAR=2 KMCPG_PROBE_G=4" "KMCPG_PROBE_VAR=2 KMCPG_PROBE_G=16" "KMCPG_PROBE_VAR=2 KMCPG_PROBE_G=32" \
           "KMCPG_PROBE_VAR=1 KMCPG_PROBE_G=16" "KMCPG_PROBE_VAR=1 KMCPG_PROBE_MINB=3 KMCPG_PROBE_G=16" "KMCPG_PROBE_VAR=1 KMCPG_PROBE_MINB=3 KMCPG_PROBE_CAP=24"; do
  env $cfg python tools/probe_one.py 2>&1 | tail -1
done
