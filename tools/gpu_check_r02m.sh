#!/bin/bash
# round 2, call M (product build): the whole GPU suite, smoke(), the reference arm (helper-process staging), the ncu launch list of a bench step
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02m_tests.log 2>&1
echo "gpu tests exit $?"; tail -4 gpurun_out/r02m_tests.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -3
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02m_reference_arm.json 2> gpurun_out/r02m_reference_arm.err
echo "reference arm exit $?"; cut -c1-700 gpurun_out/r02m_reference_arm.json; tail -3 gpurun_out/r02m_reference_arm.err
( time timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(probe_kernel|hash_kernel|locs_kernel|pack_hits_kernel|DeviceRadixSort.*)$' -s 120 -c 140 \
    --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gtdb ) > gpurun_out/r02m_ncu_list.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_r02.csv; tail -2 gpurun_out/r02m_ncu_list.log | cut -c1-200
