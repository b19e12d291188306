#!/bin/bash
mkdir -p gpurun_out
for e in 4096 65536; do timeout 120 tools/quick_bench.sh $e; done > gpurun_out/k_quick.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:'^k_tick$' -s 700 -c 1 -f -o gpurun_out/r01_serial_v7_65536 python tools/prof_env.py 65536 800 > gpurun_out/k_ncu3.log 2>&1
cat gpurun_out/k_quick.log; tail -3 gpurun_out/k_ncu3.log | cut -c1-200
