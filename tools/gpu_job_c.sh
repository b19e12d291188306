#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_autoreset.py 0 > gpurun_out/c_dbg.log 2>&1
timeout 300 python tools/debug_autoreset.py 1 >> gpurun_out/c_dbg.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_tick_quad -s 2300 -c 1 -f -o gpurun_out/r01_quad_v6_4096_steady python tools/prof_env.py 4096 2400 > gpurun_out/c_ncu1.log 2>&1
tail -5 gpurun_out/c_ncu1.log
cat gpurun_out/c_dbg.log
