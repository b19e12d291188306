#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_tick$ -s 1000 -c 1 -f -o gpurun_out/r02_serial_65536_a python tools/prof_env.py 65536 1100 > gpurun_out/r2e_ncu1.log 2>&1
timeout 400 $NCU -k regex:k_tick_quad -s 2300 -c 1 -f -o gpurun_out/r02_quad_4096_a python tools/prof_env.py 4096 2400 > gpurun_out/r2e_ncu2.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r2e_tests.log
tail -3 gpurun_out/r2e_ncu1.log gpurun_out/r2e_ncu2.log; cat gpurun_out/r2e_tests.log
