#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/j_tests.log
for e in 1024 4096 16384 65536; do timeout 120 tools/quick_bench.sh $e; done > gpurun_out/j_quick.log 2>&1
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/j_bench.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_tick_quad -s 2300 -c 1 -f -o gpurun_out/r01_quad_v7_4096 python tools/prof_env.py 4096 2400 > gpurun_out/j_ncu1.log 2>&1
timeout 400 $NCU -k regex:k_collide -s 1150 -c 1 -f -o gpurun_out/r01_collide_v7_4096 python tools/prof_env.py 4096 2400 > gpurun_out/j_ncu2.log 2>&1
timeout 500 $NCU -k regex:'k_tick\(' -s 700 -c 1 -f -o gpurun_out/r01_serial_v7_65536 python tools/prof_env.py 65536 800 > gpurun_out/j_ncu3.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/r01_launches_v7_4096.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu4.log 2>&1
cat gpurun_out/j_tests.log gpurun_out/j_quick.log; cut -c1-600 gpurun_out/j_bench.json; tail -2 gpurun_out/j_ncu1.log gpurun_out/j_ncu2.log gpurun_out/j_ncu3.log | cut -c1-200
