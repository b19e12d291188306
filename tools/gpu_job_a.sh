#!/bin/bash
# ncu captures of the two tick kernels in the env-step workload + launch list of the bench + e2e breakdown
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_tick_quad -s 400 -c 1 -f -o gpurun_out/r01_quad_v5_4096 python tools/prof_env.py 4096 260 > gpurun_out/a_ncu1.log 2>&1
timeout 300 $NCU -k regex:'k_tick$' -s 120 -c 1 -f -o gpurun_out/r01_serial_v5_65536 python tools/prof_env.py 65536 100 > gpurun_out/a_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_4096.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu3.log 2>&1
timeout 120 python tools/prof_env.py 4096 990 e2e > gpurun_out/a_e2e.log 2>&1
timeout 120 python tools/prof_env.py 1024 990 e2e >> gpurun_out/a_e2e.log 2>&1
tools/quick_bench.sh 4096 > gpurun_out/a_quick.log 2>&1
tools/quick_bench.sh 4096 --steps 30 >> gpurun_out/a_quick.log 2>&1
cat gpurun_out/a_e2e.log gpurun_out/a_quick.log; tail -3 gpurun_out/a_ncu1.log gpurun_out/a_ncu2.log
