#!/bin/bash
# round-2 evidence: ncu --set full of the two tick kernels at the current code, launch lists, phase clocks, then the round-end sequence
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_tick_quad -s 2301 -c 1 -f -o gpurun_out/r02_quad_4096_odd python tools/prof_env.py 4096 2400 > gpurun_out/p_ncu1.log 2>&1
timeout 400 $NCU -k regex:k_tick_quad -s 2300 -c 1 -f -o gpurun_out/r02_quad_4096_even python tools/prof_env.py 4096 2400 > gpurun_out/p_ncu1b.log 2>&1
timeout 500 $NCU -k regex:^k_tick\$ -s 1300 -c 1 -f -o gpurun_out/r02_serial_65536 python tools/prof_env.py 65536 1400 > gpurun_out/p_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2050 -c 200 --csv --log-file gpurun_out/r02_launches_4096.csv python tools/prof_env.py 4096 2300 > gpurun_out/p_ncu3.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1950 -c 150 --csv --log-file gpurun_out/r02_launches_65536.csv python tools/prof_env.py 65536 1400 > gpurun_out/p_ncu4.log 2>&1
timeout 200 python tools/phase_tail.py 4096 2000 > gpurun_out/p_phase.log 2>&1
bash tools/gpu_round_end.sh
