#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2x_q.log
for lpc in 1 4 8 16; do echo "LPC $lpc" >> gpurun_out/r2x_q.log; PD_COLLIDE_LPC=$lpc tools/quick_bench.sh 65536 >> gpurun_out/r2x_q.log 2>&1; PD_COLLIDE_LPC=$lpc tools/quick_bench.sh 16384 >> gpurun_out/r2x_q.log 2>&1; done
echo V1 >> gpurun_out/r2x_q.log; PD_COLLIDE_V1=1 tools/quick_bench.sh 16384 >> gpurun_out/r2x_q.log 2>&1
tools/quick_bench.sh 4096 >> gpurun_out/r2x_q.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_collide -s 20 -c 12 --csv --log-file gpurun_out/r2x_collide.csv python tools/prof_env.py 65536 200 > gpurun_out/r2x_ncu.log 2>&1
grep -o 'k_collide[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*"' gpurun_out/r2x_collide.csv | tail -3
cat gpurun_out/r2x_q.log
