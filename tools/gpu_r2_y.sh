#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh r2y -x -k "other_cars or brake_disc"
: > gpurun_out/r2y_q.log
tools/quick_bench.sh 4096 >> gpurun_out/r2y_q.log 2>&1
tools/quick_bench.sh 4096 --car ks_mazda_rx7_tuned >> gpurun_out/r2y_q.log 2>&1
tools/quick_bench.sh 4096 --car dthwsh_mazda_rx7_fc3s_sr20 >> gpurun_out/r2y_q.log 2>&1
tools/quick_bench.sh 8192 --car ks_toyota_supra_mkiv_drift >> gpurun_out/r2y_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2y_q.log 2>&1
cat gpurun_out/r2y_q.log
