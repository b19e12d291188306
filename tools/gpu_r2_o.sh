#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh r2o -x
tools/quick_bench.sh 4096 > gpurun_out/r2o_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2o_q.log 2>&1
tools/quick_bench.sh 65536 --car ks_mazda_rx7_tuned >> gpurun_out/r2o_q.log 2>&1
tools/quick_bench.sh 4096 --car dthwsh_mazda_rx7_fc3s_sr20 >> gpurun_out/r2o_q.log 2>&1
cat gpurun_out/r2o_q.log
