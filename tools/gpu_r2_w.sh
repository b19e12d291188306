#!/bin/bash
mkdir -p gpurun_out
PD_COLLIDE_V1=1 tools/quick_bench.sh 65536 > gpurun_out/r2w_q.log 2>&1
tools/quick_bench.sh 4096 >> gpurun_out/r2w_q.log 2>&1
tools/quick_bench.sh 1024 >> gpurun_out/r2w_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2w_q.log 2>&1
tools/quick_bench.sh 16384 >> gpurun_out/r2w_q.log 2>&1
tools/quick_bench.sh 65536 --car ks_mazda_rx7_tuned >> gpurun_out/r2w_q.log 2>&1
cat gpurun_out/r2w_q.log
bash tools/gpu_tests.sh r2w -k "collision or autoreset or single_tick_parity_identical_states or shard or ragged or other_cars"
