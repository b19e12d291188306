#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh r2an -x -k "multilink or other_cars or brake_disc or single_tick_parity_identical_states"
tools/quick_bench.sh 4096 > gpurun_out/r2an_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2an_q.log 2>&1
cat gpurun_out/r2an_q.log
