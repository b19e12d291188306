#!/bin/bash
bash tools/gpu_tests.sh r2h "$@"
tools/quick_bench.sh 4096 > gpurun_out/r2h_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2h_q.log 2>&1
cat gpurun_out/r2h_q.log
