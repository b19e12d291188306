#!/bin/bash
mkdir -p gpurun_out
tools/quick_bench.sh 4096 > gpurun_out/r2l_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2l_q.log 2>&1
PD_QUAD_CPW=8 tools/quick_bench.sh 8192 >> gpurun_out/r2l_q.log 2>&1
tools/quick_bench.sh 1024 >> gpurun_out/r2l_q.log 2>&1
(timeout 300 python tools/phase_tail.py 4096 2000 2>&1 | tail -22) > gpurun_out/r2l_phase.log
cat gpurun_out/r2l_q.log gpurun_out/r2l_phase.log
bash tools/gpu_tests.sh r2l
