#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "single_tick_parity_identical or free_running or shard or ragged" 2>&1 | tail -15) > gpurun_out/r2d_tests.log
tools/quick_bench.sh 4096 > gpurun_out/r2d_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2d_q.log 2>&1
PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 65536 >> gpurun_out/r2d_q.log 2>&1
PD_QUAD_CPW=8 tools/quick_bench.sh 4096 >> gpurun_out/r2d_q.log 2>&1
PD_QUAD_CPW=8 tools/quick_bench.sh 8192 >> gpurun_out/r2d_q.log 2>&1
tools/quick_bench.sh 1024 >> gpurun_out/r2d_q.log 2>&1
(timeout 300 python tools/phase_tail.py 4096 2000 2>&1 | tail -50) > gpurun_out/r2d_phase.log
cat gpurun_out/r2d_tests.log gpurun_out/r2d_q.log; tail -24 gpurun_out/r2d_phase.log
