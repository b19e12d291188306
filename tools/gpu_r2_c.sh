#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "single_tick or free_running or shard or ragged" 2>&1 | tail -15) > gpurun_out/r2c_tests.log
tools/quick_bench.sh 4096 > gpurun_out/r2c_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2c_q.log 2>&1
PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 65536 >> gpurun_out/r2c_q.log 2>&1
PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 16384 >> gpurun_out/r2c_q.log 2>&1
PD_QUAD_CPW=8 tools/quick_bench.sh 4096 >> gpurun_out/r2c_q.log 2>&1
PD_QUAD_CPW=8 tools/quick_bench.sh 8192 >> gpurun_out/r2c_q.log 2>&1
tools/quick_bench.sh 1024 >> gpurun_out/r2c_q.log 2>&1
(timeout 300 python tools/debug_offtrack.py 2>&1 | head -150) > gpurun_out/r2c_offtrack.log
cat gpurun_out/r2c_tests.log gpurun_out/r2c_q.log
