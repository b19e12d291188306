#!/bin/bash
mkdir -p gpurun_out
tools/quick_bench.sh 65536 > gpurun_out/r2n_q.log 2>&1
PD_SERIAL_INLINE_COLLIDE=1 tools/quick_bench.sh 65536 >> gpurun_out/r2n_q.log 2>&1
PD_SERIAL_INLINE_COLLIDE=1 tools/quick_bench.sh 16384 >> gpurun_out/r2n_q.log 2>&1
cat gpurun_out/r2n_q.log
